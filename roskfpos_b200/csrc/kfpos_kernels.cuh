// kfpos_kernels.cuh -- kernel parameter blocks and launch entry points shared
// between the kernels (*.cu) and the C-ABI layer (kfpos_api.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kfpos_math.cuh"

namespace kfpos {


// Range stream of a replay: SoA [T][M][N] in `fmt`, optional per-ranging
// errorEstimation of the same shape.
struct RangeStream {
    const void *ranges;
    const double *err; // null -> err_scalar
    double err_scalar;
    int fmt;
    int m_slots;
};

// ---- error statistics with a summation order that depends on nothing but the filter index:
// partial c covers filters [128c, 128c + 128) -- one replay block -- summed by a fixed 128-leaf binary
// tree in shared memory; the partials are then folded by a pairwise tree over the partial index
// (launch_error_stats_tree).  Every replay kernel can leave the partials of its final state behind
// (params.truth / params.partials): the reduction is fused into the last step of the replay, and a
// shard that starts at a multiple of its power-of-two size is a complete subtree of the global tree, so
// the reduced result is BIT-IDENTICAL for 1, 2, 4 or 8 GPUs (kfpos_stats_allreduce).
constexpr int STATS_CHUNK = 128;
// v = {|e|^2, |e_xy|^2, 1, bad} of this thread's filter (zeros for a thread without one); `sh` = at least
// 4 * 128 doubles of shared memory nobody else is using any more; every thread of the block calls it
KF_DEV void block_stats_partial(const double (&v)[4], double *sh, double *partial_out) {
    const int t = threadIdx.x;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) sh[q * STATS_CHUNK + t] = v[q];
    __syncthreads();
    for (int s = STATS_CHUNK / 2; s > 0; s >>= 1) {
        if (t < s) {
#pragma unroll
            for (int q = 0; q < 4; ++q) sh[q * STATS_CHUNK + t] += sh[q * STATS_CHUNK + t + s];
        }
        __syncthreads();
    }
    if (t < 4) partial_out[t] = sh[t * STATS_CHUNK];
}
KF_DEV void filter_error_terms(double px, double py, double pz, const double *truth, int64_t N, int64_t f, bool bad,
                               double (&v)[4]) {
    const double ex = px - truth[f], ey = py - truth[N + f], ez = pz - truth[2 * N + f];
    const double e2 = ex * ex + ey * ey + ez * ez;
    v[0] = v[1] = v[2] = 0.0;
    if (isfinite(e2)) {
        v[0] = e2;
        v[1] = ex * ex + ey * ey;
        v[2] = 1.0;
    }
    v[3] = bad ? 1.0 : 0.0;
}

struct T6Params {
    AnchorTable anchors;
    RangeStream rs;
    int64_t N;
    int T;
    int ignore_worst;
    int variant, n_ignore, best_mode; // EKF-side NLOS variants (config_pos.xml; 0 = normal)
    double ignore_thr;
    double accel_noise;
    const double *dt;   // device [T], common to the batch
    const double *dt_f; // or null: per-filter SoA [T][N]; a value < 0 = this filter has no epoch t
    double *x;        // SoA [6][N]  (rows 3..5 stay 0)
    double *P;        // SoA [21][N] packed lower triangle
    int32_t *status;  // [N], OR-ed
    double *traj;     // SoA [T][3][N] or null
    int32_t *sel;     // SoA [T][N] or null
    unsigned long long *counters;
    const double *truth; // SoA [3][N] or null: leave the error partials of the final state in `partials`
    double *partials;    // [ceil(N / 128)][4]
};

struct MlParams {
    AnchorTable anchors;
    RangeStream rs;
    int64_t N;
    int use2d, variant, n_ignore, best_mode;
    int zero_tz, _pad; // ml2d_zero_tentative_z
    double start[3];
    double min_z, max_z; // output gate (config_pos.xml minZ / maxZ), active when max_z > min_z
    double *pos;     // SoA [3][N] or null
    double *cov;     // SoA [9][N] or null
    int32_t *iters;  // [N] or null
    int32_t *sel;    // SoA [2][N] or null
    int32_t *status; // [N] or null
    unsigned long long *counters;
    // straggler queues (kfpos_mlk.cu): two buffers of queue_cap 64-byte records + their counters;
    // launch_ml_solve points q_in / q_out at them launch by launch
    void *queue[2];
    int *queue_count; // device int[2]
    int queue_cap;
    const void *q_in;      // records this launch resumes / advances (RESUME and cooperative kernels)
    const int *q_in_count;
    void *q_out;           // where this launch parks; null = finish in place
    int *q_out_count;
    unsigned first_cap;    // Newton iterations a solve gets in this launch before it is parked
    int coop_min, coop_max; // queue lengths [min, max) for which a cooperative launch does its work
    // exact-order solver (kfpos_exact.cu): kfpos_config.ml_exact_order, and the queue of epoch indices
    // whose residual order the fast solver found too close to a tie to decide (variant 1)
    int exact_mode;
    int xq_cap;
    int32_t *xq;
    int *xq_count;
    // scratch of the warp-per-epoch BestGroup kernels (parked 3-D subset solves), see ml_exact_scratch_bytes
    void *xw_scratch;
    size_t xw_scratch_bytes;
    int *stream_counter; // device int: the epoch counter of ml_stream3_kernel
};
constexpr int KFPOS_MAX_ANCHORS_DEV = 32;
// relative margin below which the fast solver does not trust its own order of two squared residuals
// (its positions agree with the reference's to ~1e-11: four orders of magnitude of room)
constexpr double ML_TIE_MARGIN = 1e-6;

// ---- sensor event streams (K8, T9).  The schedule (kind, dt) is common to the batch;
// `offset` is the first ROW (units of N elements) of the event's payload: rows of the
// range tensor for EV_TOA, rows of the f64 `sensors` tensor otherwise:
//   EV_PX4: integration_x, integration_y, integration_rot_z, integration_time_us, quality
//   EV_IMU: K8: ang_vel_z, lin_acc_x, lin_acc_y   T9: lin_acc_x, lin_acc_y, lin_acc_z
//   EV_MAG: mag_x, mag_y      EV_COMPASS: compass (rad)
// aux: EV_IMU covariances, common to the batch: K8: cov_acc[0],[1],[3],[4], cov_ang_vel[8];
// T9: cov_acc[0..8].  Mirrors kfpos_event of include/kfpos_b200.h.
enum { EV_TOA = 0, EV_PX4 = 1, EV_IMU = 2, EV_MAG = 3, EV_COMPASS = 4 };
struct EventDesc {
    int32_t kind;
    int32_t _pad;
    double dt;
    int64_t offset;
    double aux[9];
};

// constants of the five XML files + launch parameters (KF.cpp:766-844, node_pos.cpp:48-113)
struct K8Cfg {
    double accel_noise, jolt;
    double tag_z;                  // mUWBtagZ
    double px4_height, arm1, arm2; // mPX4flowHeight, mPX4FlowArmP1/P2
    double px4_cov_vel, px4_cov_gyro;
    double imu_cov_acc, imu_cov_gyro;
    double mag_offset, mag_cov;
    int imu_fix_acc, imu_fix_gyro;
    int variant, n_ignore, best_mode; // EKF-side NLOS variants (config_pos.xml; 0 = normal)
    int zero_tz;                      // ml2d_zero_tentative_z (App. B-1 as a zero-initialising build computes it)
    // the constructor without initialPosition (kfpos_config.ml_initial_position, KF.cpp:244-285): a filter whose
    // position is NaN initialises itself from its first epoch with rangings -- 2-D ML from (1, 1, tag height)
    // when use_fixed_height, else 3-D ML from (1, 1, 4), whose z becomes THIS filter's tag height (latch row 9)
    int ml_init, use_fixed_height;
};

struct K8Params {
    AnchorTable anchors;
    K8Cfg cfg;
    RangeStream rs;
    const double *sensors;   // SoA rows [R][N] f64
    const EventDesc *events; // device [n_events]
    int n_events;
    int64_t N;
    double *x;       // SoA [8][N]
    double *P;       // SoA [36][N] packed
    int32_t *status; // [N]
    double *latch;   // SoA [16][N]: px4 vx,vy,gz,cv | imu ax,ay,wz | mag angle | [8] dt carry | [9] tag height
    int32_t *has;    // [N] bit0 px4, bit1 imu, bit2 mag latched
    int *uninit;     // null, or a flag set by every filter that is still uninitialised when the launch ends
    double *latch_u; // batch-wide latched IMU covariances c00,c01,c11,cw: read from [0..3], written to [16..19]
    const double *dt_f; // null, or per-filter time steps SoA [n_events][N] (see T9Params)
    double *traj;    // SoA [n_toa][3][N] (px, py, theta) after each TOA event, or null
    unsigned long long *counters;
    const double *truth; // see T6Params
    double *partials;
};

cudaError_t launch_t6_replay(const T6Params &p, cudaStream_t s);
cudaError_t launch_k8_replay(const K8Params &p, cudaStream_t s);
cudaError_t launch_k8_get_pose(int64_t N, double dt, double accel_noise, double jolt, const double *x,
                               const double *P, double *x_pred, double *P_pred_full, cudaStream_t s);
cudaError_t launch_ml_solve(const MlParams &p, cudaStream_t s);
cudaError_t launch_ml_exact(const MlParams &p, bool queued, cudaStream_t s); // kfpos_exact.cu
size_t ml_exact_scratch_bytes(int64_t epochs);
int64_t ml_exact_scratch_epochs(int64_t N);
cudaError_t launch_selftest_ieee(int64_t n, const double *a, const double *b, double *div_fast, double *div_ieee,
                                 double *sqrt_fast, double *sqrt_ieee, int32_t *flags, cudaStream_t s);

struct T9Params {
    AnchorTable anchors;
    double accel_noise, jolt;
    RangeStream rs;
    const double *sensors;
    const EventDesc *events;
    int n_events;
    int64_t N;
    double *x;       // SoA [9][N]
    double *P;       // SoA [45][N] packed
    int32_t *status; // [N]
    double *latch;   // SoA rows 0..2: latched acceleration
    int32_t *has;    // [N] bit1: imu latched
    double *latch_u; // batch-wide latched 3x3 acceleration covariance: read from [0..8], written to [16..24]
    const double *dt_f; // null, or per-filter time steps SoA [n_events][N] replacing EventDesc::dt (< 0: the
                        // filter has no such event) -- the ragged epochs of kfpos_batch_replay_epochs
    int no_imu;      // host knowledge: no IMU sample latched and none in this schedule -> lean kernel
    int variant, n_ignore, best_mode; // EKF-side NLOS variants (config_pos.xml; 0 = normal)
    int ml_init;     // the constructor without initialPosition (TOAIMU.cpp:118-162), see K8Cfg::ml_init
    int *uninit;     // see K8Params
    double *traj;    // SoA [n_toa][3][N] or null
    unsigned long long *counters;
    const double *truth; // see T6Params
    double *partials;
};
cudaError_t launch_t9_replay(const T9Params &p, cudaStream_t s);
cudaError_t launch_t9_get_pose(int64_t N, double dt, double jolt, const double *x, const double *P, double *x_pred,
                               double *P_pred_full, cudaStream_t s);
cudaError_t launch_i32_to_f64(int64_t N, const int32_t *in, double *out, cudaStream_t s);

// full row-major SoA [n*n][N]  <->  packed lower-triangle SoA [n(n+1)/2][N]
cudaError_t launch_pack_cov(int n, int64_t N, const double *full, double *packed, cudaStream_t s);
cudaError_t launch_unpack_cov(int n, int64_t N, const double *packed, double *full, cudaStream_t s);

// predict-only pose poll (getPose)
cudaError_t launch_t6_get_pose(int64_t N, double dt, double accel_noise, const double *x,
                               const double *P, double *x_pred, double *P_pred_full, cudaStream_t s);

// error statistics (see block_stats_partial): partial[c][4] over the 128-filter chunks -- skipped when
// `truth` is null, i.e. when a replay kernel has already left them -- then the pairwise tree -> out[4]
// (position = rows 0,1 and row `zrow`, or the constant `zconst` when zrow < 0)
cudaError_t launch_error_stats(int64_t N, const double *x, int zrow, double zconst, const int32_t *status,
                               const double *truth, double *partials, double *out4, cudaStream_t s);
// pairwise tree over n vectors of 4 doubles (IN PLACE on `part`): level by level part[i] += part[i + stride]
cudaError_t launch_error_stats_tree(int64_t n, double *part, double *out4, cudaStream_t s);

// getPose report in the publisher's layout (kfpos_misc.cu); model 1 = T6, 2 = K8, 3 = T9
// tagz: null, or the per-filter tag height replacing tag_z (K8 after a 3-D ML initialisation)
cudaError_t launch_pose_msg(int model, int64_t N, double tag_z, const double *tagz, const double *x_pred,
                            const double *P_pred_full, double *pose13, double *cov36, cudaStream_t s);
cudaError_t launch_fill(double *p, int64_t n, double v, cudaStream_t s);

// epoch assembler (kfpos_assemble.cu): N logs of L messages, SoA [L][N]
struct AssembleParams {
    int64_t N, L, max_epochs;
    int M, fix_b12;
    double first_dt;
    const uint8_t *anchor, *seq;
    const int32_t *range_mm;
    const double *err, *t;
    int32_t *tbl_r; // [rows][M][N], rows = 256 (as written) or 1 (fix_b12), pre-set to -1
    double *tbl_e;  // same shape, pre-set to 0
    int32_t *ranges_out;
    double *err_out, *dt_out;
    int32_t *n_epochs;
    double *t_out; // [max_epochs][N] or null: the time of every report (-1 = no such epoch)
};
cudaError_t launch_assemble(const AssembleParams &p, cudaStream_t s);

// stream merger (kfpos_assemble.cu): stream 0 = ranging reports, 1..4 = PX4 / IMU / MAG / COMPASS samples
struct MergeParams {
    int64_t N;
    int M, n_slots;
    int64_t L[5];
    const double *t_src[5];   // [L_k][N] arrival times; negative / NaN ends the stream
    const int32_t *ranges;    // [L_0][M][N]
    const double *err_src;    // [L_0][M][N] or null
    const double *src[5];     // k > 0: [L_k][rows_k][N]
    const int32_t *slot_kind; // device [n_slots]
    const int64_t *slot_row;  // device [n_slots]: first output row of the slot
    double first_dt;
    double *dt_f;             // [n_slots][N]
    int32_t *ranges_out;      // [range rows][N]
    double *err_out;          // [range rows][N] or null
    double *sensors_out;      // [sensor rows][N]
    int32_t *n_dropped;       // [N] or null
};
cudaError_t launch_merge(const MergeParams &p, cudaStream_t s);

cudaError_t launch_selftest_math(int64_t n, const double *x, double *rcp, double *rsq, double *sn, double *cs,
                                 cudaStream_t s);

// Monte Carlo input generator for K8 batches (kfpos_synth.cu)
struct SynthEvent {
    int32_t kind;
    int32_t global_index; // index of the event in the whole run (RNG counter)
    double t;             // time of the event, s
    int64_t offset;       // first output row (rows of `ranges` for EV_TOA, of `sensors` otherwise)
};
struct SynthK8Params {
    AnchorTable anchors;
    int64_t N, filter0;
    uint64_t seed;
    int M, n_events;
    double tag_z, sigma_r, t_end;
    const SynthEvent *events; // device [n_events]
    int32_t *ranges;          // SoA [rows][N] int32 mm
    double *sensors;          // SoA [rows][N]
    double *x0;               // SoA [8][N] or null
    double *truth_end;        // SoA [3][N] or null: (x, y, tag z) at t_end
};
cudaError_t launch_synth_k8(const SynthK8Params &p, cudaStream_t s);

// DFMA-only microbenchmark on the current device (the FP64 roofline denominator)
cudaError_t measure_fp64_peak(double *flops_per_s);

} // namespace kfpos
