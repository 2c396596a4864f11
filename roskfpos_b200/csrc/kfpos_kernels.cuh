// kfpos_kernels.cuh -- kernel parameter blocks and launch entry points shared
// between the kernels (*.cu) and the C-ABI layer (kfpos_api.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kfpos_math.cuh"

namespace kfpos {

// device counters (unsigned long long each)
enum { CNT_UPDATES = 0, CNT_ML_ITERS, CNT_COST_EVALS, CNT_GAIN_EVALS, CNT_BAD, CNT_IGNORED, CNT_N = 8 };

// Range stream of a replay: SoA [T][M][N] in `fmt`, optional per-ranging
// errorEstimation of the same shape.
struct RangeStream {
    const void *ranges;
    const double *err; // null -> err_scalar
    double err_scalar;
    int fmt;
    int m_slots;
};

struct T6Params {
    AnchorTable anchors;
    RangeStream rs;
    int64_t N;
    int T;
    int ignore_worst;
    double ignore_thr;
    double accel_noise;
    const double *dt; // device [T]
    double *x;        // SoA [6][N]  (rows 3..5 stay 0)
    double *P;        // SoA [21][N] packed lower triangle
    int32_t *status;  // [N], OR-ed
    double *traj;     // SoA [T][3][N] or null
    int32_t *sel;     // SoA [T][N] or null
    unsigned long long *counters;
};

struct MlParams {
    AnchorTable anchors;
    RangeStream rs;
    int64_t N;
    int use2d, variant, n_ignore, best_mode;
    double start[3];
    double *pos;     // SoA [3][N] or null
    double *cov;     // SoA [9][N] or null
    int32_t *iters;  // [N] or null
    int32_t *sel;    // SoA [2][N] or null
    int32_t *status; // [N] or null
    unsigned long long *counters;
};

cudaError_t launch_t6_replay(const T6Params &p, cudaStream_t s);
cudaError_t launch_ml_solve(const MlParams &p, cudaStream_t s);

// full row-major SoA [n*n][N]  <->  packed lower-triangle SoA [n(n+1)/2][N]
cudaError_t launch_pack_cov(int n, int64_t N, const double *full, double *packed, cudaStream_t s);
cudaError_t launch_unpack_cov(int n, int64_t N, const double *packed, double *full, cudaStream_t s);

// predict-only pose poll (getPose)
cudaError_t launch_t6_get_pose(int64_t N, double dt, double accel_noise, const double *x,
                               const double *P, double *x_pred, double *P_pred_full, cudaStream_t s);

// error statistics: partial[chunk][4] over fixed 1024-filter chunks, then a
// fixed-shape tree in the second kernel -> out[4]
cudaError_t launch_error_stats(int64_t N, const double *x, const int32_t *status, const double *truth,
                               double *partials, double *out4, cudaStream_t s);

// DFMA-only microbenchmark on the current device (the FP64 roofline denominator)
cudaError_t measure_fp64_peak(double *flops_per_s);

} // namespace kfpos
