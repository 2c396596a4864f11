/*
 * kfpos_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, dense, as-written restatement of the arithmetic of the roskfpos
 * filter core (reference: /root/reference/src/kfpos/algorithms/*.cpp).  Every
 * function cites the reference file:line it follows.  It is used ONLY by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs as the checker / the CPU baseline.  The product (the CUDA
 * library behind include/kfpos_b200.h) never links, loads or calls it.
 *
 * Pinning status: see the PINNING paragraph of oracle/README.md (kept in one
 * place so that it cannot drift from what the tests actually check).
 *
 * Restatement decisions for reference defects are those of SURVEY.md App. B
 * (B-1 .. B-11) and are repeated at each site.
 */
#ifndef KFPOS_ORACLE_H
#define KFPOS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KO_MAX_ANCHORS 32
#define KO_MAX_ROWS (KO_MAX_ANCHORS + 8)

/* one valid ranging: RangingMeasurement of sensor_types.h:27-32 flattened */
typedef struct {
    double r;          /* ranging, metres                    */
    double e;          /* errorEstimation (variance-like)    */
    double bx, by, bz; /* beacon.position                    */
    int slot;          /* original anchor slot (for reports) */
} ko_meas;

/* status bits shared with the CUDA library (include/kfpos_b200.h) */
#define KO_ST_OK 0
#define KO_ST_NO_MEAS 1      /* no valid ranging in a TOA update                 */
#define KO_ST_ML_FEW 2       /* inner ML had < minRangings, returned its start   */
#define KO_ST_SINGULAR 4     /* a solve/inv hit an exactly singular matrix       */
#define KO_ST_NAN 8          /* non-finite state after the update                */
#define KO_ST_ML_NAN 16      /* T6 NaN guard fired (TOA.cpp:270-272)             */
#define KO_ST_MAXITER 32     /* IEKF ran out of iterations without the break     */
#define KO_ST_UNINIT 256     /* K8/T9 without a fixed initial position: the event met a filter whose
                                position was still NaN and ended in the ML-initialisation branch
                                (KF.cpp:244-285, TOAIMU.cpp:118-162): no predict, no update        */

typedef struct {
    int status;
    int ml_iters;   /* Newton iterations of the inner ML solve   */
    int cost_evals; /* IEKF cost evaluations (I_c)               */
    int gain_evals; /* IEKF gain computations (I_g)              */
    int ignored;    /* T6 leave-one-out: ignored slot or -1      */
    double cost;    /* last assigned IEKF cost                   */
} ko_info;

/* ---------------------------------------------------------------- dense */
int ko_inv(int n, const double *A, double *Ainv);                 /* arma::inv  */
void ko_pinv(int n, const double *A, double *Apinv);              /* arma::pinv */
int ko_solve(int n, const double *A, const double *b, double *x, int equilibrate);
void ko_svd_jacobi(int n, const double *A, double *U, double *S, double *V);

/* ------------------------------------------------------------------- ML */
double ko_sse(const ko_meas *m, int n, const double p[3]);
void ko_dist(const ko_meas *m, int n, const double p[3], double *d);
/* return: 0 ok, 1 too few rangings (pos = start, cov untouched), -1 singular */
int ko_ml2d(const ko_meas *m, int n, const double start[3], int b1_zero_z,
            double pos[3], double cov[4], int *iters);
int ko_ml3d(const ko_meas *m, int n, const double start[3], double pos[3], double cov[9],
            int *iters);
/* order[] = innerIndex sorted by squared residual ascending (ties: lower index first) */
void ko_best_rangings(const ko_meas *m, int n, const double p[3], int *order);
int ko_ml_ignore_n(const ko_meas *m, int n, const double start[3], int use2d, int n_ignore,
                   int b1_zero_z, double pos[3], double *cov, int *iters, int *order,
                   int *n_dropped);
int ko_ml_best_group(const ko_meas *m, int n, const double start[3], int use2d, int best_mode,
                     int b1_zero_z, double pos[3], double *cov, int *iters, int *best_index,
                     uint32_t *best_mask, int *n_groups);
/* MLLocation::getPose dispatch on stored measurements; raw inputs as newTOAMeasurement */
int ko_ml_epoch(int n_slots, const double *ranges, const double *anchors /*[n][3]*/,
                const double *errs, const double start[3], int use2d, int variant,
                int n_ignore, int best_mode, int b1_zero_z, double pos[3], double cov[9],
                int *iters, int32_t *sel /* [2]: mask, index */);

/* ------------------------------------------------------------------- T6 */
typedef struct {
    double accel_noise;
    int ignore_worst;
    double ignore_cost_threshold;
    double pos[3];
    double vel[3]; /* stays 0: never written back (TOA.cpp:110-112,159-183) */
    double P[36];
} ko_t6;
void ko_t6_init(ko_t6 *f, double accel_noise, int ignore_worst, double thr, const double p0[3]);
void ko_t6_new_toa(ko_t6 *f, double dt, int n_slots, const double *ranges,
                   const double *anchors, const double *errs, ko_info *info);
void ko_t6_new_toa_sel(ko_t6 *f, double dt, int n_slots, const double *ranges, const double *anchors,
                       const double *errs, int variant, int n_ignore, int best_mode, ko_info *info,
                       uint32_t *mask_out);
void ko_t6_get_pose(const ko_t6 *f, double dt, double pos[3], double Ppred[36]);

/* ------------------------------------------------------------------- K8 */
typedef struct {
    /* launch params */
    double accel_noise, jolt;
    /* XML (KF.cpp:766-844) */
    double tag_z;
    int use_fixed_height;
    double px4_height, px4_arm1, px4_arm2, px4_cov_vel, px4_cov_gyro;
    int imu_fixed_cov_acc;
    double imu_cov_acc;
    int imu_fixed_cov_gyro;
    double imu_cov_gyro;
    double mag_offset, mag_cov;
    /* members */
    double pos[2], vel[2], acc[2], angle, omega;
    double P[64];
    int has_mag, has_px4, has_imu;
    double px4_itime, px4_vx, px4_vy, px4_gz, px4_cv, px4_cg; /* lastPX4FlowMeasurement */
    double imu_wz, imu_cwz, imu_ax, imu_ay, imu_cxy[4];       /* lastImuMeasurement     */
    double mag_angle, mag_c;                                  /* lastMagMeasurement     */
    /* EKF-side NLOS variants (config_pos.xml; see ko_t6_new_toa_sel): 0 = normal */
    int variant, n_ignore, best_mode;
    /* 1 = the constructor WITHOUT initialPosition (KF.cpp:6-32, mUseFixedInitialPosition = false): while
     * pos is NaN every event ends in the ML-initialisation branch (KF.cpp:244-285) */
    int ml_init;
} ko_k8;
void ko_k8_init(ko_k8 *f, double accel_noise, double init_angle, double jolt, const double p0[2]);
void ko_k8_new_toa(ko_k8 *f, double dt, int n_slots, const double *ranges,
                   const double *anchors, const double *errs, int b1_zero_z, ko_info *info);
void ko_k8_new_px4(ko_k8 *f, double dt, double ix, double iy, double irz, double itime_us,
                   int quality, ko_info *info);
void ko_k8_new_imu(ko_k8 *f, double dt, const double angvel[3], const double cov_av[9],
                   const double acc[3], const double cov_acc[9], ko_info *info);
void ko_k8_new_mag(ko_k8 *f, double dt, const double mag[3], ko_info *info);
void ko_k8_new_compass(ko_k8 *f, double dt, double compass, ko_info *info);
void ko_k8_get_pose(const ko_k8 *f, double dt, double x[8], double Ppred[64]);

/* ------------------------------------------------------------------- T9 */
typedef struct {
    double accel_noise, jolt;
    double pos[3], vel[3], acc[3];
    double P[81];
    int has_imu;
    double imu_a[3], imu_cov[9];
    int variant, n_ignore, best_mode; /* EKF-side NLOS variants (see ko_t6_new_toa_sel); 0 = normal */
    int ml_init; /* 1 = the constructor without initialPosition (TOAIMU.cpp:6-24): see ko_k8.ml_init */
} ko_t9;
void ko_t9_init(ko_t9 *f, double accel_noise, double jolt, const double p0[3]);
void ko_t9_new_toa(ko_t9 *f, double dt, int n_slots, const double *ranges,
                   const double *anchors, const double *errs, ko_info *info);
void ko_t9_new_imu(ko_t9 *f, double dt, const double acc[3], const double cov_acc[9],
                   ko_info *info);
void ko_t9_get_pose(const ko_t9 *f, double dt, double x[9], double Ppred[81]);

/* ------------------------------------------------- batch drivers (OpenMP) */
/* Same SoA layouts as the CUDA C-ABI (include/kfpos_b200.h).  ranges: fmt 0 =
 * f64 metres, 1 = i32 mm, 2 = u16 mm; [T][M][N].  err: scalar if err_arr NULL.
 * x: [3][N] in/out, P: [36][N] in/out (full, row-major index r*6+c).
 * Optional outputs (NULL to skip): traj [T][3][N], sel [T][N], counters[4]
 * (sum ml_iters, cost_evals, gain_evals, status!=0 count), status [N] (OR over steps). */
void ko_t6_replay(int64_t N, int T, int M, const double *anchors, const double *dt,
                  const void *ranges, int fmt, double err_scalar, const double *err_arr,
                  double accel_noise, int ignore_worst, double thr, double *x, double *P,
                  double *traj, int32_t *sel, double *counters, int32_t *status, int threads);

void ko_ml_batch(int64_t N, int M, const double *anchors, const void *ranges, int fmt,
                 double err_scalar, const double *err_arr, const double start[3], int use2d,
                 int variant, int n_ignore, int best_mode, double *pos /*[3][N]*/,
                 double *cov /*[9][N]*/, int32_t *iters /*[N]*/, int32_t *sel /*[2][N]*/,
                 int32_t *status /*[N]*/, int threads);

void ko_t9_replay(int64_t N, int T, int M, const double *anchors, const double *dt,
                  const void *ranges, int fmt, double err_scalar, const double *err_arr,
                  double accel_noise, double jolt, double *x /*[9][N]*/, double *P /*[81][N]*/,
                  double *traj, double *counters, int32_t *status, int threads);

/* sensor-event schedules: same layout as kfpos_event of include/kfpos_b200.h */
typedef struct {
    int32_t kind; /* 0 TOA, 1 PX4, 2 IMU, 3 MAG, 4 COMPASS */
    int32_t _pad;
    double dt;
    int64_t offset;
    double aux[9];
} ko_event;
void ko_k8_replay(int64_t N, int n_events, const ko_event *ev, int M, const double *anchors, const void *ranges,
                  int fmt, double err_scalar, const double *err_arr, const double *sensors, const ko_k8 *cfg,
                  int b1_zero_z, double *x /*[8][N]*/, double *P /*[64][N]*/, double *traj, double *counters /*[5]*/,
                  int32_t *status, int threads, double *tagz /*[N] in/out or NULL*/);
void ko_t9_events_sel(int64_t N, int n_events, const ko_event *ev, int M, const double *anchors, const void *ranges,
                      int fmt, double err_scalar, const double *err_arr, const double *sensors, double accel_noise,
                      double jolt, int variant, int n_ignore, int best_mode, double *x, double *P, double *traj,
                      double *counters, int32_t *status, int threads, int ml_init);
void ko_t9_events(int64_t N, int n_events, const ko_event *ev, int M, const double *anchors, const void *ranges,
                  int fmt, double err_scalar, const double *err_arr, const double *sensors, double accel_noise,
                  double jolt, double *x /*[9][N]*/, double *P /*[81][N]*/, double *traj, double *counters /*[5]*/,
                  int32_t *status, int threads);

int ko_version(void);
int ko_max_threads(void);

void ko_pose_msg(int model, const double *x, const double *P, double tag_z, double pose13[13], double cov36[36]);

/* ------------------------------------------------------- ranging aggregation (ko_assemble.c) */
int64_t ko_assemble(int64_t L, int M, int64_t stride, const uint8_t *anchor, const uint8_t *seq,
                    const int32_t *range_mm, const double *err, const double *t, int64_t max_epochs,
                    int fix_b12, double first_dt, int64_t out_stride, int32_t *ranges_out, double *err_out,
                    double *dt_out, double *t_out);
void ko_assemble_batch(int64_t N, int64_t L, int M, const uint8_t *anchor, const uint8_t *seq,
                       const int32_t *range_mm, const double *err, const double *t, int64_t max_epochs,
                       int fix_b12, double first_dt, int32_t *ranges_out, double *err_out, double *dt_out,
                       int32_t *n_epochs, double *t_out);
void ko_merge_batch(int64_t N, int M, const int64_t L[5], const double *const t_src[5], const int32_t *ranges,
                    const double *err_src, const double *const src[5], int S, const int32_t *slot_kind,
                    const int64_t *slot_row, double first_dt, double *dt_f, int32_t *ranges_out, double *err_out,
                    double *sensors_out, int32_t *n_dropped);

#ifdef __cplusplus
}
#endif
#endif
