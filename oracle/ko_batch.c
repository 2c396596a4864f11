/*
 * ko_batch.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Batch drivers: loop the per-filter oracle over the same structure-of-arrays
 * tensors the CUDA C-ABI takes (include/kfpos_b200.h), with OpenMP over
 * filters.  Used by tests as the checker and by bench.py as the timed CPU
 * baseline ("port").  One filter = one reference object fed at sensor rate
 * (PG.cpp:476-496 -> newTOAMeasurement).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "kfpos_oracle.h"

int ko_version(void) { return 1; }

int ko_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* wire range -> metres: (double) ranges[i] / 1000, PG.cpp:484 */
static inline double load_range(const void *ranges, int fmt, int64_t idx) {
    switch (fmt) {
    case 0: return ((const double *)ranges)[idx];
    case 1: return (double)((const int32_t *)ranges)[idx] / 1000;
    default: return (double)((const uint16_t *)ranges)[idx] / 1000;
    }
}

void ko_t6_replay_sel(int64_t N, int T, int M, const double *anchors, const double *dt,
                      const void *ranges, int fmt, double err_scalar, const double *err_arr,
                      double accel_noise, int ignore_worst, double thr, int variant, int n_ignore, int best_mode,
                      double *x, double *P, double *traj, int32_t *sel, double *counters, int32_t *status,
                      int threads);

void ko_t6_replay(int64_t N, int T, int M, const double *anchors, const double *dt,
                  const void *ranges, int fmt, double err_scalar, const double *err_arr,
                  double accel_noise, int ignore_worst, double thr, double *x, double *P,
                  double *traj, int32_t *sel, double *counters, int32_t *status, int threads) {
    ko_t6_replay_sel(N, T, M, anchors, dt, ranges, fmt, err_scalar, err_arr, accel_noise, ignore_worst, thr, 0, 0, 0,
                     x, P, traj, sel, counters, status, threads);
}

/* variant != 0: the EKF-side NLOS variants (ko_t6_new_toa_sel); sel then holds the slot mask used */
void ko_t6_replay_sel(int64_t N, int T, int M, const double *anchors, const double *dt,
                      const void *ranges, int fmt, double err_scalar, const double *err_arr,
                      double accel_noise, int ignore_worst, double thr, int variant, int n_ignore, int best_mode,
                      double *x, double *P, double *traj, int32_t *sel, double *counters, int32_t *status,
                      int threads) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : c0, c1, c2, c3)
    for (int64_t f = 0; f < N; ++f) {
        ko_t6 flt;
        double p0[3] = {x[0 * N + f], x[1 * N + f], x[2 * N + f]};
        ko_t6_init(&flt, accel_noise, ignore_worst, thr, p0);
        for (int k = 0; k < 36; ++k) flt.P[k] = P[(int64_t)k * N + f];
        int st_or = 0;
        for (int t = 0; t < T; ++t) {
            double r[KO_MAX_ANCHORS], e[KO_MAX_ANCHORS];
            for (int a = 0; a < M; ++a) {
                int64_t idx = ((int64_t)t * M + a) * N + f;
                r[a] = load_range(ranges, fmt, idx);
                e[a] = err_arr ? err_arr[idx] : err_scalar;
            }
            ko_info info;
            uint32_t used = 0;
            if (variant) ko_t6_new_toa_sel(&flt, dt[t], M, r, anchors, e, variant, n_ignore, best_mode, &info, &used);
            else ko_t6_new_toa(&flt, dt[t], M, r, anchors, e, &info);
            c0 += info.ml_iters; c1 += info.cost_evals; c2 += info.gain_evals;
            if (info.status & ~(KO_ST_MAXITER)) c3 += 1;
            st_or |= info.status;
            if (traj)
                for (int k = 0; k < 3; ++k) traj[((int64_t)t * 3 + k) * N + f] = flt.pos[k];
            if (sel) sel[(int64_t)t * N + f] = variant ? (int32_t)used : info.ignored;
        }
        for (int k = 0; k < 3; ++k) x[(int64_t)k * N + f] = flt.pos[k];
        for (int k = 0; k < 36; ++k) P[(int64_t)k * N + f] = flt.P[k];
        if (status) status[f] = st_or;
    }
    if (counters) {
        counters[0] = c0; counters[1] = c1; counters[2] = c2; counters[3] = c3;
    }
}

void ko_t9_replay(int64_t N, int T, int M, const double *anchors, const double *dt,
                  const void *ranges, int fmt, double err_scalar, const double *err_arr,
                  double accel_noise, double jolt, double *x, double *P, double *traj,
                  double *counters, int32_t *status, int threads) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : c0, c1, c2, c3)
    for (int64_t f = 0; f < N; ++f) {
        ko_t9 flt;
        double p0[3] = {x[0 * N + f], x[1 * N + f], x[2 * N + f]};
        ko_t9_init(&flt, accel_noise, jolt, p0);
        for (int k = 0; k < 3; ++k) flt.vel[k] = x[(int64_t)(3 + k) * N + f];
        for (int k = 0; k < 81; ++k) flt.P[k] = P[(int64_t)k * N + f];
        int st_or = 0;
        for (int t = 0; t < T; ++t) {
            double r[KO_MAX_ANCHORS], e[KO_MAX_ANCHORS];
            for (int a = 0; a < M; ++a) {
                int64_t idx = ((int64_t)t * M + a) * N + f;
                r[a] = load_range(ranges, fmt, idx);
                e[a] = err_arr ? err_arr[idx] : err_scalar;
            }
            ko_info info;
            ko_t9_new_toa(&flt, dt[t], M, r, anchors, e, &info);
            c0 += info.ml_iters; c1 += info.cost_evals; c2 += info.gain_evals;
            if (info.status & ~(KO_ST_MAXITER)) c3 += 1;
            st_or |= info.status;
            if (traj)
                for (int k = 0; k < 3; ++k) traj[((int64_t)t * 3 + k) * N + f] = flt.pos[k];
        }
        for (int k = 0; k < 3; ++k) {
            x[(int64_t)k * N + f] = flt.pos[k];
            x[(int64_t)(3 + k) * N + f] = flt.vel[k];
            x[(int64_t)(6 + k) * N + f] = 0.0; /* acceleration never persisted (B-9) */
        }
        for (int k = 0; k < 81; ++k) P[(int64_t)k * N + f] = flt.P[k];
        if (status) status[f] = st_or;
    }
    if (counters) {
        counters[0] = c0; counters[1] = c1; counters[2] = c2; counters[3] = c3;
    }
}

void ko_ml_batch(int64_t N, int M, const double *anchors, const void *ranges, int fmt,
                 double err_scalar, const double *err_arr, const double start[3], int use2d,
                 int variant, int n_ignore, int best_mode, double *pos, double *cov,
                 int32_t *iters, int32_t *sel, int32_t *status, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < N; ++f) {
        double r[KO_MAX_ANCHORS], e[KO_MAX_ANCHORS], p[3], c[9];
        for (int a = 0; a < M; ++a) {
            int64_t idx = (int64_t)a * N + f;
            r[a] = load_range(ranges, fmt, idx);
            e[a] = err_arr ? err_arr[idx] : err_scalar;
        }
        int it = 0;
        int32_t s2[2];
        int rc = ko_ml_epoch(M, r, anchors, e, start, use2d, variant, n_ignore, best_mode, 0, p, c,
                             &it, s2);
        int d = use2d ? 2 : 3;
        for (int k = 0; k < 3; ++k) pos[(int64_t)k * N + f] = p[k];
        /* cov reported as 3x3 with the d x d block in the top-left corner */
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                cov[(int64_t)(a * 3 + b) * N + f] = (a < d && b < d && rc == 0) ? c[a * d + b] : 0.0;
        if (iters) iters[f] = it;
        if (sel) {
            sel[f] = s2[0];
            sel[N + f] = s2[1];
        }
        if (status) status[f] = rc == 0 ? KO_ST_OK : (rc == 1 ? KO_ST_ML_FEW : KO_ST_SINGULAR);
    }
}
