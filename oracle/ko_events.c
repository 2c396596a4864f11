/*
 * ko_events.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Batch drivers for sensor-event schedules (K8, T9): loop the per-filter oracle
 * objects over the same event array and SoA tensors kfpos_batch_replay_events takes
 * (include/kfpos_b200.h), OpenMP over filters.  Event semantics: the callback
 * sequence PosGenerator would issue (PG.cpp:99-141,476-496).
 */
#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "kfpos_oracle.h"

static inline double load_range(const void *ranges, int fmt, int64_t idx) {
    switch (fmt) {
    case 0: return ((const double *)ranges)[idx];
    case 1: return (double)((const int32_t *)ranges)[idx] / 1000;
    default: return (double)((const uint16_t *)ranges)[idx] / 1000;
    }
}

/* cfg: a ko_k8 whose configuration fields are set (state fields ignored).
 * x: [8][N] in/out; P: [64][N] in/out; traj: [n_toa][3][N] (px, py, theta) or NULL.
 * b1_zero_z: see ko_ml2d.  counters[4] as ko_t6_replay. */
void ko_k8_replay(int64_t N, int n_events, const ko_event *ev, int M, const double *anchors, const void *ranges,
                  int fmt, double err_scalar, const double *err_arr, const double *sensors, const ko_k8 *cfg,
                  int b1_zero_z, double *x, double *P, double *traj, double *counters, int32_t *status,
                  int threads, double *tagz /* [N] in/out or NULL: per-filter tag height (cfg->ml_init, 3-D) */) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : c0, c1, c2, c3, c4)
    for (int64_t f = 0; f < N; ++f) {
        ko_k8 flt = *cfg;
        flt.pos[0] = x[0 * N + f]; flt.pos[1] = x[1 * N + f];
        flt.vel[0] = x[2 * N + f]; flt.vel[1] = x[3 * N + f];
        flt.acc[0] = flt.acc[1] = 0.0;
        flt.angle = x[6 * N + f]; flt.omega = x[7 * N + f];
        for (int k = 0; k < 64; ++k) flt.P[k] = P[(int64_t)k * N + f];
        flt.has_mag = flt.has_px4 = flt.has_imu = 0;
        if (tagz) flt.tag_z = tagz[f];
        int st_or = 0, n_toa = 0;
        double carry = 0.0;
        for (int e = 0; e < n_events; ++e) {
            ko_info info;
            memset(&info, 0, sizeof info);
            const double dt = ev[e].dt + carry;
            const int64_t o = ev[e].offset;
            int skipped = 0;
            switch (ev[e].kind) {
            case 0: {
                double r[KO_MAX_ANCHORS], er[KO_MAX_ANCHORS];
                for (int a = 0; a < M; ++a) {
                    int64_t idx = (o + a) * N + f;
                    r[a] = load_range(ranges, fmt, idx);
                    er[a] = err_arr ? err_arr[idx] : err_scalar;
                }
                ko_k8_new_toa(&flt, dt, M, r, anchors, er, b1_zero_z, &info);
                if (traj) {
                    traj[((int64_t)n_toa * 3 + 0) * N + f] = flt.pos[0];
                    traj[((int64_t)n_toa * 3 + 1) * N + f] = flt.pos[1];
                    traj[((int64_t)n_toa * 3 + 2) * N + f] = flt.angle;
                }
                ++n_toa;
                break;
            }
            case 1: {
                const int q = (int)sensors[(o + 4) * N + f];
                if (q == 0) { skipped = 1; break; } /* returns before the clock is read (KF.cpp:111-113) */
                ko_k8_new_px4(&flt, dt, sensors[o * N + f], sensors[(o + 1) * N + f], sensors[(o + 2) * N + f],
                              sensors[(o + 3) * N + f], q, &info);
                break;
            }
            case 2: {
                double w[3] = {0, 0, sensors[o * N + f]};
                double a[3] = {sensors[(o + 1) * N + f], sensors[(o + 2) * N + f], 0};
                double cav[9] = {0}, cac[9] = {0};
                cac[0] = ev[e].aux[0]; cac[1] = ev[e].aux[1]; cac[3] = ev[e].aux[2]; cac[4] = ev[e].aux[3];
                cav[8] = ev[e].aux[4];
                ko_k8_new_imu(&flt, dt, w, cav, a, cac, &info);
                break;
            }
            case 3: {
                double m[3] = {sensors[o * N + f], sensors[(o + 1) * N + f], 0};
                ko_k8_new_mag(&flt, dt, m, &info);
                break;
            }
            default: ko_k8_new_compass(&flt, dt, sensors[o * N + f], &info); break;
            }
            if (skipped) { carry = dt; continue; }
            carry = 0.0;
            c0 += info.ml_iters; c1 += info.cost_evals; c2 += info.gain_evals;
            if (!(info.status & KO_ST_UNINIT)) c4 += 1; /* an update is an event that reaches the predict / update */
            if (info.status & ~(KO_ST_MAXITER | KO_ST_UNINIT)) c3 += 1;
            st_or |= info.status;
        }
        if (tagz) tagz[f] = flt.tag_z;
        x[0 * N + f] = flt.pos[0]; x[1 * N + f] = flt.pos[1];
        x[2 * N + f] = flt.vel[0]; x[3 * N + f] = flt.vel[1];
        x[4 * N + f] = 0.0; x[5 * N + f] = 0.0;
        x[6 * N + f] = flt.angle; x[7 * N + f] = flt.omega;
        for (int k = 0; k < 64; ++k) P[(int64_t)k * N + f] = flt.P[k];
        if (status) status[f] = st_or;
    }
    if (counters) {
        counters[0] = c0; counters[1] = c1; counters[2] = c2; counters[3] = c3; counters[4] = c4;
    }
}

/* T9: events of kind 0 (TOA) and 2 (IMU: rows ax, ay, az; aux = 3x3 covariance).
 * x: [9][N] in/out; P: [81][N] in/out; traj: [n_toa][3][N] or NULL. */
void ko_t9_events(int64_t N, int n_events, const ko_event *ev, int M, const double *anchors, const void *ranges,
                  int fmt, double err_scalar, const double *err_arr, const double *sensors, double accel_noise,
                  double jolt, double *x, double *P, double *traj, double *counters, int32_t *status, int threads) {
    ko_t9_events_sel(N, n_events, ev, M, anchors, ranges, fmt, err_scalar, err_arr, sensors, accel_noise, jolt, 0, 0, 0,
                     x, P, traj, counters, status, threads, 0);
}

/* variant != 0: the EKF-side NLOS variants (ko_t9_new_toa selects the rangings first) */
void ko_t9_events_sel(int64_t N, int n_events, const ko_event *ev, int M, const double *anchors, const void *ranges,
                      int fmt, double err_scalar, const double *err_arr, const double *sensors, double accel_noise,
                      double jolt, int variant, int n_ignore, int best_mode, double *x, double *P, double *traj,
                      double *counters, int32_t *status, int threads, int ml_init) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : c0, c1, c2, c3, c4)
    for (int64_t f = 0; f < N; ++f) {
        ko_t9 flt;
        double p0[3] = {x[0 * N + f], x[1 * N + f], x[2 * N + f]};
        ko_t9_init(&flt, accel_noise, jolt, p0);
        flt.variant = variant; flt.n_ignore = n_ignore; flt.best_mode = best_mode;
        flt.ml_init = ml_init;
        for (int k = 0; k < 3; ++k) flt.vel[k] = x[(int64_t)(3 + k) * N + f];
        for (int k = 0; k < 81; ++k) flt.P[k] = P[(int64_t)k * N + f];
        int st_or = 0, n_toa = 0;
        for (int e = 0; e < n_events; ++e) {
            ko_info info;
            const int64_t o = ev[e].offset;
            if (ev[e].kind == 0) {
                double r[KO_MAX_ANCHORS], er[KO_MAX_ANCHORS];
                for (int a = 0; a < M; ++a) {
                    int64_t idx = (o + a) * N + f;
                    r[a] = load_range(ranges, fmt, idx);
                    er[a] = err_arr ? err_arr[idx] : err_scalar;
                }
                ko_t9_new_toa(&flt, ev[e].dt, M, r, anchors, er, &info);
                if (traj)
                    for (int k = 0; k < 3; ++k) traj[((int64_t)n_toa * 3 + k) * N + f] = flt.pos[k];
                ++n_toa;
            } else {
                double a[3] = {sensors[o * N + f], sensors[(o + 1) * N + f], sensors[(o + 2) * N + f]};
                ko_t9_new_imu(&flt, ev[e].dt, a, ev[e].aux, &info);
            }
            c0 += info.ml_iters; c1 += info.cost_evals; c2 += info.gain_evals;
            if (!(info.status & KO_ST_UNINIT)) c4 += 1;
            if (info.status & ~(KO_ST_MAXITER | KO_ST_UNINIT)) c3 += 1;
            st_or |= info.status;
        }
        for (int k = 0; k < 3; ++k) {
            x[(int64_t)k * N + f] = flt.pos[k];
            x[(int64_t)(3 + k) * N + f] = flt.vel[k];
            x[(int64_t)(6 + k) * N + f] = 0.0;
        }
        for (int k = 0; k < 81; ++k) P[(int64_t)k * N + f] = flt.P[k];
        if (status) status[f] = st_or;
    }
    if (counters) {
        counters[0] = c0; counters[1] = c1; counters[2] = c2; counters[3] = c3; counters[4] = c4;
    }
}
