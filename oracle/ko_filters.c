/*
 * ko_filters.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * As-written (dense, inv/pinv) restatement of the three iterated-EKF classes:
 *   T6  KalmanFilterTOA      src/kfpos/algorithms/KalmanFilterTOA.cpp
 *   K8  KalmanFilter         src/kfpos/algorithms/KalmanFilter.cpp
 *   T9  KalmanFilterTOAIMU   src/kfpos/algorithms/KalmanFilterTOAIMU.cpp
 * dt is an explicit argument (B-8: the reference reads the wall clock).
 * The fixed-initial-position constructors are the ones restated (B-7).
 */
#include <math.h>
#include <string.h>

#include "kfpos_oracle.h"

#define NS 9           /* max state dim */
#define MR KO_MAX_ROWS /* max measurement rows */

static int gather(int n_slots, const double *ranges, const double *anchors, const double *errs,
                  ko_meas *m) {
    /* newTOAMeasurement: keep rangings[i] > 0 in arrival order
     * (TOA.cpp:48-57, KF.cpp:69-78, TOAIMU.cpp:54-63) */
    int n = 0;
    for (int i = 0; i < n_slots; ++i)
        if (ranges[i] > 0) {
            m[n].r = ranges[i];
            m[n].e = errs[i];
            m[n].bx = anchors[3 * i];
            m[n].by = anchors[3 * i + 1];
            m[n].bz = anchors[3 * i + 2];
            m[n].slot = i;
            ++n;
        }
    return n;
}

/* P <- F P F^T + Q, dense as written (TOA.cpp:122, KF.cpp:302, TOAIMU.cpp:179) */
static void predict_cov(int n, const double *F, const double *Q, double *P) {
    double T[NS * NS];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += F[i * n + k] * P[k * n + j];
            T[i * n + j] = s;
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += T[i * n + k] * F[j * n + k];
            P[i * n + j] = s + Q[i * n + j];
        }
}

static void matvec(int n, const double *F, const double *x, double *y) {
    for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int k = 0; k < n; ++k) s += F[i * n + k] * x[k];
        y[i] = s;
    }
}

/* KalmanFilter::normalizeAngle, KF.cpp:699-706 (single wrap) */
static double wrap_angle(double a) {
    if (a > M_PI) return a - 2 * M_PI;
    else if (a <= -M_PI) return a + 2 * M_PI;
    return a;
}

/* --------------------------------------------------------------------------
 * The iterated update shared by the three classes (TOA.cpp:285-326,
 * KF.cpp:444-499, TOAIMU.cpp:296-338).  `model` evaluates the sensor outputs
 * h(x) and the Jacobian at x.  Returns -1 if an inv() hit a singular matrix
 * (arma::inv would throw std::runtime_error).
 * -------------------------------------------------------------------------- */
typedef void (*ko_model_fn)(const void *ctx, const double *x, double *h, double *J);

static int iekf_dense(int n, int M, const double *xpred, const double *Ppred, const double *z,
                      const double *R, ko_model_fn model, const void *ctx, int max_steps,
                      double min_rel, int mag_row, double *x_out, double *P_out, ko_info *info) {
    double Rinv[MR * MR], Pinv[NS * NS];
    double J[MR * NS], Jx[MR * NS], K[NS * MR], h[MR], eps[MR], delta[NS];
    double x[NS];
    int have_gain = 0;
    memset(J, 0, sizeof J);
    memset(K, 0, sizeof K);
    if (ko_inv(M, R, Rinv) != 0) return -1;
    ko_pinv(n, Ppred, Pinv);
    memcpy(x, xpred, sizeof(double) * n);
    double cost = 1e20;
    int broke = 0;
    for (int iter = 0; iter < max_steps; ++iter) {
        model(ctx, x, h, Jx); /* Jx is only adopted after the break test below */
        for (int i = 0; i < M; ++i) eps[i] = z[i] - h[i];
        if (mag_row >= 0) eps[mag_row] = wrap_angle(eps[mag_row]); /* KF.cpp:461-463 */
        for (int i = 0; i < n; ++i) delta[i] = xpred[i] - x[i];
        double c = 0;
        for (int i = 0; i < M; ++i) {
            double s = 0;
            for (int j = 0; j < M; ++j) s += Rinv[i * M + j] * eps[j];
            c += eps[i] * s;
        }
        double c2 = 0;
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int j = 0; j < n; ++j) s += Pinv[i * n + j] * delta[j];
            c2 += delta[i] * s;
        }
        double newCost = c + c2;
        info->cost_evals++;
        if (fabs(cost - newCost) / cost < min_rel) { broke = 1; break; }
        cost = newCost;
        memcpy(J, Jx, sizeof(double) * M * n); /* jacobian*(...) calls come after the break */
        /* K = P J^T inv(J P J^T + R) */
        double PJt[NS * MR], S[MR * MR], Sinv[MR * MR];
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < M; ++r) {
                double s = 0;
                for (int k = 0; k < n; ++k) s += Ppred[i * n + k] * J[r * n + k];
                PJt[i * M + r] = s;
            }
        for (int r = 0; r < M; ++r)
            for (int q = 0; q < M; ++q) {
                double s = 0;
                for (int k = 0; k < n; ++k) s += J[r * n + k] * PJt[k * M + q];
                S[r * M + q] = s + R[r * M + q];
            }
        if (ko_inv(M, S, Sinv) != 0) return -1;
        for (int i = 0; i < n; ++i)
            for (int q = 0; q < M; ++q) {
                double s = 0;
                for (int r = 0; r < M; ++r) s += PJt[i * M + r] * Sinv[r * M + q];
                K[i * M + q] = s;
            }
        have_gain = 1;
        info->gain_evals++;
        /* direction = delta + K (eps - J delta) */
        double y[MR];
        for (int r = 0; r < M; ++r) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += J[r * n + k] * delta[k];
            y[r] = eps[r] - s;
        }
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int r = 0; r < M; ++r) s += K[i * M + r] * y[r];
            x[i] = x[i] + (delta[i] + s);
        }
    }
    if (!broke) info->status |= KO_ST_MAXITER;
    /* P = (I - K J) P with the last K, J (TOA.cpp:326, KF.cpp:499, TOAIMU.cpp:338) */
    if (!have_gain) {
        /* kalmanGain is an empty arma::mat: the product would throw; cannot
         * happen because the first test is against cost = 1e20. */
        memcpy(P_out, Ppred, sizeof(double) * n * n);
    } else {
        double IKJ[NS * NS];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int r = 0; r < M; ++r) s += K[i * M + r] * J[r * n + j];
                IKJ[i * n + j] = (i == j ? 1.0 : 0.0) - s;
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int k = 0; k < n; ++k) s += IKJ[i * n + k] * Ppred[k * n + j];
                P_out[i * n + j] = s;
            }
    }
    memcpy(x_out, x, sizeof(double) * n);
    info->cost = cost;
    return 0;
}

/* ========================================================================== T6 */
typedef struct {
    const ko_meas *m;
    int n_meas;
} t6_ctx;

/* sensorOutputs TOA.cpp:341-358 + jacobianRangings TOA.cpp:421-433 */
static void t6_model(const void *vctx, const double *x, double *h, double *J) {
    const t6_ctx *c = (const t6_ctx *)vctx;
    ko_dist(c->m, c->n_meas, x, h);
    for (int i = 0; i < c->n_meas; ++i) {
        J[i * 6 + 0] = (x[0] - c->m[i].bx) / h[i];
        J[i * 6 + 1] = (x[1] - c->m[i].by) / h[i];
        J[i * 6 + 2] = (x[2] - c->m[i].bz) / h[i];
        J[i * 6 + 3] = J[i * 6 + 4] = J[i * 6 + 5] = 0;
    }
}

void ko_t6_init(ko_t6 *f, double accel_noise, int ignore_worst, double thr, const double p0[3]) {
    memset(f, 0, sizeof *f); /* estimationCovariance.zeros(6,6), mVelocity = 0: TOA.cpp:29-32 */
    f->accel_noise = accel_noise;
    f->ignore_worst = ignore_worst;
    f->ignore_cost_threshold = thr;
    memcpy(f->pos, p0, sizeof f->pos);
}

/* predictionMatrix TOA.cpp:362-369, predictionErrorCovariance TOA.cpp:371-391 */
static void t6_FQ(double a, double t, double *F, double *Q) {
    memset(F, 0, sizeof(double) * 36);
    memset(Q, 0, sizeof(double) * 36);
    for (int i = 0; i < 6; ++i) F[i * 6 + i] = 1;
    for (int i = 0; i < 3; ++i) F[i * 6 + i + 3] = t;
    double t2 = pow(t, 2) / 2, a2 = a * a;
    for (int i = 0; i < 3; ++i) {
        Q[i * 6 + i] = a2 * t2 * t2;
        Q[i * 6 + i + 3] = a2 * t2 * t;
        Q[(i + 3) * 6 + i] = a2 * t2 * t;
        Q[(i + 3) * 6 + i + 3] = a2 * t * t;
    }
}

/* kalmanStep3DIgnoreAnchor, TOA.cpp:242-338 */
static int t6_step_ignore(const double *xpred, const double *Ppred, const ko_meas *all, int n_all,
                          int ignored, double *x_out, double *P_out, ko_info *info) {
    ko_meas m[KO_MAX_ANCHORS];
    int n = 0;
    for (int i = 0; i < n_all; ++i)
        if (i != ignored) m[n++] = all[i];
    double mlp[3], mlcov[9];
    int it = 0;
    int rc = ko_ml3d(m, n, xpred, mlp, mlcov, &it);
    info->ml_iters += it;
    if (rc == 1) info->status |= KO_ST_ML_FEW;
    if (rc < 0) return -1; /* arma::solve / inv threw */
    if (isnan(mlp[0]) || isnan(mlp[1]) || isnan(mlp[2])) { /* TOA.cpp:270-272 */
        mlp[0] = xpred[0]; mlp[1] = xpred[1]; mlp[2] = xpred[2];
        info->status |= KO_ST_ML_NAN;
    }
    double sse = ko_sse(m, n, mlp);
    double R[MR * MR], z[MR];
    memset(R, 0, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) {
        z[i] = m[i].r;
        R[i * n + i] = fmax(sse, m[i].e); /* std::max(mlRangingError, errorEstimation) */
    }
    t6_ctx ctx = {m, n};
    return iekf_dense(6, n, xpred, Ppred, z, R, t6_model, &ctx, 10, 1e-3, -1, x_out, P_out, info);
}

/* estimatePositionKF, TOA.cpp:70-156 (fixed-initial-position branch) */
static void t6_estimate(ko_t6 *f, double dt, const ko_meas *m, int n, ko_info *info) {
    double F[36], Q[36], x[6], xp[6];
    memset(info, 0, sizeof *info);
    info->ignored = -1;
    x[0] = f->pos[0]; x[1] = f->pos[1]; x[2] = f->pos[2];
    x[3] = f->vel[0]; x[4] = f->vel[1]; x[5] = f->vel[2];
    t6_FQ(f->accel_noise, dt, F, Q);
    matvec(6, F, x, xp);
    predict_cov(6, F, Q, f->P); /* member overwritten before the try block: TOA.cpp:122 */
    if (n == 0) info->status |= KO_ST_NO_MEAS;
    double xs[6], Ps[36];
    int rc;
    if (n > 4 && f->ignore_worst) {
        /* kalmanStep3DCanIgnoreAnAnchor, TOA.cpp:185-238.  Counters and status
         * accumulate over the 1 + n solves. */
        rc = t6_step_ignore(xp, f->P, m, n, -1, xs, Ps, info);
        double cost_all = info->cost;
        double maxDist = 0, worstCost = 0, xb[6], Pb[36];
        int idx = -1;
        for (int i = 0; i < n && rc == 0; ++i) {
            double xi[6], Pi[36];
            rc = t6_step_ignore(xp, f->P, m, n, i, xi, Pi, info);
            if (rc != 0) break;
            double dd = sqrt(pow(m[i].bx - xi[0], 2) + pow(m[i].by - xi[1], 2) +
                             pow(m[i].bz - xi[2], 2));
            double diff = m[i].r - dd;
            if (i == 0 || diff > maxDist) { /* strict >, i = 0 seeds: TOA.cpp:209 */
                maxDist = diff;
                worstCost = info->cost;
                memcpy(xb, xi, sizeof xb);
                memcpy(Pb, Pi, sizeof Pb);
                idx = i;
            }
        }
        info->cost = cost_all;
        if (rc == 0 && maxDist > 0 && (cost_all - worstCost) > f->ignore_cost_threshold) {
            memcpy(xs, xb, sizeof xs);
            memcpy(Ps, Pb, sizeof Ps);
            info->ignored = m[idx].slot; /* reported as the anchor SLOT */
            info->cost = worstCost;
        }
    } else {
        rc = t6_step_ignore(xp, f->P, m, n, -1, xs, Ps, info);
    }
    if (rc != 0) { /* catch (std::runtime_error): update skipped, P stays predicted */
        info->status |= KO_ST_SINGULAR;
        return;
    }
    memcpy(f->P, Ps, sizeof f->P);
    f->pos[0] = xs[0]; f->pos[1] = xs[1]; f->pos[2] = xs[2]; /* stateToPose: velocity dropped */
    if (!(isfinite(xs[0]) && isfinite(xs[1]) && isfinite(xs[2]))) info->status |= KO_ST_NAN;
}

void ko_t6_new_toa(ko_t6 *f, double dt, int n_slots, const double *ranges, const double *anchors,
                   const double *errs, ko_info *info) {
    ko_meas m[KO_MAX_ANCHORS];
    int n = gather(n_slots, ranges, anchors, errs, m);
    t6_estimate(f, dt, m, n, info);
}

/* EKF-side NLOS variants (README.md:85-108, config_pos.xml:5-28).  The reference documents
 * variant / numIgnoredRangings / bestMode for "type = 1: EKF" but implements the selection only in
 * MLLocation, so there is no reference behaviour to pin: this is the restatement of north_star (3).
 * The ML estimator is started at the predicted position (= the stored position: the predicted
 * velocity is 0) and selects the rangings; the normal update then runs on the survivors, kept in
 * slot order.  variant 1: solve with all rangings, sort by squared residual at that solution,
 * drop the last min(n - 4, n_ignore) (ML.cpp:307-347; no selection when that solve fails);
 * variant 2: the 4-ranging group of ko_ml_best_group (ML.cpp:351-414) when n >= 4.
 * The Newton iterations of the selection solves are added to info->ml_iters.  *mask_out = slots used. */
void ko_t6_new_toa_sel(ko_t6 *f, double dt, int n_slots, const double *ranges, const double *anchors,
                       const double *errs, int variant, int n_ignore, int best_mode, ko_info *info,
                       uint32_t *mask_out) {
    ko_meas m[KO_MAX_ANCHORS], sub[KO_MAX_ANCHORS];
    int n = gather(n_slots, ranges, anchors, errs, m);
    unsigned char keep[KO_MAX_ANCHORS];
    for (int i = 0; i < n; ++i) keep[i] = 1;
    int it_sel = 0, it = 0;
    const double start[3] = {f->pos[0], f->pos[1], f->pos[2]};
    if (variant == 1 && n > 0) {
        double p0[3], c0[9];
        if (ko_ml3d(m, n, start, p0, c0, &it) == 0) {
            int order[KO_MAX_ANCHORS];
            ko_best_rangings(m, n, p0, order);
            int drop = n - 4 < n_ignore ? n - 4 : n_ignore;
            if (drop < 0) drop = 0;
            for (int i = n - drop; i < n; ++i) keep[order[i]] = 0;
        }
        it_sel += it;
    } else if (variant == 2 && n >= 4) {
        double p0[3], c0[9];
        int bi, ng;
        uint32_t bm = 0;
        ko_ml_best_group(m, n, start, 0, best_mode, 0, p0, c0, &it, &bi, &bm, &ng);
        it_sel += it;
        if (bi >= 0) /* a failed solve selects nothing: the update then runs on all rangings */
            for (int i = 0; i < n; ++i) keep[i] = (bm >> i) & 1u;
    }
    int ns = 0;
    uint32_t mask = 0;
    for (int i = 0; i < n; ++i)
        if (keep[i]) {
            sub[ns++] = m[i];
            mask |= 1u << m[i].slot;
        }
    t6_estimate(f, dt, sub, ns, info);
    info->ml_iters += it_sel;
    if (mask_out) *mask_out = mask;
}

/* getPose, TOA.cpp:438-473: predict only, state untouched */
void ko_t6_get_pose(const ko_t6 *f, double dt, double pos[3], double Ppred[36]) {
    double F[36], Q[36], x[6], xp[6];
    x[0] = f->pos[0]; x[1] = f->pos[1]; x[2] = f->pos[2];
    x[3] = f->vel[0]; x[4] = f->vel[1]; x[5] = f->vel[2];
    t6_FQ(f->accel_noise, dt, F, Q);
    matvec(6, F, x, xp);
    memcpy(Ppred, f->P, sizeof(double) * 36);
    predict_cov(6, F, Q, Ppred);
    pos[0] = xp[0]; pos[1] = xp[1]; pos[2] = xp[2];
}

/* ========================================================================== K8 */
typedef struct {
    const ko_k8 *f;
    const ko_meas *m;
    int n_meas; /* 0 when !hasRangingMeasurements */
    int has_px4, has_imu, has_mag;
    double dt;
} k8_ctx;

void ko_k8_init(ko_k8 *f, double accel_noise, double init_angle, double jolt, const double p0[2]) {
    memset(f, 0, sizeof *f);
    f->accel_noise = accel_noise;
    f->jolt = jolt;
    f->angle = init_angle;
    f->pos[0] = p0[0];
    f->pos[1] = p0[1];
}

/* predictionMatrix KF.cpp:583-592, predictionErrorCovariance KF.cpp:594-609 */
static void k8_FQ(double a, double j, double t, double *F, double *Q) {
    memset(F, 0, sizeof(double) * 64);
    memset(Q, 0, sizeof(double) * 64);
    for (int i = 0; i < 8; ++i) F[i * 8 + i] = 1;
    F[0 * 8 + 2] = t; F[0 * 8 + 4] = t * t / 2;
    F[1 * 8 + 3] = t; F[1 * 8 + 5] = t * t / 2;
    F[2 * 8 + 4] = t; F[3 * 8 + 5] = t;
    F[6 * 8 + 7] = t;
    double t3 = pow(t, 3) / 6, t2 = pow(t, 2) / 2;
    double u[3] = {t3, t2, t};
    for (int ax = 0; ax < 2; ++ax)
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) Q[(ax + 2 * p) * 8 + (ax + 2 * q)] = j * u[p < q ? p : q] * u[p < q ? q : p];
    Q[6 * 8 + 6] = a * t2 * t2;
    Q[6 * 8 + 7] = a * t2 * t;
    Q[7 * 8 + 6] = a * t2 * t;
    Q[7 * 8 + 7] = a * t * t;
}

/* sensorOutputs KF.cpp:506-561, px4flowOutput :563-571, imuOutput :573-581,
 * Jacobians :611-697.  Row layout [ranges | px4(3) | imu(3) | mag(1)]. */
static void k8_model(const void *vctx, const double *x, double *h, double *J) {
    const k8_ctx *c = (const k8_ctx *)vctx;
    const ko_k8 *f = c->f;
    int row = 0;
    double pos[3] = {x[0], x[1], f->tag_z};
    double vx = x[2], vy = x[3], ax = x[4], ay = x[5], ang = x[6], w = x[7], t = c->dt;
    if (c->n_meas > 0) {
        ko_dist(c->m, c->n_meas, pos, h);
        for (int i = 0; i < c->n_meas; ++i) {
            for (int k = 0; k < 8; ++k) J[i * 8 + k] = 0;
            J[i * 8 + 0] = (pos[0] - c->m[i].bx) / h[i];
            J[i * 8 + 1] = (pos[1] - c->m[i].by) / h[i];
        }
        row = c->n_meas;
    }
    if (c->has_px4) {
        double A1 = f->px4_arm1, A2 = f->px4_arm2;
        h[row] = cos(ang) * vx + sin(ang) * vy + 1 / t * ((1 - cos(w * t)) * A1 - sin(w * t) * A2);
        h[row + 1] = -sin(ang) * vx + cos(ang) * vy + 1 / t * (sin(w * t) * A1 + (1 - cos(w * t)) * A2);
        h[row + 2] = w;
        for (int k = 0; k < 24; ++k) J[row * 8 + k] = 0;
        J[row * 8 + 2] = cos(ang);
        J[row * 8 + 3] = sin(ang);
        J[row * 8 + 6] = -sin(ang) * vx + cos(ang) * vy;
        J[row * 8 + 7] = A1 * sin(w * t) - A2 * cos(w * t);
        J[(row + 1) * 8 + 2] = -sin(ang);
        J[(row + 1) * 8 + 3] = cos(ang);
        J[(row + 1) * 8 + 6] = -cos(ang) * vx - sin(ang) * vy;
        J[(row + 1) * 8 + 7] = A1 * cos(w * t) + A2 * sin(w * t);
        J[(row + 2) * 8 + 7] = 1;
        row += 3;
    }
    if (c->has_imu) {
        h[row] = cos(ang) * ax + sin(ang) * ay;
        h[row + 1] = -sin(ang) * ax + cos(ang) * ay;
        h[row + 2] = w;
        for (int k = 0; k < 24; ++k) J[row * 8 + k] = 0;
        J[row * 8 + 4] = cos(ang);
        J[row * 8 + 5] = sin(ang);
        J[row * 8 + 6] = -sin(ang) * ax + cos(ang) * ay;
        J[(row + 1) * 8 + 4] = -sin(ang);
        J[(row + 1) * 8 + 5] = cos(ang);
        J[(row + 1) * 8 + 6] = -cos(ang) * ax - sin(ang) * ay;
        J[(row + 2) * 8 + 7] = 1;
        row += 3;
    }
    if (c->has_mag) {
        h[row] = ang;
        for (int k = 0; k < 8; ++k) J[row * 8 + k] = 0;
        J[row * 8 + 6] = 1;
    }
}

/* estimatePositionKF KF.cpp:224-321 + kalmanStep3D KF.cpp:365-501 */
static void k8_estimate(ko_k8 *f, double dt, int has_r, const ko_meas *m, int n, int has_px4,
                        int has_imu, int has_mag, int b1_zero_z, ko_info *info) {
    double F[64], Q[64], x[8], xp[8];
    memset(info, 0, sizeof *info);
    info->ignored = -1;
    if (f->ml_init && (isnan(f->pos[0]) || isnan(f->pos[1]))) {
        /* KF.cpp:244-285: the first position comes from ML, started at (1, 1, tag height) in 2-D or (1, 1, 4)
         * in 3-D (then the tag height becomes the estimated z); the 2x2 position block of the ML covariance
         * goes into the all-zero estimation covariance; nothing else happens in this call. */
        info->status |= KO_ST_UNINIT;
        if (has_r) {
            double p[3], c[9];
            int it = 0, rc, d;
            if (f->use_fixed_height) {
                const double start[3] = {1.0, 1.0, f->tag_z};
                rc = ko_ml2d(m, n, start, b1_zero_z, p, c, &it);
                d = 2;
            } else {
                const double start[3] = {1.0, 1.0, 4.0};
                rc = ko_ml3d(m, n, start, p, c, &it);
                d = 3;
            }
            info->ml_iters += it;
            if (rc < 0) { info->status |= KO_ST_SINGULAR; return; } /* throws before mPosition is assigned */
            f->pos[0] = p[0]; f->pos[1] = p[1];
            if (!f->use_fixed_height) f->tag_z = p[2];
            /* too few rangings: the solver returns its start point with an EMPTY covariance matrix; the
             * position (and tag height) are assigned, then covarianceMatrix(0,0) throws (KF.cpp:267) */
            if (rc == 1) { info->status |= KO_ST_ML_FEW; return; }
            f->P[0 * 8 + 0] = c[0 * d + 0]; f->P[1 * 8 + 0] = c[1 * d + 0];
            f->P[0 * 8 + 1] = c[0 * d + 1]; f->P[1 * 8 + 1] = c[1 * d + 1];
        }
        return;
    }
    x[0] = f->pos[0]; x[1] = f->pos[1];
    x[2] = f->vel[0]; x[3] = f->vel[1];
    x[4] = f->acc[0]; x[5] = f->acc[1]; /* always 0: never written back (B-9) */
    x[6] = f->angle; x[7] = f->omega;
    k8_FQ(f->accel_noise, f->jolt, dt, F, Q);
    matvec(8, F, x, xp);
    predict_cov(8, F, Q, f->P);
    xp[6] = wrap_angle(xp[6]); /* KF.cpp:305 */

    int nr = has_r ? n : 0;
    int M = nr, ipx4 = 0, iimu = 0, imag = -1;
    if (has_px4) { ipx4 = M; M += 3; }
    if (has_imu) { iimu = M; M += 3; }
    if (has_mag) { imag = M; M += 1; }
    double R[MR * MR], z[MR];
    memset(R, 0, sizeof(double) * M * M);
    for (int i = 0; i < M; ++i) R[i * M + i] = 1; /* arma::eye */
    if (has_r) { /* KF.cpp:403-411 */
        double start[3] = {xp[0], xp[1], f->tag_z}, mlp[3], mlcov[4];
        int it = 0;
        int rc = ko_ml2d(m, n, start, b1_zero_z, mlp, mlcov, &it);
        info->ml_iters += it;
        if (rc == 1) info->status |= KO_ST_ML_FEW;
        if (rc < 0) { info->status |= KO_ST_SINGULAR; return; } /* uncaught throw in the reference */
        double sse = ko_sse(m, n, mlp);
        for (int i = 0; i < n; ++i) {
            z[i] = m[i].r;
            R[i * M + i] = fmax(sse, m[i].e);
        }
        if (n == 0) info->status |= KO_ST_NO_MEAS;
    }
    if (has_px4) { /* KF.cpp:413-424 */
        z[ipx4] = f->px4_vx; z[ipx4 + 1] = f->px4_vy; z[ipx4 + 2] = f->px4_gz;
        R[ipx4 * M + ipx4] = f->px4_cv;
        R[(ipx4 + 1) * M + ipx4 + 1] = f->px4_cv;
        R[(ipx4 + 2) * M + ipx4 + 2] = f->px4_cg;
    }
    if (has_imu) { /* KF.cpp:426-436 */
        z[iimu] = f->imu_ax; z[iimu + 1] = f->imu_ay; z[iimu + 2] = f->imu_wz;
        R[iimu * M + iimu] = f->imu_cxy[0];
        R[iimu * M + iimu + 1] = f->imu_cxy[1];
        R[(iimu + 1) * M + iimu] = f->imu_cxy[2];
        R[(iimu + 1) * M + iimu + 1] = f->imu_cxy[3];
        R[(iimu + 2) * M + iimu + 2] = f->imu_cwz;
    }
    if (has_mag) { /* KF.cpp:438-442 */
        z[imag] = f->mag_angle;
        R[imag * M + imag] = f->mag_c;
    }
    k8_ctx ctx = {f, m, nr, has_px4, has_imu, has_mag, dt};
    double xs[8], Ps[64];
    int rc = iekf_dense(8, M, xp, f->P, z, R, k8_model, &ctx, 20, 1e-4, imag, xs, Ps, info);
    if (rc != 0) { info->status |= KO_ST_SINGULAR; return; }
    memcpy(f->P, Ps, sizeof f->P);
    f->vel[0] = xs[2]; f->vel[1] = xs[3]; /* KF.cpp:315-318 */
    f->angle = xs[6];
    f->omega = xs[7];
    f->pos[0] = xs[0]; f->pos[1] = xs[1]; /* stateToPose KF.cpp:326-327 */
    for (int i = 0; i < 8; ++i)
        if (!isfinite(xs[i])) info->status |= KO_ST_NAN;
}

/* newTOAMeasurement KF.cpp:64-97: rangings + latched px4/imu/mag */
void ko_k8_new_toa(ko_k8 *f, double dt, int n_slots, const double *ranges, const double *anchors,
                   const double *errs, int b1_zero_z, ko_info *info) {
    ko_meas m[KO_MAX_ANCHORS];
    int n = gather(n_slots, ranges, anchors, errs, m);
    if (f->variant == 0 || (f->ml_init && (isnan(f->pos[0]) || isnan(f->pos[1])))) {
        k8_estimate(f, dt, 1, m, n, f->has_px4, f->has_imu, f->has_mag, b1_zero_z, info);
        return;
    }
    /* EKF-side NLOS variants, the 2-D counterpart of ko_t6_new_toa_sel: the ML estimator, started at the
     * predicted position (p + dt v, tag height), selects the rangings -- variant 1 drops the
     * min(n - 3, n_ignore) with the largest residual, variant 2 keeps the best 3-anchor group
     * (ML.cpp:351-414 with the 2-D criterion of App. B-4) -- and the update runs on the survivors. */
    ko_meas sub[KO_MAX_ANCHORS];
    unsigned char keep[KO_MAX_ANCHORS];
    for (int i = 0; i < n; ++i) keep[i] = 1;
    int it_sel = 0, it = 0;
    const double start[3] = {f->pos[0] + dt * f->vel[0], f->pos[1] + dt * f->vel[1], f->tag_z};
    if (f->variant == 1 && n > 0) {
        double p0[3], c0[4];
        if (ko_ml2d(m, n, start, b1_zero_z, p0, c0, &it) == 0) {
            int order[KO_MAX_ANCHORS];
            ko_best_rangings(m, n, p0, order);
            int drop = n - 3 < f->n_ignore ? n - 3 : f->n_ignore;
            if (drop < 0) drop = 0;
            for (int i = n - drop; i < n; ++i) keep[order[i]] = 0;
        }
        it_sel += it;
    } else if (f->variant == 2 && n >= 3) {
        double p0[3], c0[9];
        int bi, ng;
        uint32_t bm = 0;
        ko_ml_best_group(m, n, start, 1, f->best_mode, b1_zero_z, p0, c0, &it, &bi, &bm, &ng);
        it_sel += it;
        if (bi >= 0) /* a failed solve selects nothing: the update then runs on all rangings */
            for (int i = 0; i < n; ++i) keep[i] = (bm >> i) & 1u;
    }
    int ns = 0;
    for (int i = 0; i < n; ++i)
        if (keep[i]) sub[ns++] = m[i];
    k8_estimate(f, dt, 1, sub, ns, f->has_px4, f->has_imu, f->has_mag, b1_zero_z, info);
    info->ml_iters += it_sel;
}

/* newPX4FlowMeasurement KF.cpp:100-133 */
void ko_k8_new_px4(ko_k8 *f, double dt, double ix, double iy, double irz, double itime_us,
                   int quality, ko_info *info) {
    memset(info, 0, sizeof *info);
    info->ignored = -1;
    double vy = iy / (itime_us / 1000000.0) * f->px4_height;
    double vx = ix / (itime_us / 1000000.0) * f->px4_height;
    double gz = irz / (itime_us / 1000000.0);
    double it = itime_us / 1000000.0;
    if (quality == 0) { info->status |= KO_ST_NO_MEAS; return; }
    double cv;
    if (itime_us > 0) cv = f->px4_cov_vel / it * f->px4_height / quality;
    else cv = f->px4_cov_vel * quality;
    f->px4_itime = it; f->px4_vx = vx; f->px4_vy = vy; f->px4_gz = gz;
    f->px4_cv = cv; f->px4_cg = f->px4_cov_gyro;
    f->has_px4 = 1;
    k8_estimate(f, dt, 0, NULL, 0, 1, 0, 0, 0, info);
}

/* newIMUMeasurement KF.cpp:137-176 */
void ko_k8_new_imu(ko_k8 *f, double dt, const double angvel[3], const double cov_av[9],
                   const double acc[3], const double cov_acc[9], ko_info *info) {
    if (f->imu_fixed_cov_acc) {
        f->imu_cxy[0] = f->imu_cov_acc; f->imu_cxy[1] = cov_acc[1];
        f->imu_cxy[2] = cov_acc[3];     f->imu_cxy[3] = f->imu_cov_acc;
    } else {
        f->imu_cxy[0] = cov_acc[0]; f->imu_cxy[1] = cov_acc[1];
        f->imu_cxy[2] = cov_acc[3]; f->imu_cxy[3] = cov_acc[4];
    }
    f->imu_cwz = f->imu_fixed_cov_gyro ? f->imu_cov_gyro : cov_av[8];
    f->imu_wz = angvel[2];
    f->imu_ax = acc[0];
    f->imu_ay = acc[1];
    f->has_imu = 1;
    k8_estimate(f, dt, 0, NULL, 0, 0, 1, 0, 0, info);
}

/* newMAGMeasurement KF.cpp:179-193: mag only, angle NOT normalised */
void ko_k8_new_mag(ko_k8 *f, double dt, const double mag[3], ko_info *info) {
    f->mag_angle = atan2(mag[1], mag[0]) - f->mag_offset;
    f->mag_c = f->mag_cov;
    f->has_mag = 1;
    k8_estimate(f, dt, 0, NULL, 0, 0, 0, 1, 0, info);
}

/* newCompassMeasurement KF.cpp:195-221: mag + latched px4 + latched imu */
void ko_k8_new_compass(ko_k8 *f, double dt, double compass, ko_info *info) {
    f->mag_angle = wrap_angle(compass);
    f->mag_c = f->mag_cov;
    f->has_mag = 1;
    k8_estimate(f, dt, 0, NULL, 0, f->has_px4, f->has_imu, 1, 0, info);
}

/* getPose KF.cpp:709-747 */
void ko_k8_get_pose(const ko_k8 *f, double dt, double xo[8], double Ppred[64]) {
    double F[64], Q[64], x[8];
    x[0] = f->pos[0]; x[1] = f->pos[1]; x[2] = f->vel[0]; x[3] = f->vel[1];
    x[4] = f->acc[0]; x[5] = f->acc[1]; x[6] = f->angle; x[7] = f->omega;
    k8_FQ(f->accel_noise, f->jolt, dt, F, Q);
    matvec(8, F, x, xo);
    memcpy(Ppred, f->P, sizeof(double) * 64);
    predict_cov(8, F, Q, Ppred);
    xo[6] = wrap_angle(xo[6]);
}

/* ========================================================================== T9 */
typedef struct {
    const ko_meas *m;
    int n_meas;
    int has_imu;
} t9_ctx;

void ko_t9_init(ko_t9 *f, double accel_noise, double jolt, const double p0[3]) {
    memset(f, 0, sizeof *f);
    f->accel_noise = accel_noise;
    f->jolt = jolt;
    memcpy(f->pos, p0, sizeof f->pos);
}

/* predictionMatrix TOAIMU.cpp:392-402, predictionErrorCovariance :405-421 */
static void t9_FQ(double j, double t, double *F, double *Q) {
    memset(F, 0, sizeof(double) * 81);
    memset(Q, 0, sizeof(double) * 81);
    for (int i = 0; i < 9; ++i) F[i * 9 + i] = 1;
    for (int i = 0; i < 3; ++i) {
        F[i * 9 + i + 3] = t;
        F[i * 9 + i + 6] = t * t / 2;
        F[(i + 3) * 9 + i + 6] = t;
    }
    double t3 = pow(t, 3) / 6, t2 = pow(t, 2) / 2;
    double u[3] = {t3, t2, t};
    for (int ax = 0; ax < 3; ++ax)
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) Q[(ax + 3 * p) * 9 + (ax + 3 * q)] = j * u[p < q ? p : q] * u[p < q ? q : p];
}

/* sensorOutputs TOAIMU.cpp:345-388, jacobianRangings :423-438.
 * B-5: the reference IMU block is defective (9 rows reserved for 3 values,
 * jacobian(row,9) out of range, d h/d a written as diag(a)).  Restated as 3 IMU
 * rows with h = a and J = I3 on the acceleration block. */
static void t9_model(const void *vctx, const double *x, double *h, double *J) {
    const t9_ctx *c = (const t9_ctx *)vctx;
    int row = 0;
    if (c->n_meas > 0) {
        ko_dist(c->m, c->n_meas, x, h);
        for (int i = 0; i < c->n_meas; ++i) {
            for (int k = 0; k < 9; ++k) J[i * 9 + k] = 0;
            J[i * 9 + 0] = (x[0] - c->m[i].bx) / h[i];
            J[i * 9 + 1] = (x[1] - c->m[i].by) / h[i];
            J[i * 9 + 2] = (x[2] - c->m[i].bz) / h[i];
        }
        row = c->n_meas;
    }
    if (c->has_imu) {
        for (int k = 0; k < 27; ++k) J[row * 9 + k] = 0;
        for (int a = 0; a < 3; ++a) {
            h[row + a] = x[6 + a];
            J[(row + a) * 9 + 6 + a] = 1;
        }
    }
}

/* estimatePositionKF TOAIMU.cpp:100-195 + kalmanStep3D :242-340 */
static void t9_estimate(ko_t9 *f, double dt, int has_r, const ko_meas *m, int n, int has_imu,
                        ko_info *info) {
    double F[81], Q[81], x[9], xp[9];
    memset(info, 0, sizeof *info);
    info->ignored = -1;
    if (f->ml_init && (isnan(f->pos[0]) || isnan(f->pos[1]))) {
        /* TOAIMU.cpp:118-162: 3-D ML from (1, 1, 4); only the 2x2 x-y block of its covariance is copied */
        info->status |= KO_ST_UNINIT;
        if (has_r) {
            const double start[3] = {1.0, 1.0, 4.0};
            double p[3], c[9];
            int it = 0;
            int rc = ko_ml3d(m, n, start, p, c, &it);
            info->ml_iters += it;
            if (rc < 0) { info->status |= KO_ST_SINGULAR; return; }
            f->pos[0] = p[0]; f->pos[1] = p[1]; f->pos[2] = p[2];
            if (rc == 1) { info->status |= KO_ST_ML_FEW; return; } /* empty covariance: (0,0) throws (:133) */
            f->P[0 * 9 + 0] = c[0]; f->P[1 * 9 + 0] = c[3];
            f->P[0 * 9 + 1] = c[1]; f->P[1 * 9 + 1] = c[4];
        }
        return;
    }
    for (int i = 0; i < 3; ++i) { x[i] = f->pos[i]; x[3 + i] = f->vel[i]; x[6 + i] = f->acc[i]; }
    t9_FQ(f->jolt, dt, F, Q);
    matvec(9, F, x, xp);
    predict_cov(9, F, Q, f->P);
    int nr = has_r ? n : 0, M = nr, iimu = 0;
    if (has_imu) { iimu = M; M += 3; }
    double R[MR * MR], z[MR];
    memset(R, 0, sizeof(double) * M * M);
    for (int i = 0; i < M; ++i) R[i * M + i] = 1;
    if (has_r) { /* TOAIMU.cpp:268-276 (no NaN guard here) */
        double mlp[3], mlcov[9];
        int it = 0;
        int rc = ko_ml3d(m, n, xp, mlp, mlcov, &it);
        info->ml_iters += it;
        if (rc == 1) info->status |= KO_ST_ML_FEW;
        if (rc < 0) { info->status |= KO_ST_SINGULAR; return; }
        double sse = ko_sse(m, n, mlp);
        for (int i = 0; i < n; ++i) {
            z[i] = m[i].r;
            R[i * M + i] = fmax(sse, m[i].e);
        }
        if (n == 0) info->status |= KO_ST_NO_MEAS;
    }
    if (has_imu) {
        for (int a = 0; a < 3; ++a) {
            z[iimu + a] = f->imu_a[a];
            for (int b = 0; b < 3; ++b) R[(iimu + a) * M + iimu + b] = f->imu_cov[a * 3 + b];
        }
    }
    t9_ctx ctx = {m, nr, has_imu};
    double xs[9], Ps[81];
    int rc = iekf_dense(9, M, xp, f->P, z, R, t9_model, &ctx, 20, 1e-4, -1, xs, Ps, info);
    if (rc != 0) { info->status |= KO_ST_SINGULAR; return; }
    memcpy(f->P, Ps, sizeof f->P);
    for (int i = 0; i < 3; ++i) { f->pos[i] = xs[i]; f->vel[i] = xs[3 + i]; } /* :189-194 */
    for (int i = 0; i < 9; ++i)
        if (!isfinite(xs[i])) info->status |= KO_ST_NAN;
}

void ko_t9_new_toa(ko_t9 *f, double dt, int n_slots, const double *ranges, const double *anchors,
                   const double *errs, ko_info *info) {
    ko_meas m[KO_MAX_ANCHORS];
    int n = gather(n_slots, ranges, anchors, errs, m);
    if (f->variant == 0 || (f->ml_init && (isnan(f->pos[0]) || isnan(f->pos[1])))) {
        t9_estimate(f, dt, 1, m, n, f->has_imu, info);
        return;
    }
    /* EKF-side NLOS variants, as ko_t6_new_toa_sel: 3-D ML selection from the predicted position (F x)[0:3] */
    ko_meas sub[KO_MAX_ANCHORS];
    unsigned char keep[KO_MAX_ANCHORS];
    for (int i = 0; i < n; ++i) keep[i] = 1;
    int it_sel = 0, it = 0;
    double Fs[81], Qs[81], xs0[9], xps[9];
    for (int i = 0; i < 3; ++i) { xs0[i] = f->pos[i]; xs0[3 + i] = f->vel[i]; xs0[6 + i] = f->acc[i]; }
    t9_FQ(f->jolt, dt, Fs, Qs);
    matvec(9, Fs, xs0, xps);
    const double start[3] = {xps[0], xps[1], xps[2]};
    if (f->variant == 1 && n > 0) {
        double p0[3], c0[9];
        if (ko_ml3d(m, n, start, p0, c0, &it) == 0) {
            int order[KO_MAX_ANCHORS];
            ko_best_rangings(m, n, p0, order);
            int drop = n - 4 < f->n_ignore ? n - 4 : f->n_ignore;
            if (drop < 0) drop = 0;
            for (int i = n - drop; i < n; ++i) keep[order[i]] = 0;
        }
        it_sel += it;
    } else if (f->variant == 2 && n >= 4) {
        double p0[3], c0[9];
        int bi, ng;
        uint32_t bm = 0;
        ko_ml_best_group(m, n, start, 0, f->best_mode, 0, p0, c0, &it, &bi, &bm, &ng);
        it_sel += it;
        if (bi >= 0) /* a failed solve selects nothing: the update then runs on all rangings */
            for (int i = 0; i < n; ++i) keep[i] = (bm >> i) & 1u;
    }
    int ns = 0;
    for (int i = 0; i < n; ++i)
        if (keep[i]) sub[ns++] = m[i];
    t9_estimate(f, dt, 1, sub, ns, f->has_imu, info);
    info->ml_iters += it_sel;
}

void ko_t9_new_imu(ko_t9 *f, double dt, const double acc[3], const double cov_acc[9],
                   ko_info *info) {
    memcpy(f->imu_a, acc, sizeof f->imu_a);
    memcpy(f->imu_cov, cov_acc, sizeof f->imu_cov);
    f->has_imu = 1;
    t9_estimate(f, dt, 0, NULL, 0, 1, info);
}

void ko_t9_get_pose(const ko_t9 *f, double dt, double xo[9], double Ppred[81]) {
    double F[81], Q[81], x[9];
    for (int i = 0; i < 3; ++i) { x[i] = f->pos[i]; x[3 + i] = f->vel[i]; x[6 + i] = f->acc[i]; }
    t9_FQ(f->jolt, dt, F, Q);
    matvec(9, F, x, xo);
    memcpy(Ppred, f->P, sizeof(double) * 81);
    predict_cov(9, F, Q, Ppred);
}

/* ============================================================ pose message
 * stateToPose of the three filters (TOA.cpp:159-183, KF.cpp:324-363, TOAIMU.cpp:198-241) applied to
 * the predicted state / covariance of getPose, read the way the publisher reads a report
 * (Posgenerator.cpp:385-470): pose13 = x, y, z, rotX, rotY, rotZ, rotW, linearSpeed(3),
 * angularSpeed(3); cov36[i] = covarianceMatrix(i), i < 36, Armadillo's column-major linear index
 * (for T9 the matrix is 9 x 9, so the message gets its first four columns).
 * model: 1 = T6, 2 = K8, 3 = T9.  x / P: predicted state and row-major n x n covariance.  The
 * speed fields stateToPose of T6 never writes are reported as 0. */
void ko_pose_msg(int model, const double *x, const double *P, double tag_z, double pose13[13], double cov36[36]) {
    double C[81];
    int n = 6;
    for (int i = 0; i < 13; ++i) pose13[i] = 0.0;
    for (int i = 0; i < 81; ++i) C[i] = 0.0;
    if (model == 1) {
        pose13[0] = x[0]; pose13[1] = x[1]; pose13[2] = x[2];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) C[i * 6 + j] = P[i * 6 + j];
    } else if (model == 2) {
        const double half = x[6] * 0.5;
        pose13[0] = x[0]; pose13[1] = x[1]; pose13[2] = tag_z;
        pose13[5] = sin(half); pose13[6] = cos(half);
        pose13[7] = x[2]; pose13[8] = x[3];
        pose13[12] = x[7];
        for (int i = 0; i < 6; ++i) C[i * 6 + i] = 0.01;
        C[0 * 6 + 0] = P[0 * 8 + 0]; C[0 * 6 + 1] = P[0 * 8 + 1];
        C[1 * 6 + 0] = P[1 * 8 + 0]; C[1 * 6 + 1] = P[1 * 8 + 1];
        C[0 * 6 + 5] = P[0 * 8 + 6]; C[1 * 6 + 5] = P[1 * 8 + 6];
        C[5 * 6 + 0] = P[6 * 8 + 0]; C[5 * 6 + 1] = P[6 * 8 + 1];
        C[5 * 6 + 5] = P[6 * 8 + 6];
    } else {
        n = 9;
        for (int i = 0; i < 3; ++i) {
            pose13[i] = x[i];
            pose13[7 + i] = x[3 + i];
            pose13[10 + i] = x[6 + i]; /* the acceleration, in the angular-speed fields (:213-215) */
        }
        for (int i = 0; i < 9; ++i) C[i * 9 + i] = 0.01;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) C[i * 9 + j] = P[i * 9 + j];
            C[i * 9 + 7] = P[i * 9 + 8];
            C[7 * 9 + i] = P[8 * 9 + i];
        }
        C[7 * 9 + 7] = P[8 * 9 + 8];
    }
    for (int i = 0; i < 36; ++i) cov36[i] = C[(i % n) * n + i / n];
}
