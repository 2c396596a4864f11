// kfpos_ml.cuh -- per-thread ML (Newton) multilateration, the B200 formulation of
// MLLocation::estimatePosition / estimatePosition2D (ML.cpp:48-257).
//
// One thread owns one epoch.  Each Newton iteration is ONE pass over the anchor
// table that yields, at the current point, the stopping cost, the unweighted SSE
// (the EKFs' mlRangingError), the gradient and the Hessian together -- the
// reference evaluates the distances three times per iteration (ML.cpp:171,216,229).
#pragma once
#include "kfpos_math.cuh"

namespace kfpos {

// Valid rangings of one epoch: z[i] metres for the slots set in `valid`.
// PME = per-measurement errorEstimation array; otherwise one scalar for all.
template <int MAXM, bool PME>
struct Epoch {
    double z[MAXM];
    double e[PME ? MAXM : 1];
    unsigned valid;
    KF_DEV double err(int i) const { return PME ? e[i] : e[0]; }
};

struct MlPass3 {
    double wcost, sse;
    double g[3];
    double H[6]; // packed Sym<3>: xx, xy, yy, xz, yz, zz
};

// gradient / Hessian / costs at p over the slots in `mask` (ML.cpp:171-222)
template <int MAXM, bool PME>
KF_DEV void ml_pass3(const AnchorTable &A, const Epoch<MAXM, PME> &ep, unsigned mask,
                     const double (&p)[3], MlPass3 &o) {
    o.wcost = 0.0; o.sse = 0.0;
    o.g[0] = o.g[1] = o.g[2] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) o.H[k] = 0.0;
    const double inv_e0 = 1.0 / ep.e[0];
#pragma unroll
    for (int i = 0; i < MAXM; ++i) {
        if (!((mask >> i) & 1u)) continue;
        const double dx = A.x[i] - p[0], dy = A.y[i] - p[1], dz = A.z[i] - p[2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        const double d = sqrt(d2);
        const double invd = 1.0 / d;
        const double w = PME ? 1.0 / ep.e[i] : inv_e0;
        const double r = ep.z[i];
        const double res = r - d;
        o.sse = fma(res, res, o.sse);
        o.wcost = fma(res * res, w, o.wcost);
        const double t = res * invd * w;
        o.g[0] = fma(t, dx, o.g[0]);
        o.g[1] = fma(t, dy, o.g[1]);
        o.g[2] = fma(t, dz, o.g[2]);
        const double rid = r * invd;
        const double c1 = (1.0 - rid) * w;
        const double c2 = rid * invd * invd * w;
        o.H[0] += fma(c2 * dx, dx, c1);
        o.H[2] += fma(c2 * dy, dy, c1);
        o.H[5] += fma(c2 * dz, dz, c1);
        o.H[1] = fma(c2 * dx, dy, o.H[1]);
        o.H[3] = fma(c2 * dx, dz, o.H[3]);
        o.H[4] = fma(c2 * dy, dz, o.H[4]);
    }
}

// return codes of the ML solvers
#define ML_OK 0
#define ML_FEW 1       // fewer than minRangings: position = start (ML.cpp:54-58,158-161)
#define ML_SINGULAR -1 // arma::solve / inv would throw

// estimatePosition (3-D), ML.cpp:153-257.  p: in = start, out = estimate.
// sse_out = estimationError at the returned point.  The covariance
// inv(J^T W^-1 J) (ML.cpp:229-254) is produced only when cov != nullptr.
template <int MAXM, bool PME>
KF_DEV int ml_solve3(const AnchorTable &A, const Epoch<MAXM, PME> &ep, unsigned mask, double (&p)[3],
                     double &sse_out, unsigned &iters, double *cov /* packed Sym<3> or null */) {
    MlPass3 ps;
    ml_pass3<MAXM, PME>(A, ep, mask, p, ps);
    sse_out = ps.sse;
    if (__popc(mask) < 4) return ML_FEW;
    double cost = 1e20, newCost = 1.0;
    unsigned iter = 0;
    while ((fabs(cost - newCost) / cost > 1e-3) && (iter < 10000u)) {
        iter += 1;
        cost = newCost;
        double s[3];
        if (!solve_sym3(ps.H, ps.g, s)) { iters += iter; return ML_SINGULAR; }
        // newPos = solve(H, H pos - g)  ==  pos - H^-1 g
        p[0] -= s[0]; p[1] -= s[1]; p[2] -= s[2];
        ml_pass3<MAXM, PME>(A, ep, mask, p, ps);
        newCost = ps.wcost;
    }
    iters += iter;
    sse_out = ps.sse;
    if (cov) {
        // J_i = (p - b_i)/d_i ; W = diag(max(e_i, SSE)) ; cov = inv(J^T W^-1 J)
        double M[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < MAXM; ++i) {
            if (!((mask >> i) & 1u)) continue;
            const double dx = p[0] - A.x[i], dy = p[1] - A.y[i], dz = p[2] - A.z[i];
            const double id2 = 1.0 / (dx * dx + dy * dy + dz * dz);
            const double w = id2 / fmax(ep.err(i), ps.sse);
            M[0] = fma(w * dx, dx, M[0]);
            M[1] = fma(w * dx, dy, M[1]);
            M[2] = fma(w * dy, dy, M[2]);
            M[3] = fma(w * dx, dz, M[3]);
            M[4] = fma(w * dy, dz, M[4]);
            M[5] = fma(w * dz, dz, M[5]);
        }
        double I[6];
        if (!inv_sym3(M, I)) return ML_SINGULAR;
#pragma unroll
        for (int k = 0; k < 6; ++k) cov[k] = I[k];
    }
    return ML_OK;
}

struct MlPass2 {
    double sse;
    double g[2];
    double H[3]; // xx, xy, yy
};

// 2-D pass: distances are 3-D with z fixed, derivatives in x,y only (ML.cpp:74-95)
template <int MAXM, bool PME>
KF_DEV void ml_pass2(const AnchorTable &A, const Epoch<MAXM, PME> &ep, unsigned mask, double px,
                     double py, double pz, MlPass2 &o) {
    o.sse = 0.0;
    o.g[0] = o.g[1] = 0.0;
    o.H[0] = o.H[1] = o.H[2] = 0.0;
    const double inv_e0 = 1.0 / ep.e[0];
#pragma unroll
    for (int i = 0; i < MAXM; ++i) {
        if (!((mask >> i) & 1u)) continue;
        const double dx = A.x[i] - px, dy = A.y[i] - py, dz = A.z[i] - pz;
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        const double invd = 1.0 / d;
        const double w = PME ? 1.0 / ep.e[i] : inv_e0;
        const double r = ep.z[i];
        const double res = r - d;
        o.sse = fma(res, res, o.sse);
        const double t = res * invd * w;
        o.g[0] = fma(t, dx, o.g[0]);
        o.g[1] = fma(t, dy, o.g[1]);
        const double rid = r * invd;
        const double c1 = (1.0 - rid) * w;
        const double c2 = rid * invd * invd * w;
        o.H[0] += fma(c2 * dx, dx, c1);
        o.H[2] += fma(c2 * dy, dy, c1);
        o.H[1] = fma(c2 * dx, dy, o.H[1]);
    }
}

// estimatePosition2D, ML.cpp:48-143.  z stays at its start value.  The damping
// `step` of the reference never takes effect: a rejected step leaves
// newCost == cost, so the while-test fails on the next evaluation (ML.cpp:109-116).
// B-1: the tentative cost is evaluated at z = start z.
template <int MAXM, bool PME>
KF_DEV int ml_solve2(const AnchorTable &A, const Epoch<MAXM, PME> &ep, unsigned mask, double (&p)[3],
                     double &sse_out, unsigned &iters, double *cov /* xx, xy, yy or null */) {
    MlPass2 ps;
    ml_pass2<MAXM, PME>(A, ep, mask, p[0], p[1], p[2], ps);
    sse_out = ps.sse;
    if (__popc(mask) < 3) return ML_FEW;
    double cost = 1e20, newCost = ps.sse;
    unsigned iter = 0;
    while ((fabs(cost - newCost) / cost > 1e-3) && (iter < 10000u)) {
        iter += 1;
        cost = newCost;
        double s0, s1;
        if (!solve_sym2(ps.H[0], ps.H[1], ps.H[2], ps.g[0], ps.g[1], s0, s1)) {
            iters += iter;
            return ML_SINGULAR;
        }
        const double nx = p[0] - s0, ny = p[1] - s1;
        MlPass2 pt;
        ml_pass2<MAXM, PME>(A, ep, mask, nx, ny, p[2], pt);
        if (pt.sse > cost) break; // step /= 2; position kept; loop ends
        newCost = pt.sse;
        p[0] = nx; p[1] = ny;
        ps = pt;
    }
    iters += iter;
    sse_out = ps.sse;
    if (cov) {
        double m00 = 0, m01 = 0, m11 = 0;
#pragma unroll
        for (int i = 0; i < MAXM; ++i) {
            if (!((mask >> i) & 1u)) continue;
            const double dx = p[0] - A.x[i], dy = p[1] - A.y[i], dz = p[2] - A.z[i];
            const double id2 = 1.0 / (dx * dx + dy * dy + dz * dz);
            const double w = id2 / fmax(ep.err(i), ps.sse);
            m00 = fma(w * dx, dx, m00);
            m01 = fma(w * dx, dy, m01);
            m11 = fma(w * dy, dy, m11);
        }
        const double det = m00 * m11 - m01 * m01;
        if (!(det != 0.0)) return ML_SINGULAR;
        const double id = 1.0 / det;
        cov[0] = m11 * id;
        cov[1] = -m01 * id;
        cov[2] = m00 * id;
    }
    return ML_OK;
}

} // namespace kfpos
