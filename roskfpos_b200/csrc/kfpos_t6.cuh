// kfpos_t6.cuh -- KalmanFilterTOA (6-state constant-velocity iterated EKF on
// rangings), per-thread formulation.  Reference: algorithms/KalmanFilterTOA.cpp.
#pragma once
#include "kfpos_math.cuh"
#include "kfpos_ml.cuh"

namespace kfpos {

struct StepStats {
    unsigned ml_iters, cost_evals, gain_evals, status;
};

// P^- = F P F^T + Q for F = [[I, tI],[0, I]] and the per-axis Q of
// predictionErrorCovariance (TOA.cpp:362-391), in place on the packed matrix.
KF_DEV void t6_predict_cov(Sym<6> &P, double t, double accel_noise) {
    const double t2 = t * t / 2, a2 = accel_noise * accel_noise;
    // pp block first (needs the old pv and vv blocks)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const double bij = P.get(3 + j, i), bji = P.get(3 + i, j), cij = P.get(3 + i, 3 + j);
            P.at(i, j) = fma(t, (bij + bji) + t * cij, P.get(i, j));
        }
    // pv block: B + t C
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) P.at(3 + j, i) = fma(t, P.get(3 + i, 3 + j), P.get(3 + j, i));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        P.at(i, i) += a2 * t2 * t2;
        P.at(3 + i, i) += a2 * t2 * t;
        P.at(3 + i, 3 + i) += a2 * t * t;
    }
}

// kalmanStep3DIgnoreAnchor (TOA.cpp:242-338) on the slots in `mask`.
//   xp   : predicted position (predicted velocity is 0: TOA.cpp:110-120)
//   Pm   : P^- (kept), Pw: out = (I - K J) P^-
//   dx   : out = x - x^-  (6)
// Returns ML_SINGULAR when a solve failed (the reference's catch at TOA.cpp:151
// then skips the update), else 0.  `cost_out` = last assigned IEKF cost.
template <int MAXM, bool PME>
KF_DEV int t6_update(const AnchorTable &A, const Epoch<MAXM, PME> &ep, unsigned mask,
                     const double (&xp)[3], const Sym<6> &Pm, Sym<6> &Pw, double (&dx)[6],
                     double &cost_out, StepStats &st) {
    // ---- inner ML solve from the predicted position (TOA.cpp:268-273)
    double pml[3] = {xp[0], xp[1], xp[2]};
    double sse;
    const int rc = ml_solve3<MAXM, PME>(A, ep, mask, pml, sse, st.ml_iters, nullptr);
    if (rc == ML_FEW) st.status |= 2u;
    if (rc == ML_SINGULAR) return ML_SINGULAR;
    if (isnan(pml[0]) || isnan(pml[1]) || isnan(pml[2])) { // TOA.cpp:270-272
        st.status |= 16u;
        MlPass3 ps;
        ml_pass3<MAXM, PME>(A, ep, mask, xp, ps);
        sse = ps.sse;
    }
    if (mask == 0u) sse = -1.0; // estimationError of an empty list (ML.cpp:265-267)

    // R_ii = max(mlRangingError, errorEstimation_i) (TOA.cpp:281); inverse kept
    const double R0 = fmax(sse, ep.e[0]);
    const double invR0 = 1.0 / R0;

#pragma unroll
    for (int k = 0; k < 6; ++k) dx[k] = 0.0;
    Pw = Pm;
    double cost = 1e20;
    double prior = 0.0; // delta^T pinv(P^-) delta, pinv-free
    bool broke = false;
    for (int iter = 0; iter < 10; ++iter) {
        const double px = xp[0] + dx[0], py = xp[1] + dx[1], pz = xp[2] + dx[2];
        // ---- pass A: cost at the current iterate (TOA.cpp:297-305)
        double d[MAXM];
        double c = 0.0;
#pragma unroll
        for (int i = 0; i < MAXM; ++i) {
            d[i] = 0.0;
            if (!((mask >> i) & 1u)) continue;
            const double ex = px - A.x[i], ey = py - A.y[i], ez = pz - A.z[i];
            d[i] = sqrt(ex * ex + ey * ey + ez * ez);
            const double eps = ep.z[i] - d[i];
            const double iR = PME ? 1.0 / fmax(sse, ep.e[i]) : invR0;
            c = fma(eps * eps, iR, c);
        }
        const double newCost = c + prior;
        st.cost_evals += 1;
        if (fabs(cost - newCost) / cost < 1e-3) { broke = true; break; }
        cost = newCost;
        // ---- pass B: sequential scalar updates from (x^-, P^-) with the rows
        //      linearised at the current iterate (TOA.cpp:313-320)
        st.gain_evals += 1;
        Pw = Pm;
        double dn[6] = {0, 0, 0, 0, 0, 0};
        double b[3] = {0, 0, 0};          // J^T R^-1 y
        double G[6] = {0, 0, 0, 0, 0, 0}; // J^T R^-1 J (position block, packed)
#pragma unroll
        for (int i = 0; i < MAXM; ++i) {
            if (!((mask >> i) & 1u)) continue;
            const double invd = 1.0 / d[i];
            double h[6];
            h[0] = (px - A.x[i]) * invd;
            h[1] = (py - A.y[i]) * invd;
            h[2] = (pz - A.z[i]) * invd;
            h[3] = h[4] = h[5] = 0.0;
            const double eps = ep.z[i] - d[i];
            // y = eps - J delta, delta = x^- - x = -dx
            const double y = fma(h[0], dx[0], fma(h[1], dx[1], fma(h[2], dx[2], eps)));
            const double R = PME ? fmax(sse, ep.e[i]) : R0;
            const double iR = PME ? 1.0 / R : invR0;
            scalar_update<6, 0x7u>(Pw, dn, h, y, R);
            const double yr = y * iR;
            b[0] = fma(h[0], yr, b[0]);
            b[1] = fma(h[1], yr, b[1]);
            b[2] = fma(h[2], yr, b[2]);
            const double h0r = h[0] * iR, h1r = h[1] * iR, h2r = h[2] * iR;
            G[0] = fma(h0r, h[0], G[0]);
            G[1] = fma(h0r, h[1], G[1]);
            G[2] = fma(h1r, h[1], G[2]);
            G[3] = fma(h0r, h[2], G[3]);
            G[4] = fma(h1r, h[2], G[4]);
            G[5] = fma(h2r, h[2], G[5]);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) dx[k] = dn[k];
        // w = J^T R^-1 (y - J dx) ; delta^T P^+ delta = w . dx   (position part only)
        const double w0 = b[0] - (G[0] * dx[0] + G[1] * dx[1] + G[3] * dx[2]);
        const double w1 = b[1] - (G[1] * dx[0] + G[2] * dx[1] + G[4] * dx[2]);
        const double w2 = b[2] - (G[3] * dx[0] + G[4] * dx[1] + G[5] * dx[2]);
        prior = w0 * dx[0] + w1 * dx[1] + w2 * dx[2];
    }
    if (!broke) st.status |= 32u;
    cost_out = cost;
    return 0;
}

} // namespace kfpos
