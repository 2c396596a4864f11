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

// Per-thread scratch columns in shared memory used by one IEKF update.
struct T6Scratch {
    Col Pm;   // 21 rows: P^- (read at the start of every gain pass)
    Col invd; // MAXM rows: 1/dist at the current iterate (pass A -> pass B)
    Col eps;  // MAXM rows: z - dist at the current iterate
};

// kalmanStep3DIgnoreAnchor (TOA.cpp:242-338) on the slots in `mask`.
//   xp : predicted position (predicted velocity is 0: TOA.cpp:110-120)
//   sc.Pm : P^- (kept);  Pw: out = (I - K J) P^- ;  dx: out = x - x^- (6)
// Returns ML_SINGULAR when a solve failed (the reference's catch at TOA.cpp:151
// then skips the update), else 0.  `cost_out` = last assigned IEKF cost.
//
// Formulation (validated against the dense oracle): inside one
// IEKF iteration the gain step is a linear-KF update from (x^-, P^-), processed
// one ranging at a time as rank-1 updates of a register copy of P^-; the prior
// term delta^T pinv(P^-) delta of the cost equals w . dx with
// w = J^T R^-1 (y - J dx), accumulated on the fly (b = J^T R^-1 y, G = J^T R^-1 J).
template <bool PME>
KF_DEV int t6_update(const AnchorTable &A, const Epoch<PME> &ep, unsigned mask, const double (&xp)[3],
                     const T6Scratch &sc, Sym<6> &Pw, double (&dx)[6], double &cost_out, StepStats &st) {
    // ---- inner ML solve from the predicted position (TOA.cpp:268-273)
    double pml[3] = {xp[0], xp[1], xp[2]};
    double sse, sse_xp;
    // the first Newton pass runs at xp: it also leaves 1/d_i and eps_i(xp) in the scratch
    // columns for the IEKF's first cost evaluation
    const DistStore ds = {sc.invd, sc.eps};
    const int rc = ml_solve3<PME, true>(A, ep, mask, pml, sse, st.ml_iters, nullptr, &ds, &sse_xp);
    if (rc == ML_FEW) st.status |= 2u;
    if (rc == ML_SINGULAR) return ML_SINGULAR;
    if (isnan(pml[0]) || isnan(pml[1]) || isnan(pml[2])) { // TOA.cpp:270-272: back to xp
        st.status |= 16u;
        sse = sse_xp;
    }
    if (mask == 0u) sse = -1.0; // estimationError of an empty list (ML.cpp:265-267)

    // R_ii = max(mlRangingError, errorEstimation_i) (TOA.cpp:281)
    const double R0 = fmax(sse, ep.e0);
    const double invR0 = 1.0 / R0;

#pragma unroll
    for (int k = 0; k < 6; ++k) dx[k] = 0.0;
#pragma unroll
    for (int k = 0; k < Sym<6>::SZ; ++k) Pw.a[k] = sc.Pm[k];
    double cost = 1e20;
    double prior = 0.0; // delta^T pinv(P^-) delta
    bool broke = false;
    for (int iter = 0; iter < 10; ++iter) {
        const double px = xp[0] + dx[0], py = xp[1] + dx[1], pz = xp[2] + dx[2];
        // ---- pass A: cost at the current iterate (TOA.cpp:297-305)
        double c = 0.0;
        if (iter == 0) { // x = xp: distances already in the scratch columns
            if (PME) {
                for (int i = 0; i < ep.m_slots; ++i) {
                    if (!((mask >> i) & 1u)) continue;
                    const double e = sc.eps[i];
                    c = fma(e * e, 1.0 / fmax(sse, ep.e[i]), c);
                }
            } else {
                c = sse_xp;
            }
        } else {
#pragma unroll 2
            for (int i = 0; i < ep.m_slots; ++i) {
                if (!((mask >> i) & 1u)) continue;
                const double ex = px - A.x[i], ey = py - A.y[i], ez = pz - A.z[i];
                const double d2 = fma(ez, ez, fma(ey, ey, ex * ex));
                const double id = fast_rsqrt(d2);
                const double e = ep.z[i] - d2 * id;
                sc.invd[i] = id;
                sc.eps[i] = e;
                c = PME ? fma(e * e, 1.0 / fmax(sse, ep.e[i]), c) : fma(e, e, c);
            }
        }
        const double newCost = (PME ? c : c * invR0) + prior;
        st.cost_evals += 1;
        if (fabs(cost - newCost) / cost < 1e-3) { broke = true; break; }
        cost = newCost;
        // ---- pass B: sequential scalar updates from (x^-, P^-) with the rows
        //      linearised at the current iterate (TOA.cpp:313-320)
        st.gain_evals += 1;
        if (iter > 0) {
#pragma unroll
            for (int k = 0; k < Sym<6>::SZ; ++k) Pw.a[k] = sc.Pm[k];
        }
        double dn[6] = {0, 0, 0, 0, 0, 0};
        double b0 = 0, b1 = 0, b2 = 0;                      // J^T R^-1 y
        double G0 = 0, G1 = 0, G2 = 0, G3 = 0, G4 = 0, G5 = 0; // J^T R^-1 J (position block, packed)
#pragma unroll 1
        for (int i = 0; i < ep.m_slots; ++i) {
            if (!((mask >> i) & 1u)) continue;
            const double id = sc.invd[i];
            double h[6];
            h[0] = (px - A.x[i]) * id;
            h[1] = (py - A.y[i]) * id;
            h[2] = (pz - A.z[i]) * id;
            h[3] = h[4] = h[5] = 0.0;
            // y = eps - J delta, delta = x^- - x = -dx
            const double y = fma(h[0], dx[0], fma(h[1], dx[1], fma(h[2], dx[2], sc.eps[i])));
            const double R = PME ? fmax(sse, ep.e[i]) : R0;
            scalar_update<6, 0x7u>(Pw, dn, h, y, R);
            if (PME) {
                const double iR = 1.0 / R;
                const double yr = y * iR;
                b0 = fma(h[0], yr, b0); b1 = fma(h[1], yr, b1); b2 = fma(h[2], yr, b2);
                const double h0r = h[0] * iR, h1r = h[1] * iR, h2r = h[2] * iR;
                G0 = fma(h0r, h[0], G0); G1 = fma(h0r, h[1], G1); G2 = fma(h1r, h[1], G2);
                G3 = fma(h0r, h[2], G3); G4 = fma(h1r, h[2], G4); G5 = fma(h2r, h[2], G5);
            } else { // common R: scale once after the loop
                b0 = fma(h[0], y, b0); b1 = fma(h[1], y, b1); b2 = fma(h[2], y, b2);
                G0 = fma(h[0], h[0], G0); G1 = fma(h[0], h[1], G1); G2 = fma(h[1], h[1], G2);
                G3 = fma(h[0], h[2], G3); G4 = fma(h[1], h[2], G4); G5 = fma(h[2], h[2], G5);
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) dx[k] = dn[k];
        // w = J^T R^-1 (y - J dx) ; delta^T P^+ delta = w . dx   (position part only)
        const double w0 = b0 - (G0 * dx[0] + G1 * dx[1] + G3 * dx[2]);
        const double w1 = b1 - (G1 * dx[0] + G2 * dx[1] + G4 * dx[2]);
        const double w2 = b2 - (G3 * dx[0] + G4 * dx[1] + G5 * dx[2]);
        prior = w0 * dx[0] + w1 * dx[1] + w2 * dx[2];
        if (!PME) prior *= invR0;
    }
    if (!broke) st.status |= 32u;
    cost_out = cost;
    return 0;
}

} // namespace kfpos
