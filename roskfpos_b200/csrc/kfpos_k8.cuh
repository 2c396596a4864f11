// kfpos_k8.cuh -- KalmanFilter (8-state planar iterated EKF fusing UWB rangings,
// PX4Flow, IMU and magnetometer/compass), per-thread formulation.
// Reference: algorithms/KalmanFilter.cpp.  State [px,py,vx,vy,ax,ay,theta,omega].
#pragma once
#include "kfpos_kernels.cuh" // K8Cfg, EventDesc
#include "kfpos_math.cuh"
#include "kfpos_solve.cuh"

namespace kfpos {

// symmetric congruence with the elementary matrix E = I + c e_i e_k^T :  P <- E P E^T
template <int N>
KF_DEV void sym_add_row(Sym<N> &P, int i, int k, double c) {
    P.at(i, i) = fma(c, fma(c, P.get(k, k), 2.0 * P.get(i, k)), P.get(i, i));
#pragma unroll
    for (int b = 0; b < N; ++b)
        if (b != i) P.at(i, b) = fma(c, P.get(k, b), P.get(i, b));
}

// P^- = F P F^T + Q (KF.cpp:583-609).  F = I + N with N(p,v)=t, N(p,a)=t^2/2,
// N(v,a)=t per axis and N(theta,omega)=t, applied as elementary congruences
// (E(p<-v,t) E(p<-a,-t^2/2) E(v<-a,t) = F restricted to one axis).
KF_DEV void k8_predict_cov(Sym<8> &P, double t, double accel_noise, double jolt) {
    const double h = -0.5 * t * t;
#pragma unroll
    for (int ax = 0; ax < 2; ++ax) {
        sym_add_row<8>(P, 2 + ax, 4 + ax, t);
        sym_add_row<8>(P, ax, 4 + ax, h);
        sym_add_row<8>(P, ax, 2 + ax, t);
    }
    sym_add_row<8>(P, 6, 7, t);
    const double t3 = t * t * t / 6, t2 = t * t / 2;
    const double u[3] = {t3, t2, t};
#pragma unroll
    for (int ax = 0; ax < 2; ++ax)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) P.at(ax + 2 * a, ax + 2 * b) += jolt * u[b] * u[a];
    P.at(6, 6) += accel_noise * t2 * t2;
    P.at(7, 6) += accel_noise * t2 * t;
    P.at(7, 7) += accel_noise * t * t;
}

// what one kalmanStep3D call fuses (KF.cpp:365-501): row layout [ranges | px4(3) | imu(3) | mag(1)]
struct K8Meas {
    bool has_px4, has_imu, has_mag;
    double imu_c00, imu_c01, imu_c11, imu_cw; // covarianceAccelerationXY, covarianceAngularVelocityZ
    double px4_cg, mag_c;                     // covarianceGyroZ, covarianceMag (XML constants)
    // latched samples (shared-memory column, read where they are used so that they do not
    // occupy registers across the iteration):
    //   [0..3] PX4Flow vx, vy, gz, cv   [4..6] IMU ax, ay, wz   [7] magnetometer angle
    Col latch;
};

// kalmanStep3D (KF.cpp:365-501).  xp: predicted state with theta already wrapped
// (KF.cpp:305); has_r: ranging rows present (ep.valid may still be empty).
//   Pm : P^- (shared-memory column, read only);  Pw: out = (I - K J) P^-;  dx: out = x - x^-
// One IEKF iteration = one pass over the anchors (cost + information-form accumulators of the
// ranging rows, kfpos_solve.cuh) + the sensor residuals; its gain step applies the ranging rows
// as ONE 2x2 block update and the <= 7 sensor rows as sequential scalar updates (the IMU
// accelerometer pair as a 2x2 block: its covariance may carry off-diagonals, KF.cpp:431-434).
// wmask: lanes that run this event together, re-converged after the Newton loop (0 = none).
template <bool PME, int MT>
KF_DEV int k8_update(const AnchorTable &A, const K8Cfg &cfg, double tag_z, const EpochT<PME, MT> &ep, bool has_r,
                     unsigned used, const K8Meas &ms, double dt, const double (&xp)[8], const Col &Pm, Sym<8> &Pw,
                     double (&dx)[8], StepStats &st, unsigned wmask) {
    const unsigned mask = has_r ? used : 0u; // used = ep.valid, or the EKF-side variant's selection
    double sse = -1.0;
    int rc = ML_OK;
    if (has_r) { // inner ML 2-D solve from (x^-_0, x^-_1, tagZ) (KF.cpp:403-405)
        double pml[3] = {xp[0], xp[1], tag_z};
        rc = ml_solve2<PME, MT>(A, ep, mask, pml, sse, st.ml_iters, nullptr, nullptr, cfg.zero_tz != 0);
        if (mask == 0u) sse = -1.0; // estimationError of an empty list
        // has_r is a property of the event, common to the batch: every lane of wmask is here
        if (wmask) __syncwarp(wmask);
    }
    if (rc == ML_FEW) st.status |= 2u;
    if (rc == ML_SINGULAR) return ML_SINGULAR;
    const double invR0 = mask ? fast_rcp(fmax(sse, ep.e0)) : 0.0; // no ranging rows: nothing is weighted by it
    // inverse of the IMU accelerometer block and the scalar variances (arma::inv of the
    // block-diagonal observationCovariance, KF.cpp:446)
    const double idet = ms.has_imu ? 1.0 / (ms.imu_c00 * ms.imu_c11 - ms.imu_c01 * ms.imu_c01) : 0.0;
    const double ii00 = ms.imu_c11 * idet, ii01 = -ms.imu_c01 * idet, ii11 = ms.imu_c00 * idet;
    const double px4_cv = ms.has_px4 ? ms.latch[3] : 1.0;
    const double i_cv = ms.has_px4 ? 1.0 / px4_cv : 0.0, i_cg = ms.has_px4 ? 1.0 / ms.px4_cg : 0.0;
    const double i_cw = ms.has_imu ? 1.0 / ms.imu_cw : 0.0, i_cm = ms.has_mag ? 1.0 / ms.mag_c : 0.0;

#pragma unroll
    for (int k = 0; k < 8; ++k) dx[k] = 0.0;
    double cost = 1e20, prior = 0.0;
    bool broke = false;
    for (int iter = 0; iter < 20; ++iter) {
        const double vx = xp[2] + dx[2], vy = xp[3] + dx[3];
        const double ax = xp[4] + dx[4], ay = xp[5] + dx[5];
        const double th = xp[6] + dx[6], om = xp[7] + dx[7];
        // ---- sensor outputs and cost at the current iterate (KF.cpp:451-469); the ranging pass
        //      also accumulates b = J^T R^-1 y and G = J^T R^-1 J of the rows linearised there
        double c = 0.0, b[3] = {0.0, 0.0, 0.0}, G[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (mask) {
            const double dx3[3] = {dx[0], dx[1], 0.0};
            iekf_pass<PME, MT, 2>(A, ep, mask, sse, xp[0] + dx[0], xp[1] + dx[1], tag_z, dx3, c, b, G);
            if (!PME) c *= invR0;
        }
        double sn = 0.0, cs = 1.0, sw = 0.0, cw = 1.0;
        if (ms.has_px4 || ms.has_imu) fast_sincos(th, &sn, &cs);
        double e_p0 = 0, e_p1 = 0, e_p2 = 0, e_i0 = 0, e_i1 = 0, e_i2 = 0, e_m = 0;
        if (ms.has_px4) { // px4flowOutput (KF.cpp:563-571)
            fast_sincos(om * dt, &sw, &cw);
            const double it = 1.0 / dt;
            e_p0 = ms.latch[0] - (cs * vx + sn * vy + it * ((1.0 - cw) * cfg.arm1 - sw * cfg.arm2));
            e_p1 = ms.latch[1] - (-sn * vx + cs * vy + it * (sw * cfg.arm1 + (1.0 - cw) * cfg.arm2));
            e_p2 = ms.latch[2] - om;
            c += (e_p0 * e_p0 + e_p1 * e_p1) * i_cv + e_p2 * e_p2 * i_cg;
        }
        if (ms.has_imu) { // imuOutput (KF.cpp:573-581)
            e_i0 = ms.latch[4] - (cs * ax + sn * ay);
            e_i1 = ms.latch[5] - (-sn * ax + cs * ay);
            e_i2 = ms.latch[6] - om;
            c += e_i0 * (ii00 * e_i0 + ii01 * e_i1) + e_i1 * (ii01 * e_i0 + ii11 * e_i1) + e_i2 * e_i2 * i_cw;
        }
        if (ms.has_mag) { // residual wrapped once (KF.cpp:461-463)
            e_m = wrap_angle(ms.latch[7] - th);
            c += e_m * e_m * i_cm;
        }
        const double newCost = c + prior;
        st.cost_evals += 1;
        if (rel_change_lt(cost, newCost, 1e-4)) { broke = true; break; }
        cost = newCost;

        // ---- gain step from (x^-, P^-), rows linearised at the iterate (KF.cpp:472-495)
        st.gain_evals += 1;
#pragma unroll
        for (int k = 0; k < Sym<8>::SZ; ++k) Pw.a[k] = Pm[k]; // not before: 72 registers less in the first pass
        double dn[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (mask) { // ranging rows (KF.cpp:611-625): one 2x2 information-form block
            if (!PME) {
                b[0] *= invR0; b[1] *= invR0;
                G[0] *= invR0; G[1] *= invR0; G[2] *= invR0;
            }
            info_block<8, 2>(Pw, dn, b, G);
        }
        // Jacobian entries of the sensor rows at the iterate (KF.cpp:627-697)
        const double j_p06 = -sn * vx + cs * vy, j_p07 = cfg.arm1 * sw - cfg.arm2 * cw;
        const double j_p16 = -cs * vx - sn * vy, j_p17 = cfg.arm1 * cw + cfg.arm2 * sw;
        const double j_i06 = -sn * ax + cs * ay, j_i16 = -cs * ax - sn * ay;
        // y = eps - J delta = eps + J dx   (dx still holds the previous iterate's increment)
        const double y_p0 = e_p0 + (cs * dx[2] + sn * dx[3] + j_p06 * dx[6] + j_p07 * dx[7]);
        const double y_p1 = e_p1 + (-sn * dx[2] + cs * dx[3] + j_p16 * dx[6] + j_p17 * dx[7]);
        const double y_p2 = e_p2 + dx[7];
        const double y_i0 = e_i0 + (cs * dx[4] + sn * dx[5] + j_i06 * dx[6]);
        const double y_i1 = e_i1 + (-sn * dx[4] + cs * dx[5] + j_i16 * dx[6]);
        const double y_i2 = e_i2 + dx[7];
        const double y_m = e_m + dx[6];
        if (ms.has_px4) {
            double h[8] = {0, 0, cs, sn, 0, 0, j_p06, j_p07};
            scalar_update<8, 0xCCu>(Pw, dn, h, y_p0, px4_cv);
            h[2] = -sn; h[3] = cs; h[6] = j_p16; h[7] = j_p17;
            scalar_update<8, 0xCCu>(Pw, dn, h, y_p1, px4_cv);
            h[7] = 1.0;
            scalar_update<8, 0x80u>(Pw, dn, h, y_p2, ms.px4_cg);
        }
        if (ms.has_imu) {
            double h0[8] = {0, 0, 0, 0, cs, sn, j_i06, 0}, h1[8] = {0, 0, 0, 0, -sn, cs, j_i16, 0};
            block2_update<8, 0x70u>(Pw, dn, h0, h1, y_i0, y_i1, ms.imu_c00, ms.imu_c01, ms.imu_c11);
            h0[7] = 1.0;
            scalar_update<8, 0x80u>(Pw, dn, h0, y_i2, ms.imu_cw);
        }
        if (ms.has_mag) {
            double h[8] = {0, 0, 0, 0, 0, 0, 1.0, 0};
            scalar_update<8, 0x40u>(Pw, dn, h, y_m, ms.mag_c);
        }
        // ---- prior term for the next cost: w = J^T R^-1 (y - J Delta), delta^T P^+ delta = w . Delta
        double w[8];
        w[0] = b[0] - (G[0] * dn[0] + G[1] * dn[1]);
        w[1] = b[1] - (G[1] * dn[0] + G[2] * dn[1]);
#pragma unroll
        for (int k = 2; k < 8; ++k) w[k] = 0.0;
        if (ms.has_px4) {
            const double u0 = (y_p0 - (cs * dn[2] + sn * dn[3] + j_p06 * dn[6] + j_p07 * dn[7])) * i_cv;
            const double u1 = (y_p1 - (-sn * dn[2] + cs * dn[3] + j_p16 * dn[6] + j_p17 * dn[7])) * i_cv;
            const double u2 = (y_p2 - dn[7]) * i_cg;
            w[2] += cs * u0 - sn * u1;
            w[3] += sn * u0 + cs * u1;
            w[6] += j_p06 * u0 + j_p16 * u1;
            w[7] += j_p07 * u0 + j_p17 * u1 + u2;
        }
        if (ms.has_imu) {
            const double r0 = y_i0 - (cs * dn[4] + sn * dn[5] + j_i06 * dn[6]);
            const double r1 = y_i1 - (-sn * dn[4] + cs * dn[5] + j_i16 * dn[6]);
            const double u0 = ii00 * r0 + ii01 * r1, u1 = ii01 * r0 + ii11 * r1;
            const double u2 = (y_i2 - dn[7]) * i_cw;
            w[4] += cs * u0 - sn * u1;
            w[5] += sn * u0 + cs * u1;
            w[6] += j_i06 * u0 + j_i16 * u1;
            w[7] += u2;
        }
        if (ms.has_mag) w[6] += (y_m - dn[6]) * i_cm;
        prior = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            prior = fma(w[k], dn[k], prior);
            dx[k] = dn[k];
        }
    }
    if (!broke) st.status |= 32u;
    return 0;
}

// ---- kalmanStep3D for an event that fuses ONE sensor and no rangings (KF.cpp:100-193: a PX4Flow frame,
// an IMU sample or a raw magnetometer sample; 13 of the 15 events of a kfpos_multi macro-step), with the
// covariance held in REGISTERS across predict and update and the covariance update DEFERRED:
// every such event runs exactly two gain steps and three cost evaluations (the first gain step's covariance
// is discarded by the reference as well: each step restarts from P^-, KF.cpp:472-495), so a gain step here
// computes only what the state increment needs -- the <= 3 vectors v_r = P_{r-1} h_r, obtained from the
// columns of P^- and the earlier vectors instead of from an updated matrix -- and the sequential rank-1
// updates are applied once, after the loop, from the vectors of the last gain step:
//     P = fma(-k2_i, v2_j, fma(-k1_i, v1_j, fma(-k0_i, v0_j, P^-_ij)))        k_r = v_r / s_r
// Entry for entry these are the operations of scalar_update / block2_update in the same order, so the
// result is bit-identical to the sequential form (kfpos_math.cuh); the covariance never visits shared
// memory, and the first gain step costs a third of a full one.
//   KIND: EV_IMU (accelerometer pair as a 2x2 block + gyro row), EV_PX4 (two flow rows + gyro row), EV_MAG.
template <int KIND>
KF_DEV void k8_update_light(const K8Cfg &cfg, const K8Meas &ms, double dt, const double (&xp)[8], Sym<8> &P,
                            double (&dx)[8], StepStats &st) {
    // inverse of the IMU accelerometer block and the scalar variances, as in k8_update
    const double idet = KIND == EV_IMU ? 1.0 / (ms.imu_c00 * ms.imu_c11 - ms.imu_c01 * ms.imu_c01) : 0.0;
    const double ii00 = ms.imu_c11 * idet, ii01 = -ms.imu_c01 * idet, ii11 = ms.imu_c00 * idet;
    const double px4_cv = KIND == EV_PX4 ? ms.latch[3] : 1.0;
    const double i_cv = KIND == EV_PX4 ? 1.0 / px4_cv : 0.0, i_cg = KIND == EV_PX4 ? 1.0 / ms.px4_cg : 0.0;
    const double i_cw = KIND == EV_IMU ? 1.0 / ms.imu_cw : 0.0, i_cm = KIND == EV_MAG ? 1.0 / ms.mag_c : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) dx[k] = 0.0;
    double cost = 1e20, prior = 0.0;
    bool broke = false;
    // the vectors of the last gain step and their gains (IMU: k0, k1 are the 2x2 block's)
    double v0[8], v1[8], v2[8], q00 = 0.0, q01 = 0.0, q11 = 0.0, q2 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v0[k] = v1[k] = v2[k] = 0.0;
    for (int iter = 0; iter < 20; ++iter) {
        const double vx = xp[2] + dx[2], vy = xp[3] + dx[3];
        const double ax = xp[4] + dx[4], ay = xp[5] + dx[5];
        const double th = xp[6] + dx[6], om = xp[7] + dx[7];
        double c = 0.0;
        double sn = 0.0, cs = 1.0, sw = 0.0, cw = 1.0;
        if (KIND != EV_MAG) fast_sincos(th, &sn, &cs);
        double e0 = 0, e1 = 0, e2 = 0;
        if (KIND == EV_PX4) { // px4flowOutput (KF.cpp:563-571)
            fast_sincos(om * dt, &sw, &cw);
            const double it = 1.0 / dt;
            e0 = ms.latch[0] - (cs * vx + sn * vy + it * ((1.0 - cw) * cfg.arm1 - sw * cfg.arm2));
            e1 = ms.latch[1] - (-sn * vx + cs * vy + it * (sw * cfg.arm1 + (1.0 - cw) * cfg.arm2));
            e2 = ms.latch[2] - om;
            c += (e0 * e0 + e1 * e1) * i_cv + e2 * e2 * i_cg;
        } else if (KIND == EV_IMU) { // imuOutput (KF.cpp:573-581)
            e0 = ms.latch[4] - (cs * ax + sn * ay);
            e1 = ms.latch[5] - (-sn * ax + cs * ay);
            e2 = ms.latch[6] - om;
            c += e0 * (ii00 * e0 + ii01 * e1) + e1 * (ii01 * e0 + ii11 * e1) + e2 * e2 * i_cw;
        } else { // residual wrapped once (KF.cpp:461-463)
            e0 = wrap_angle(ms.latch[7] - th);
            c += e0 * e0 * i_cm;
        }
        const double newCost = c + prior;
        st.cost_evals += 1;
        if (rel_change_lt(cost, newCost, 1e-4)) { broke = true; break; }
        cost = newCost;
        st.gain_evals += 1;

        double dn[8];
        if (KIND == EV_MAG) {
            const double y = e0 + dx[6];
#pragma unroll
            for (int i = 0; i < 8; ++i) v0[i] = P.get(i, 6) * 1.0;
            const double s = fma(1.0, v0[6], ms.mag_c), nu = y;
            q00 = fast_rcp(s);
            const double g = nu * q00;
#pragma unroll
            for (int i = 0; i < 8; ++i) dn[i] = fma(v0[i], g, 0.0);
            const double wm = (y - dn[6]) * i_cm;
            prior = fma(wm, dn[6], 0.0);
        } else if (KIND == EV_IMU) {
            const double j06 = -sn * ax + cs * ay, j16 = -cs * ax - sn * ay;
            const double y0 = e0 + (cs * dx[4] + sn * dx[5] + j06 * dx[6]);
            const double y1 = e1 + (-sn * dx[4] + cs * dx[5] + j16 * dx[6]);
            const double y2 = e2 + dx[7];
            // rows (cs, sn, j06) and (-sn, cs, j16) on states 4, 5, 6 as one 2x2 block (block2_update<8, 0x70>)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v0[i] = fma(P.get(i, 6), j06, fma(P.get(i, 5), sn, fma(P.get(i, 4), cs, 0.0)));
                v1[i] = fma(P.get(i, 6), j16, fma(P.get(i, 5), cs, fma(P.get(i, 4), -sn, 0.0)));
            }
            const double s00 = fma(j06, v0[6], fma(sn, v0[5], fma(cs, v0[4], ms.imu_c00)));
            const double s01 = fma(j06, v1[6], fma(sn, v1[5], fma(cs, v1[4], ms.imu_c01)));
            const double s11 = fma(j16, v1[6], fma(cs, v1[5], fma(-sn, v1[4], ms.imu_c11)));
            const double id = fast_rcp(s00 * s11 - s01 * s01);
            q00 = s11 * id; q01 = -s01 * id; q11 = s00 * id;
            const double g0 = q00 * y0 + q01 * y1, g1 = q01 * y0 + q11 * y1;
#pragma unroll
            for (int i = 0; i < 8; ++i) dn[i] = fma(v0[i], g0, fma(v1[i], g1, 0.0));
            // gyro row e_7 on the covariance after the block: its column 7, entry (7, i) of the packed matrix
            const double k07 = v0[7] * q00 + v1[7] * q01, k17 = v0[7] * q01 + v1[7] * q11;
#pragma unroll
            for (int i = 0; i < 8; ++i) v2[i] = fma(-k07, v0[i], fma(-k17, v1[i], P.get(7, i))) * 1.0;
            const double s2 = fma(1.0, v2[7], ms.imu_cw), nu2 = fma(-1.0, dn[7], y2);
            q2 = fast_rcp(s2);
            const double g2 = nu2 * q2;
#pragma unroll
            for (int i = 0; i < 8; ++i) dn[i] = fma(v2[i], g2, dn[i]);
            // prior term for the next cost: w = J^T R^-1 (y - J Delta), delta^T P^+ delta = w . Delta
            const double r0 = y0 - (cs * dn[4] + sn * dn[5] + j06 * dn[6]);
            const double r1 = y1 - (-sn * dn[4] + cs * dn[5] + j16 * dn[6]);
            const double u0 = ii00 * r0 + ii01 * r1, u1 = ii01 * r0 + ii11 * r1;
            const double u2 = (y2 - dn[7]) * i_cw;
            const double w4 = 0.0 + (cs * u0 - sn * u1), w5 = 0.0 + (sn * u0 + cs * u1);
            const double w6 = 0.0 + (j06 * u0 + j16 * u1), w7 = 0.0 + u2;
            prior = fma(w7, dn[7], fma(w6, dn[6], fma(w5, dn[5], fma(w4, dn[4], 0.0))));
        } else {
            const double j06 = -sn * vx + cs * vy, j07 = cfg.arm1 * sw - cfg.arm2 * cw;
            const double j16 = -cs * vx - sn * vy, j17 = cfg.arm1 * cw + cfg.arm2 * sw;
            const double y0 = e0 + (cs * dx[2] + sn * dx[3] + j06 * dx[6] + j07 * dx[7]);
            const double y1 = e1 + (-sn * dx[2] + cs * dx[3] + j16 * dx[6] + j17 * dx[7]);
            const double y2 = e2 + dx[7];
            // flow rows on states 2, 3, 6, 7, sequential (scalar_update<8, 0xCC> twice), then the gyro row e_7
#pragma unroll
            for (int i = 0; i < 8; ++i)
                v0[i] = fma(P.get(i, 7), j07, fma(P.get(i, 6), j06, fma(P.get(i, 3), sn, P.get(i, 2) * cs)));
            const double s0 = fma(j07, v0[7], fma(j06, v0[6], fma(sn, v0[3], fma(cs, v0[2], px4_cv))));
            q00 = fast_rcp(s0);
            const double g0 = y0 * q00; // the increment is still zero: nu = y
#pragma unroll
            for (int i = 0; i < 8; ++i) dn[i] = fma(v0[i], g0, 0.0);
            double k0[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) k0[i] = v0[i] * q00;
            // entry (i, j) of the covariance after row 0: fma(-k0[max], v0[min], P(i, j))
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                auto p1 = [&](int j) { return fma(-k0[i > j ? i : j], v0[i > j ? j : i], P.get(i, j)); };
                v1[i] = fma(p1(7), j17, fma(p1(6), j16, fma(p1(3), cs, p1(2) * -sn)));
            }
            const double s1 = fma(j17, v1[7], fma(j16, v1[6], fma(cs, v1[3], fma(-sn, v1[2], px4_cv))));
            q11 = fast_rcp(s1);
            const double nu1 = fma(-j17, dn[7], fma(-j16, dn[6], fma(-cs, dn[3], fma(sn, dn[2], y1))));
            const double g1 = nu1 * q11;
#pragma unroll
            for (int i = 0; i < 8; ++i) dn[i] = fma(v1[i], g1, dn[i]);
            const double k17 = v1[7] * q11;
#pragma unroll
            for (int i = 0; i < 8; ++i) v2[i] = fma(-k17, v1[i], fma(-k0[7], v0[i], P.get(7, i))) * 1.0;
            const double s2 = fma(1.0, v2[7], ms.px4_cg), nu2 = fma(-1.0, dn[7], y2);
            q2 = fast_rcp(s2);
            const double g2 = nu2 * q2;
#pragma unroll
            for (int i = 0; i < 8; ++i) dn[i] = fma(v2[i], g2, dn[i]);
            const double u0 = (y0 - (cs * dn[2] + sn * dn[3] + j06 * dn[6] + j07 * dn[7])) * i_cv;
            const double u1 = (y1 - (-sn * dn[2] + cs * dn[3] + j16 * dn[6] + j17 * dn[7])) * i_cv;
            const double u2 = (y2 - dn[7]) * i_cg;
            const double w2 = 0.0 + (cs * u0 - sn * u1), w3 = 0.0 + (sn * u0 + cs * u1);
            const double w6 = 0.0 + (j06 * u0 + j16 * u1), w7 = 0.0 + (j07 * u0 + j17 * u1 + u2);
            prior = fma(w7, dn[7], fma(w6, dn[6], fma(w3, dn[3], fma(w2, dn[2], 0.0))));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) dx[k] = dn[k];
    }
    if (!broke) st.status |= 32u;
    // ---- the covariance update of the last gain step, applied once
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (KIND == EV_MAG) {
            const double k = v0[i] * q00;
#pragma unroll
            for (int j = 0; j <= i; ++j) P.at(i, j) = fma(-k, v0[j], P.at(i, j));
        } else if (KIND == EV_IMU) {
            const double k0 = v0[i] * q00 + v1[i] * q01, k1 = v0[i] * q01 + v1[i] * q11, k2 = v2[i] * q2;
#pragma unroll
            for (int j = 0; j <= i; ++j) P.at(i, j) = fma(-k2, v2[j], fma(-k0, v0[j], fma(-k1, v1[j], P.at(i, j))));
        } else {
            const double k0 = v0[i] * q00, k1 = v1[i] * q11, k2 = v2[i] * q2;
#pragma unroll
            for (int j = 0; j <= i; ++j) P.at(i, j) = fma(-k2, v2[j], fma(-k1, v1[j], fma(-k0, v0[j], P.at(i, j))));
        }
    }
}


// ---- the IMU event (10 of the 12 / 15 events of a macro-step) as STRAIGHT-LINE code.  Every such event runs
// three cost evaluations and two gain steps (measured 3.000 / 2.000); written as a loop, each evaluation ends a
// basic block at its break test, so the covariance prediction cannot overlap the first evaluation's sincos chain
// and the deferred covariance update cannot overlap the third evaluation.  Here the sequence
//   eval0, gain1, eval1, gain2, [P <- update], eval2
// is one block: the update is applied SPECULATIVELY after the second gain step (the caller has backed P^- up in
// the filter's shared-memory column).  If the break tests do not come out as (continue, continue, stop) the
// function returns false and the caller runs the event through k8_update from the backed-up P^- -- the general
// form, same operations in the same order -- so results and counters are those of the loop form in every case.
KF_DEV bool k8_update_imu_fast(const K8Meas &ms, const double (&xp)[8], Sym<8> &P, double (&dx)[8], StepStats &st) {
    const double idet = 1.0 / (ms.imu_c00 * ms.imu_c11 - ms.imu_c01 * ms.imu_c01);
    const double ii00 = ms.imu_c11 * idet, ii01 = -ms.imu_c01 * idet, ii11 = ms.imu_c00 * idet;
    const double i_cw = 1.0 / ms.imu_cw;
    const double la4 = ms.latch[4], la5 = ms.latch[5], la6 = ms.latch[6];
    double v0[8], v1[8], v2[8], q00, q01, q11, q2, prior;
    double sn, cs, e0, e1, e2;
    auto eval = [&](const double(&d)[8]) { // imuOutput and cost at x^- + d (KF.cpp:451-469, 573-581)
        const double ax = xp[4] + d[4], ay = xp[5] + d[5], th = xp[6] + d[6], om = xp[7] + d[7];
        fast_sincos(th, &sn, &cs);
        e0 = la4 - (cs * ax + sn * ay);
        e1 = la5 - (-sn * ax + cs * ay);
        e2 = la6 - om;
        return 0.0 + (e0 * (ii00 * e0 + ii01 * e1) + e1 * (ii01 * e0 + ii11 * e1) + e2 * e2 * i_cw);
    };
    auto gain = [&](const double(&d)[8], double(&dn)[8]) { // the EV_IMU gain step of k8_update_light, operation for operation
        const double ax = xp[4] + d[4], ay = xp[5] + d[5];
        const double j06 = -sn * ax + cs * ay, j16 = -cs * ax - sn * ay;
        const double y0 = e0 + (cs * d[4] + sn * d[5] + j06 * d[6]);
        const double y1 = e1 + (-sn * d[4] + cs * d[5] + j16 * d[6]);
        const double y2 = e2 + d[7];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v0[i] = fma(P.get(i, 6), j06, fma(P.get(i, 5), sn, fma(P.get(i, 4), cs, 0.0)));
            v1[i] = fma(P.get(i, 6), j16, fma(P.get(i, 5), cs, fma(P.get(i, 4), -sn, 0.0)));
        }
        const double s00 = fma(j06, v0[6], fma(sn, v0[5], fma(cs, v0[4], ms.imu_c00)));
        const double s01 = fma(j06, v1[6], fma(sn, v1[5], fma(cs, v1[4], ms.imu_c01)));
        const double s11 = fma(j16, v1[6], fma(cs, v1[5], fma(-sn, v1[4], ms.imu_c11)));
        const double id = fast_rcp(s00 * s11 - s01 * s01);
        q00 = s11 * id; q01 = -s01 * id; q11 = s00 * id;
        const double g0 = q00 * y0 + q01 * y1, g1 = q01 * y0 + q11 * y1;
#pragma unroll
        for (int i = 0; i < 8; ++i) dn[i] = fma(v0[i], g0, fma(v1[i], g1, 0.0));
        const double k07 = v0[7] * q00 + v1[7] * q01, k17 = v0[7] * q01 + v1[7] * q11;
#pragma unroll
        for (int i = 0; i < 8; ++i) v2[i] = fma(-k07, v0[i], fma(-k17, v1[i], P.get(7, i))) * 1.0;
        const double s2 = fma(1.0, v2[7], ms.imu_cw), nu2 = fma(-1.0, dn[7], y2);
        q2 = fast_rcp(s2);
        const double g2 = nu2 * q2;
#pragma unroll
        for (int i = 0; i < 8; ++i) dn[i] = fma(v2[i], g2, dn[i]);
        const double r0 = y0 - (cs * dn[4] + sn * dn[5] + j06 * dn[6]);
        const double r1 = y1 - (-sn * dn[4] + cs * dn[5] + j16 * dn[6]);
        const double u0 = ii00 * r0 + ii01 * r1, u1 = ii01 * r0 + ii11 * r1;
        const double u2 = (y2 - dn[7]) * i_cw;
        const double w4 = 0.0 + (cs * u0 - sn * u1), w5 = 0.0 + (sn * u0 + cs * u1);
        const double w6 = 0.0 + (j06 * u0 + j16 * u1), w7 = 0.0 + u2;
        prior = fma(w7, dn[7], fma(w6, dn[6], fma(w5, dn[5], fma(w4, dn[4], 0.0))));
    };
    const double z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double d1[8], d2[8];
    const double c0 = eval(z) + 0.0;
    const bool stop0 = rel_change_lt(1e20, c0, 1e-4);
    gain(z, d1);
    const double c1 = eval(d1) + prior;
    const bool stop1 = rel_change_lt(c0, c1, 1e-4);
    gain(d1, d2);
    // the covariance update of the second gain step (k8_update_light's, entry for entry), before the third evaluation
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double k0 = v0[i] * q00 + v1[i] * q01, k1 = v0[i] * q01 + v1[i] * q11, k2 = v2[i] * q2;
#pragma unroll
        for (int j = 0; j <= i; ++j) P.at(i, j) = fma(-k2, v2[j], fma(-k0, v0[j], fma(-k1, v1[j], P.at(i, j))));
    }
    const double c2 = eval(d2) + prior;
    const bool stop2 = rel_change_lt(c1, c2, 1e-4);
    if (stop0 || stop1 || !stop2) return false;
    st.cost_evals += 3;
    st.gain_evals += 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) dx[k] = d2[k];
    return true;
}

} // namespace kfpos
