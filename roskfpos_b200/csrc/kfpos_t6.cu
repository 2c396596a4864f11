// kfpos_t6.cu -- persistent replay kernel for KalmanFilterTOA batches (G1+G2+G3+G5
// of SURVEY.md §2): one thread per filter, position and the packed 6x6
// covariance stay in registers across all T steps; per step the only global
// traffic is the coalesced read of the filter's M rangings (SoA, filter index
// fastest) and the optional trajectory / selection stores.
#include "kfpos_kernels.cuh"
#include "kfpos_t6.cuh"

namespace kfpos {

constexpr int T6_BLOCK = 128;

template <int MAXM, bool PME, bool LOO>
__global__ void __launch_bounds__(T6_BLOCK) t6_replay_kernel(const __grid_constant__ T6Params p) {
    const int64_t f = (int64_t)blockIdx.x * T6_BLOCK + threadIdx.x;
    const bool active = f < p.N;
    StepStats st = {0u, 0u, 0u, 0u};
    unsigned n_updates = 0, n_bad = 0, n_ignored = 0;

    if (active) {
        const int64_t N = p.N;
        double pos[3];
        Sym<6> P;
#pragma unroll
        for (int k = 0; k < 3; ++k) pos[k] = p.x[(int64_t)k * N + f];
#pragma unroll
        for (int k = 0; k < Sym<6>::SZ; ++k) P.a[k] = p.P[(int64_t)k * N + f];
        unsigned status_or = 0;

        for (int t = 0; t < p.T; ++t) {
            // ---- this step's rangings: keep rangings[i] > 0 (TOA.cpp:48-57)
            Epoch<MAXM, PME> ep;
            ep.valid = 0u;
            ep.e[0] = p.rs.err_scalar;
            const int64_t base = (int64_t)t * p.rs.m_slots * N + f;
#pragma unroll
            for (int i = 0; i < MAXM; ++i) {
                ep.z[i] = 0.0;
                if (PME) ep.e[i] = 1.0;
                if (i < p.rs.m_slots) {
                    const double r = load_range(p.rs.ranges, p.rs.fmt, base + (int64_t)i * N);
                    ep.z[i] = r;
                    if (r > 0) ep.valid |= 1u << i;
                    if (PME) ep.e[i] = __ldg(p.rs.err + base + (int64_t)i * N);
                }
            }
            const double dt = __ldg(p.dt + t);

            // ---- predict (TOA.cpp:115-123): x^- = F x with v = 0, P^- = F P F^T + Q.
            // The member covariance is overwritten before the try block, so P^-
            // is what survives a failed update.
            t6_predict_cov(P, dt, p.accel_noise);

            st.status = 0u;
            if (ep.valid == 0u) st.status |= 1u;
            Sym<6> Pw;
            double dx[6], cost;
            int rc = t6_update<MAXM, PME>(p.anchors, ep, ep.valid, pos, P, Pw, dx, cost, st);
            int ignored = -1;
            if (LOO) {
                // kalmanStep3DCanIgnoreAnAnchor (TOA.cpp:185-238): only with > 4 rangings
                if (rc == 0 && __popc(ep.valid) > 4) {
                    double maxDist = 0.0, worstCost = 0.0, dxb[6];
                    Sym<6> Pb;
                    int idx = -1;
                    bool first = true;
                    for (int i = 0; i < MAXM && rc == 0; ++i) {
                        if (!((ep.valid >> i) & 1u)) continue;
                        Sym<6> Pi;
                        double dxi[6], ci;
                        rc = t6_update<MAXM, PME>(p.anchors, ep, ep.valid & ~(1u << i), pos, P, Pi,
                                                  dxi, ci, st);
                        if (rc != 0) break;
                        const double ex = p.anchors.x[i] - (pos[0] + dxi[0]);
                        const double ey = p.anchors.y[i] - (pos[1] + dxi[1]);
                        const double ez = p.anchors.z[i] - (pos[2] + dxi[2]);
                        const double diff = ep.z[i] - sqrt(ex * ex + ey * ey + ez * ez);
                        if (first || diff > maxDist) { // strict >, first seeds (TOA.cpp:209)
                            maxDist = diff;
                            worstCost = ci;
                            Pb = Pi;
#pragma unroll
                            for (int k = 0; k < 6; ++k) dxb[k] = dxi[k];
                            idx = i;
                            first = false;
                        }
                    }
                    if (rc == 0 && maxDist > 0 && (cost - worstCost) > p.ignore_thr) {
                        Pw = Pb;
#pragma unroll
                        for (int k = 0; k < 6; ++k) dx[k] = dxb[k];
                        ignored = idx;
                        n_ignored += 1;
                    }
                }
            }
            if (rc == 0) {
                // stateToPose (TOA.cpp:159-183): position kept, velocity dropped
                pos[0] += dx[0]; pos[1] += dx[1]; pos[2] += dx[2];
                P = Pw;
                if (!(isfinite(pos[0]) && isfinite(pos[1]) && isfinite(pos[2]))) st.status |= 8u;
            } else {
                st.status |= 4u; // catch (std::runtime_error): update skipped (TOA.cpp:151)
            }
            n_updates += 1;
            if (st.status & ~32u) n_bad += 1;
            status_or |= st.status;
            if (p.traj) {
#pragma unroll
                for (int k = 0; k < 3; ++k) p.traj[((int64_t)t * 3 + k) * N + f] = pos[k];
            }
            if (p.sel) p.sel[(int64_t)t * N + f] = ignored;
        }

#pragma unroll
        for (int k = 0; k < 3; ++k) p.x[(int64_t)k * N + f] = pos[k];
#pragma unroll
        for (int k = 0; k < Sym<6>::SZ; ++k) p.P[(int64_t)k * N + f] = P.a[k];
        if (p.status) p.status[f] |= (int32_t)status_or;
    }
    warp_accumulate(p.counters + CNT_UPDATES, n_updates);
    warp_accumulate(p.counters + CNT_ML_ITERS, st.ml_iters);
    warp_accumulate(p.counters + CNT_COST_EVALS, st.cost_evals);
    warp_accumulate(p.counters + CNT_GAIN_EVALS, st.gain_evals);
    warp_accumulate(p.counters + CNT_BAD, n_bad);
    warp_accumulate(p.counters + CNT_IGNORED, n_ignored);
}

template <int MAXM>
static cudaError_t launch_m(const T6Params &p, cudaStream_t s) {
    const unsigned grid = (unsigned)((p.N + T6_BLOCK - 1) / T6_BLOCK);
    const bool pme = p.rs.err != nullptr;
    if (p.ignore_worst) {
        if (pme) t6_replay_kernel<MAXM, true, true><<<grid, T6_BLOCK, 0, s>>>(p);
        else t6_replay_kernel<MAXM, false, true><<<grid, T6_BLOCK, 0, s>>>(p);
    } else {
        if (pme) t6_replay_kernel<MAXM, true, false><<<grid, T6_BLOCK, 0, s>>>(p);
        else t6_replay_kernel<MAXM, false, false><<<grid, T6_BLOCK, 0, s>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t launch_t6_replay(const T6Params &p, cudaStream_t s) {
    if (p.N <= 0 || p.T <= 0) return cudaSuccess;
    const int m = p.rs.m_slots;
    if (m <= 4) return launch_m<4>(p, s);
    if (m <= 8) return launch_m<8>(p, s);
    if (m <= 16) return launch_m<16>(p, s);
    return launch_m<32>(p, s);
}

// getPose (TOA.cpp:438-473): predict-only, state untouched
__global__ void t6_get_pose_kernel(int64_t N, double dt, double accel_noise, const double *x,
                                   const double *P, double *x_pred, double *P_full) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= N) return;
    Sym<6> S;
#pragma unroll
    for (int k = 0; k < Sym<6>::SZ; ++k) S.a[k] = P[(int64_t)k * N + f];
    t6_predict_cov(S, dt, accel_noise);
    if (x_pred) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // velocity is always 0 at the start of a step, so F x = x
            x_pred[(int64_t)k * N + f] = x[(int64_t)k * N + f];
            x_pred[(int64_t)(3 + k) * N + f] = 0.0;
        }
    }
    if (P_full) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) P_full[(int64_t)(i * 6 + j) * N + f] = S.get(i, j);
    }
}

cudaError_t launch_t6_get_pose(int64_t N, double dt, double accel_noise, const double *x,
                               const double *P, double *x_pred, double *P_pred_full, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    t6_get_pose_kernel<<<(unsigned)((N + 127) / 128), 128, 0, s>>>(N, dt, accel_noise, x, P, x_pred,
                                                                  P_pred_full);
    return cudaGetLastError();
}

} // namespace kfpos
