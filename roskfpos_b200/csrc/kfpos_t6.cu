// kfpos_t6.cu -- persistent replay kernel for KalmanFilterTOA batches (G1+G2+G3+G5
// of SURVEY.md §2).  One thread per filter; the filter's position and the working
// covariance stay in registers and its P^- / per-anchor scratch in a private
// shared-memory column across all T steps.  Per step the only global traffic is
// the coalesced read of the filter's M rangings (SoA, filter index fastest) and
// the optional trajectory / selection stores.
#include "kfpos_kernels.cuh"
#include "kfpos_t6.cuh"

namespace kfpos {

constexpr int T6_BLOCK = 128;

// shared-memory rows per thread
__host__ __device__ inline int t6_smem_rows(int m, bool pme, bool loo) {
    return 4 * m + (pme ? m : 0) + 21 + (loo ? 2 * (21 + 6) : 0);
}

template <bool PME, bool LOO>
__global__ void __launch_bounds__(T6_BLOCK, LOO ? 2 : 4) t6_replay_kernel(const __grid_constant__ T6Params p) {
    extern __shared__ double smem[];
    const int64_t f = (int64_t)blockIdx.x * T6_BLOCK + threadIdx.x;
    const bool active = f < p.N;
    StepStats st = {0u, 0u, 0u, 0u};
    unsigned n_updates = 0, n_bad = 0, n_ignored = 0;

    if (active) {
        const int64_t N = p.N;
        const int m = p.rs.m_slots;
        // carve this thread's columns
        double *col = smem + threadIdx.x;
        int row = 0;
        auto take = [&](int rows) {
            Col c = {col + (size_t)row * T6_BLOCK, T6_BLOCK};
            row += rows;
            return c;
        };
        Epoch<PME> ep;
        ep.z = take(m);
        ep.e = PME ? take(m) : ep.z;
        ep.e0 = p.rs.err_scalar;
        ep.m_slots = m;
        Col raw = take(m); // landing zone of the cp.async prefetch of the next epoch
        T6Scratch sc;
        sc.invd = take(m);
        sc.eps = take(m);
        sc.Pm = take(21);
        Col s_all = LOO ? take(27) : sc.Pm, s_best = LOO ? take(27) : sc.Pm;

        double pos[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) pos[k] = p.x[(int64_t)k * N + f];
#pragma unroll
        for (int k = 0; k < Sym<6>::SZ; ++k) sc.Pm[k] = p.P[(int64_t)k * N + f];
        unsigned status_or = 0;

        prefetch_epoch(raw, m, p.rs.ranges, p.rs.fmt, f, N);
        for (int t = 0; t < p.T; ++t) {
            cp_async_wait_all();
            convert_epoch<PME>(ep, raw, p.rs.ranges, p.rs.fmt, p.rs.err, (int64_t)t * m * N + f, N);
            if (t + 1 < p.T) prefetch_epoch(raw, m, p.rs.ranges, p.rs.fmt, (int64_t)(t + 1) * m * N + f, N);
            const double dt = __ldg(p.dt + t);

            // ---- predict (TOA.cpp:115-123): x^- = F x with v = 0, P^- = F P F^T + Q.
            // The member covariance is overwritten before the try block, so P^-
            // is what survives a failed update.
            Sym<6> Pw;
#pragma unroll
            for (int k = 0; k < Sym<6>::SZ; ++k) Pw.a[k] = sc.Pm[k];
            t6_predict_cov(Pw, dt, p.accel_noise);
#pragma unroll
            for (int k = 0; k < Sym<6>::SZ; ++k) sc.Pm[k] = Pw.a[k];

            st.status = 0u;
            if (ep.valid == 0u) st.status |= 1u;
            double dx[6], cost;
            int rc = t6_update<PME>(p.anchors, ep, ep.valid, pos, sc, Pw, dx, cost, st);
            int ignored = -1;
            if (LOO) {
                // kalmanStep3DCanIgnoreAnAnchor (TOA.cpp:185-238): only with > 4 rangings
                if (rc == 0 && __popc(ep.valid) > 4) {
#pragma unroll
                    for (int k = 0; k < 21; ++k) s_all[k] = Pw.a[k];
#pragma unroll
                    for (int k = 0; k < 6; ++k) s_all[21 + k] = dx[k];
                    double maxDist = 0.0, worstCost = 0.0;
                    int idx = -1;
                    bool first = true;
                    for (int i = 0; i < m && rc == 0; ++i) {
                        if (!((ep.valid >> i) & 1u)) continue;
                        double ci;
                        rc = t6_update<PME>(p.anchors, ep, ep.valid & ~(1u << i), pos, sc, Pw, dx, ci, st);
                        if (rc != 0) break;
                        const double ex = p.anchors.x[i] - (pos[0] + dx[0]);
                        const double ey = p.anchors.y[i] - (pos[1] + dx[1]);
                        const double ez = p.anchors.z[i] - (pos[2] + dx[2]);
                        const double diff = ep.z[i] - sqrt(ex * ex + ey * ey + ez * ez);
                        if (first || diff > maxDist) { // strict >, first seeds (TOA.cpp:209)
                            maxDist = diff;
                            worstCost = ci;
#pragma unroll
                            for (int k = 0; k < 21; ++k) s_best[k] = Pw.a[k];
#pragma unroll
                            for (int k = 0; k < 6; ++k) s_best[21 + k] = dx[k];
                            idx = i;
                            first = false;
                        }
                    }
                    const bool take_best = rc == 0 && maxDist > 0 && (cost - worstCost) > p.ignore_thr;
                    const Col &src = take_best ? s_best : s_all;
#pragma unroll
                    for (int k = 0; k < 21; ++k) Pw.a[k] = src[k];
#pragma unroll
                    for (int k = 0; k < 6; ++k) dx[k] = src[21 + k];
                    if (take_best) {
                        ignored = idx;
                        n_ignored += 1;
                    }
                }
            }
            if (rc == 0) {
                // stateToPose (TOA.cpp:159-183): position kept, velocity dropped
                pos[0] += dx[0]; pos[1] += dx[1]; pos[2] += dx[2];
#pragma unroll
                for (int k = 0; k < Sym<6>::SZ; ++k) sc.Pm[k] = Pw.a[k];
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
        for (int k = 0; k < Sym<6>::SZ; ++k) p.P[(int64_t)k * N + f] = sc.Pm[k];
        if (p.status) p.status[f] |= (int32_t)status_or;
    }
    warp_accumulate(p.counters + CNT_UPDATES, n_updates);
    warp_accumulate(p.counters + CNT_ML_ITERS, st.ml_iters);
    warp_accumulate(p.counters + CNT_COST_EVALS, st.cost_evals);
    warp_accumulate(p.counters + CNT_GAIN_EVALS, st.gain_evals);
    warp_accumulate(p.counters + CNT_BAD, n_bad);
    warp_accumulate(p.counters + CNT_IGNORED, n_ignored);
}

template <bool PME, bool LOO>
static cudaError_t launch_k(const T6Params &p, cudaStream_t s) {
    const unsigned grid = (unsigned)((p.N + T6_BLOCK - 1) / T6_BLOCK);
    const size_t smem = (size_t)t6_smem_rows(p.rs.m_slots, PME, LOO) * T6_BLOCK * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(t6_replay_kernel<PME, LOO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    t6_replay_kernel<PME, LOO><<<grid, T6_BLOCK, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_t6_replay(const T6Params &p, cudaStream_t s) {
    if (p.N <= 0 || p.T <= 0) return cudaSuccess;
    const bool pme = p.rs.err != nullptr;
    if (p.ignore_worst) return pme ? launch_k<true, true>(p, s) : launch_k<false, true>(p, s);
    return pme ? launch_k<true, false>(p, s) : launch_k<false, false>(p, s);
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
