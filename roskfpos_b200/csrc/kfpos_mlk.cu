// kfpos_mlk.cu -- batched MLLocation epochs (G4 of SURVEY.md §2): one thread per
// epoch; variants NORMAL / IGNORE_N / BEST (ML.cpp:307-414, 421-469).
#include "kfpos_kernels.cuh"
#include "kfpos_solve.cuh"

namespace kfpos {

constexpr int ML_BLOCK = 128;

template <bool PME>
KF_DEV int ml_any(const AnchorTable &A, const EpochT<PME, 0> &ep, unsigned mask, bool use2d,
                  const double (&start)[3], double (&pos)[3], double (&cov)[6], double &sse,
                  unsigned &iters) {
    pos[0] = start[0]; pos[1] = start[1]; pos[2] = start[2];
    if (use2d) {
        double c2[3] = {0, 0, 0};
        const int rc = ml_solve2<PME, 0>(A, ep, mask, pos, sse, iters, c2);
        cov[0] = c2[0]; cov[1] = c2[1]; cov[2] = c2[2];
        cov[3] = cov[4] = cov[5] = 0.0;
        return rc;
    }
    return ml_solve3<PME, 0>(A, ep, mask, pos, sse, iters, cov);
}

template <bool PME>
__global__ void __launch_bounds__(ML_BLOCK) ml_solve_kernel(const __grid_constant__ MlParams p) {
    extern __shared__ double smem[];
    const int64_t f = (int64_t)blockIdx.x * ML_BLOCK + threadIdx.x;
    const bool active = f < p.N;
    unsigned iters = 0, bad = 0;
    if (active) {
        const int64_t N = p.N;
        const int m = p.rs.m_slots;
        EpochT<PME, 0> ep;
        ep.z = Col{smem + threadIdx.x, ML_BLOCK};
        ep.e = Col{smem + (size_t)(PME ? m : 0) * ML_BLOCK + threadIdx.x, ML_BLOCK};
        ep.e0 = p.rs.err_scalar;
        ep.m_slots = m;
        const RawCol raw = make_raw(smem + (size_t)(PME ? 2 * m : m) * ML_BLOCK, p.rs.fmt, threadIdx.x, ML_BLOCK);
        prefetch_epoch(raw, m, p.rs.ranges, p.rs.fmt, f, N); // all M loads in flight together
        cp_async_wait_all();
        convert_epoch<PME, 0>(ep, raw, p.rs.ranges, p.rs.fmt, p.rs.err, f, N);
        const bool use2d = p.use2d != 0;
        const int k = use2d ? 3 : 4; // minRangings (ML.cpp:316,319)
        const double start[3] = {p.start[0], p.start[1], p.start[2]};
        const int n = __popc(ep.valid);
        double pos[3], cov[6] = {0, 0, 0, 0, 0, 0}, sse;
        unsigned used = ep.valid;
        int index = -1;
        int rc = ml_any<PME>(p.anchors, ep, ep.valid, use2d, start, pos, cov, sse, iters);

        if (p.variant == 1 && rc != ML_SINGULAR) {
            // estimatePositionIgnoreN (ML.cpp:307-347): drop the tail of the
            // ascending residual order; ties keep the lower index (App. B-11).
            int drop = min(n - k, p.n_ignore);
            if (drop < 0) drop = 0;
            for (int dcount = 0; dcount < drop; ++dcount) {
                double worst = -1.0;
                int wi = -1;
                for (int i = 0; i < m; ++i) {
                    if (!((used >> i) & 1u)) continue;
                    const double ex = p.anchors.x[i] - pos[0], ey = p.anchors.y[i] - pos[1],
                                 ez = p.anchors.z[i] - pos[2];
                    const double d = sqrt(ex * ex + ey * ey + ez * ez);
                    const double q = (d - ep.z[i]) * (d - ep.z[i]);
                    if (q >= worst) { worst = q; wi = i; }
                }
                if (wi < 0) break; // all residuals NaN
                used &= ~(1u << wi);
            }
            index = drop;
            rc = ml_any<PME>(p.anchors, ep, used, use2d, start, pos, cov, sse, iters);
        } else if (p.variant == 2 && rc != ML_SINGULAR && n >= k) {
            // estimatePositionBestGroup (ML.cpp:351-414): all C(n,k) subsets in
            // prev_permutation (= lexicographic) order; `<=` keeps the last minimum.
            // App. B-3: subset = measurements with mask true; B-4: 2-D criterion
            // = cov(0,0)+cov(1,1); best_mode 1 = cov(2,2) (3-D only).
            unsigned char slot[32];
            int c = 0;
            for (int i = 0; i < m; ++i)
                if ((ep.valid >> i) & 1u) slot[c++] = (unsigned char)i;
            double minErr = 0.0;
            int minIdx = -1, gi = 0;
            int a[4] = {0, 1, 2, 3};
            while (true) {
                unsigned gm = 0u;
                for (int j = 0; j < k; ++j) gm |= 1u << slot[a[j]];
                double gp[3], gc[6] = {0, 0, 0, 0, 0, 0}, gs;
                const int grc = ml_any<PME>(p.anchors, ep, gm, use2d, start, gp, gc, gs, iters);
                double cur;
                if (use2d) cur = gc[0] + gc[2];
                else if (p.best_mode == 1) cur = gc[5];
                else cur = gc[0] + gc[2] + gc[5];
                if (grc != ML_OK) cur = nan("");
                if (minIdx == -1 || cur <= minErr) {
                    minIdx = gi;
                    minErr = cur;
                    pos[0] = gp[0]; pos[1] = gp[1]; pos[2] = gp[2];
#pragma unroll
                    for (int q = 0; q < 6; ++q) cov[q] = gc[q];
                    used = gm;
                    rc = grc;
                }
                ++gi;
                // next combination in lexicographic order
                int j = k - 1;
                while (j >= 0 && a[j] == n - k + j) --j;
                if (j < 0) break;
                ++a[j];
                for (int q = j + 1; q < k; ++q) a[q] = a[q - 1] + 1;
            }
            index = minIdx;
        }

        if (p.pos) {
#pragma unroll
            for (int q = 0; q < 3; ++q) p.pos[(int64_t)q * N + f] = pos[q];
        }
        if (p.cov) {
            const bool ok = rc == ML_OK;
            // packed (xx,xy,yy,xz,yz,zz) -> 3x3 row-major; 2-D fills the top-left block
            const double c00 = ok ? cov[0] : 0.0, c01 = ok ? cov[1] : 0.0, c11 = ok ? cov[2] : 0.0;
            const double c02 = ok ? cov[3] : 0.0, c12 = ok ? cov[4] : 0.0, c22 = ok ? cov[5] : 0.0;
            p.cov[0 * N + f] = c00; p.cov[1 * N + f] = c01; p.cov[2 * N + f] = c02;
            p.cov[3 * N + f] = c01; p.cov[4 * N + f] = c11; p.cov[5 * N + f] = c12;
            p.cov[6 * N + f] = c02; p.cov[7 * N + f] = c12; p.cov[8 * N + f] = c22;
        }
        if (p.iters) p.iters[f] = (int32_t)iters;
        if (p.sel) {
            p.sel[f] = (int32_t)used;
            p.sel[N + f] = index;
        }
        const int stv = rc == ML_OK ? 0 : (rc == ML_FEW ? 2 : 4);
        if (p.status) p.status[f] = stv;
        bad = stv != 0;
    }
    warp_accumulate(p.counters + CNT_UPDATES, active ? 1u : 0u);
    warp_accumulate(p.counters + CNT_ML_ITERS, iters);
    warp_accumulate(p.counters + CNT_BAD, bad);
}

cudaError_t launch_ml_solve(const MlParams &p, cudaStream_t s) {
    if (p.N <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((p.N + ML_BLOCK - 1) / ML_BLOCK);
    const bool pme = p.rs.err != nullptr;
    const size_t smem = (size_t)p.rs.m_slots * (pme ? 3 : 2) * ML_BLOCK * sizeof(double);
    if (pme) {
        cudaError_t e = cudaFuncSetAttribute(ml_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ml_solve_kernel<true><<<grid, ML_BLOCK, smem, s>>>(p);
    } else {
        cudaError_t e = cudaFuncSetAttribute(ml_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ml_solve_kernel<false><<<grid, ML_BLOCK, smem, s>>>(p);
    }
    return cudaGetLastError();
}

} // namespace kfpos
