// kfpos_mlk.cu -- batched MLLocation epochs (G4 of SURVEY.md §2): one thread per
// epoch; variants NORMAL / IGNORE_N / BEST (ML.cpp:307-414, 421-469).
#include "kfpos_kernels.cuh"
#include "kfpos_ml.cuh"

namespace kfpos {

constexpr int ML_BLOCK = 128;

template <int MAXM, bool PME>
KF_DEV int ml_any(const AnchorTable &A, const Epoch<MAXM, PME> &ep, unsigned mask, bool use2d,
                  const double (&start)[3], double (&pos)[3], double (&cov)[6], double &sse,
                  unsigned &iters) {
    pos[0] = start[0]; pos[1] = start[1]; pos[2] = start[2];
    if (use2d) {
        double c2[3] = {0, 0, 0};
        const int rc = ml_solve2<MAXM, PME>(A, ep, mask, pos, sse, iters, c2);
        cov[0] = c2[0]; cov[1] = c2[1]; cov[2] = c2[2];
        cov[3] = cov[4] = cov[5] = 0.0;
        return rc;
    }
    return ml_solve3<MAXM, PME>(A, ep, mask, pos, sse, iters, cov);
}

template <int MAXM, bool PME>
__global__ void __launch_bounds__(ML_BLOCK) ml_solve_kernel(const __grid_constant__ MlParams p) {
    const int64_t f = (int64_t)blockIdx.x * ML_BLOCK + threadIdx.x;
    const bool active = f < p.N;
    unsigned iters = 0, bad = 0;
    if (active) {
        const int64_t N = p.N;
        Epoch<MAXM, PME> ep;
        ep.valid = 0u;
        ep.e[0] = p.rs.err_scalar;
#pragma unroll
        for (int i = 0; i < MAXM; ++i) {
            ep.z[i] = 0.0;
            if (PME) ep.e[i] = 1.0;
            if (i < p.rs.m_slots) {
                const double r = load_range(p.rs.ranges, p.rs.fmt, (int64_t)i * N + f);
                ep.z[i] = r;
                if (r > 0) ep.valid |= 1u << i; // ML.cpp:478
                if (PME) ep.e[i] = __ldg(p.rs.err + (int64_t)i * N + f);
            }
        }
        const bool use2d = p.use2d != 0;
        const int k = use2d ? 3 : 4; // minRangings (ML.cpp:316,319)
        const double start[3] = {p.start[0], p.start[1], p.start[2]};
        const int n = __popc(ep.valid);
        double pos[3], cov[6] = {0, 0, 0, 0, 0, 0}, sse;
        unsigned used = ep.valid;
        int index = -1;
        int rc = ml_any<MAXM, PME>(p.anchors, ep, ep.valid, use2d, start, pos, cov, sse, iters);

        if (p.variant == 1 && rc != ML_SINGULAR) {
            // estimatePositionIgnoreN (ML.cpp:307-347): drop the tail of the
            // ascending residual order; ties keep the lower index (App. B-11).
            int drop = min(n - k, p.n_ignore);
            if (drop < 0) drop = 0;
            for (int dcount = 0; dcount < drop; ++dcount) {
                double worst = -1.0;
                int wi = -1;
#pragma unroll
                for (int i = 0; i < MAXM; ++i) {
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
            rc = ml_any<MAXM, PME>(p.anchors, ep, used, use2d, start, pos, cov, sse, iters);
        } else if (p.variant == 2 && rc != ML_SINGULAR && n >= k) {
            // estimatePositionBestGroup (ML.cpp:351-414): all C(n,k) subsets in
            // prev_permutation (= lexicographic) order; `<=` keeps the last minimum.
            // App. B-3: subset = measurements with mask true; B-4: 2-D criterion
            // = cov(0,0)+cov(1,1); best_mode 1 = cov(2,2) (3-D only).
            int slot[MAXM];
            int c = 0;
            for (int i = 0; i < MAXM; ++i)
                if ((ep.valid >> i) & 1u) slot[c++] = i;
            double minErr = 0.0;
            int minIdx = -1, gi = 0;
            int a[4] = {0, 1, 2, 3};
            while (true) {
                unsigned m = 0u;
                for (int j = 0; j < k; ++j) m |= 1u << slot[a[j]];
                double gp[3], gc[6] = {0, 0, 0, 0, 0, 0}, gs;
                const int grc = ml_any<MAXM, PME>(p.anchors, ep, m, use2d, start, gp, gc, gs, iters);
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
                    used = m;
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

template <int MAXM>
static cudaError_t launch_m(const MlParams &p, cudaStream_t s) {
    const unsigned grid = (unsigned)((p.N + ML_BLOCK - 1) / ML_BLOCK);
    if (p.rs.err) ml_solve_kernel<MAXM, true><<<grid, ML_BLOCK, 0, s>>>(p);
    else ml_solve_kernel<MAXM, false><<<grid, ML_BLOCK, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_ml_solve(const MlParams &p, cudaStream_t s) {
    if (p.N <= 0) return cudaSuccess;
    const int m = p.rs.m_slots;
    if (m <= 4) return launch_m<4>(p, s);
    if (m <= 8) return launch_m<8>(p, s);
    if (m <= 16) return launch_m<16>(p, s);
    return launch_m<32>(p, s);
}

} // namespace kfpos
