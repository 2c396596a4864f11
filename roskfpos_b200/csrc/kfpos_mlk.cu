// kfpos_mlk.cu -- batched MLLocation epochs (G4 of SURVEY.md §2): one thread per
// epoch; variants NORMAL / IGNORE_N / BEST (ML.cpp:307-414, 421-469).
//
// Straggler queue.  The reference's Newton loop runs until its relative-change test passes
// or 10000 iterations have been made (ML.cpp:165).  From the fixed start point (1,1,4) about
// one 3-D epoch in a thousand oscillates all the way to that cap, and a warp is as slow as
// its slowest lane: without counter-measures 3 % of the warps run 10000 iterations instead of
// ~10 and the batch takes 30x longer.  So the main kernel gives every epoch ML_FIRST_CAP
// iterations; an epoch that needs more PARKS its Newton state (point, cost, iteration count,
// phase) in a queue, and a second launch resumes the parked epochs 32 to a warp.  The resumed
// iteration sequence is the same arithmetic in the same order, so results do not depend on
// where an epoch was parked.
#include "kfpos_kernels.cuh"
#include "kfpos_solve.cuh"

namespace kfpos {

constexpr int ML_BLOCK = 128;
constexpr unsigned ML_FIRST_CAP = 32u;

// parked Newton state of one epoch (one 64-byte record)
struct MlParked {
    int32_t idx;    // epoch index (N < 2^31 per launch is enforced by the API)
    int32_t phase;  // 0: first solve, 1: re-solve of variant IGNORE_N
    uint32_t used;  // slot mask of the running solve
    uint32_t iter;  // Newton iterations of the running solve so far
    uint32_t iters; // iterations of the finished solves of this epoch
    uint32_t _pad;
    double cost;
    double p[3];
    double _pad2;
};
static_assert(sizeof(MlParked) == 64, "queue record layout");

template <bool PME, int MT>
KF_DEV int ml_any(const AnchorTable &A, const EpochT<PME, MT> &ep, unsigned mask, bool use2d, bool zero_tz,
                  double (&pos)[3], double *cov /* [6] or null */, double &sse, unsigned &iters, unsigned cap,
                  MlResume *rs) {
    if (use2d) {
        double c2[3] = {0, 0, 0};
        const int rc = ml_solve2<PME, MT>(A, ep, mask, pos, sse, iters, cov ? c2 : nullptr, nullptr, zero_tz);
        if (cov) {
            cov[0] = c2[0]; cov[1] = c2[1]; cov[2] = c2[2];
            cov[3] = cov[4] = cov[5] = 0.0;
        }
        return rc;
    }
    return ml_solve3<PME, MT>(A, ep, mask, pos, sse, iters, cov, nullptr, cap, rs);
}

// RESUME = false: thread f owns epoch f;  true: thread q owns parked record q.
template <bool PME, int MT, bool RESUME>
__global__ void __launch_bounds__(ML_BLOCK) ml_solve_kernel(const __grid_constant__ MlParams p) {
    extern __shared__ double smem[];
    const int64_t t = (int64_t)blockIdx.x * ML_BLOCK + threadIdx.x;
    MlParked *queue = reinterpret_cast<MlParked *>(p.queue);
    const bool active = RESUME ? t < min(*p.queue_count, p.queue_cap) : t < p.N;
    unsigned iters = 0, bad = 0, done = 0;
    if (active) {
        const int64_t N = p.N;
        const int m = MT > 0 ? MT : p.rs.m_slots;
        MlParked rec;
        if (RESUME) rec = queue[t];
        const int64_t f = RESUME ? (int64_t)rec.idx : t;
        EpochT<PME, MT> ep;
        ep.z = Col{smem + threadIdx.x, ML_BLOCK};
        ep.e = Col{smem + (size_t)(MT > 0 ? 0 : m) * ML_BLOCK + threadIdx.x, ML_BLOCK};
        ep.e0 = p.rs.err_scalar;
        ep.m_slots = m;
        const RawCol raw = make_raw(smem + (size_t)((MT > 0 ? 0 : m) + (PME ? m : 0)) * ML_BLOCK, p.rs.fmt, threadIdx.x,
                                    ML_BLOCK);
        prefetch_epoch(raw, m, p.rs.ranges, p.rs.fmt, f, N); // all M loads in flight together
        cp_async_wait_all();
        convert_epoch<PME, MT>(ep, raw, p.rs.ranges, p.rs.fmt, p.rs.err, f, N);
        const bool use2d = p.use2d != 0;
        const int k = use2d ? 3 : 4; // minRangings (ML.cpp:316,319)
        const int n = __popc(ep.valid);
        int drop = min(n - k, p.n_ignore); // estimatePositionIgnoreN (ML.cpp:322)
        if (drop < 0) drop = 0;

        double pos[3] = {p.start[0], p.start[1], p.start[2]}, cov[6] = {0, 0, 0, 0, 0, 0}, sse;
        unsigned used = ep.valid;
        int phase = 0, index = -1;
        MlResume rs = {0.0, 0u};
        if (RESUME) {
            phase = rec.phase; used = rec.used; iters = rec.iters;
            pos[0] = rec.p[0]; pos[1] = rec.p[1]; pos[2] = rec.p[2];
            rs.cost = rec.cost; rs.iter = rec.iter;
        }
        unsigned cap = (RESUME || p.variant == 2) ? 10000u : ML_FIRST_CAP; // BEST does not park
        int rc;
        bool parked = false;
        for (;;) {
            const bool last = !(p.variant == 1 && phase == 0);
            rc = ml_any<PME, MT>(p.anchors, ep, used, use2d, p.zero_tz != 0, pos, last ? cov : nullptr, sse, iters, cap,
                                 &rs);
            if (rc == ML_MORE) {
                const int slot = atomicAdd(p.queue_count, 1);
                if (slot < p.queue_cap) {
                    rec.idx = (int32_t)f; rec.phase = phase; rec.used = used; rec.iter = rs.iter; rec.iters = iters;
                    rec._pad = 0u; rec.cost = rs.cost; rec.p[0] = pos[0]; rec.p[1] = pos[1]; rec.p[2] = pos[2];
                    rec._pad2 = 0.0;
                    queue[slot] = rec;
                    parked = true;
                    break;
                }
                cap = 10000u; // queue full: finish in place
                continue;
            }
            if (!last && rc != ML_SINGULAR) {
                // estimatePositionIgnoreN (ML.cpp:307-347): drop the tail of the ascending
                // residual order; ties keep the lower index (App. B-11); re-solve from the start
                used = drop_worst<PME, MT>(p.anchors, ep, used, pos, drop);
                phase = 1;
                pos[0] = p.start[0]; pos[1] = p.start[1]; pos[2] = p.start[2];
                rs.iter = 0u;
                continue;
            }
            break;
        }
        if (p.variant == 1 && phase == 1) index = drop;

        if (!parked && p.variant == 2 && rc != ML_SINGULAR && n >= k) {
            // estimatePositionBestGroup (ML.cpp:351-414)
            const double start[3] = {p.start[0], p.start[1], p.start[2]};
            index = best_group<PME, MT>(p.anchors, ep, ep.valid, use2d, p.best_mode, start, iters, pos, cov, used, rc,
                                        p.zero_tz != 0);
        }

        if (!parked) {
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
            int stv = rc == ML_OK ? 0 : (rc == ML_FEW ? 2 : 4);
            // minZ / maxZ (config_pos.xml:22-25): "no estimation if the estimated Z is lower / greater" --
            // the estimate is still written, flagged for the consumer to drop
            if (p.max_z > p.min_z && (pos[2] < p.min_z || pos[2] > p.max_z)) stv |= 128;
            if (p.status) p.status[f] = stv;
            bad = (stv & ~128) != 0;
            done = 1u;
        } else {
            iters = 0u; // counted when the epoch completes
        }
    }
    warp_accumulate(p.counters + CNT_UPDATES, done);
    warp_accumulate(p.counters + CNT_ML_ITERS, iters);
    warp_accumulate(p.counters + CNT_BAD, bad);
}

template <bool PME, int MT>
static cudaError_t launch_k(const MlParams &p, cudaStream_t s) {
    const int m = p.rs.m_slots;
    const size_t smem = (size_t)((MT > 0 ? 0 : m) + (PME ? m : 0) + raw_rows(p.rs.fmt, m)) * ML_BLOCK * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(ml_solve_kernel<PME, MT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ml_solve_kernel<PME, MT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(p.queue_count, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    ml_solve_kernel<PME, MT, false><<<(unsigned)((p.N + ML_BLOCK - 1) / ML_BLOCK), ML_BLOCK, smem, s>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // the parked epochs, densely packed; sized for the whole queue (blocks beyond the count exit at once)
    if (!p.use2d && p.variant != 2)
        ml_solve_kernel<PME, MT, true><<<(unsigned)((p.queue_cap + ML_BLOCK - 1) / ML_BLOCK), ML_BLOCK, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_ml_solve(const MlParams &p, cudaStream_t s) {
    if (p.N <= 0) return cudaSuccess;
    if (p.rs.err != nullptr) return launch_k<true, 0>(p, s);
    // BEST solves k-anchor subsets: the branch-skipping rolled loops do k, not m, anchors of work
    if (p.variant != 2 && !p.zero_tz && p.rs.m_slots == 8) return launch_k<false, 8>(p, s);
    if (p.variant != 2 && !p.zero_tz && p.rs.m_slots == 16) return launch_k<false, 16>(p, s);
    return launch_k<false, 0>(p, s);
}

} // namespace kfpos
