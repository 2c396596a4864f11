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
//
// Cooperative advance.  Most parked epochs run to the 10000-iteration cap, one dependent Newton
// iteration after the other: ~950 instructions per iteration issued by a single warp, 18 ms
// whatever the batch size.  Between the main launch and the resume launch a third kernel therefore
// ADVANCES every parked epoch with a GROUP of lanes per epoch (16 or 32 lanes, one per anchor, when
// the queue is short; 4 or 8 lanes with four anchors each when it is long): each lane forms its
// anchors' terms of the cost, gradient and Hessian, the group sums them through shared memory,
// every lane of the group solves the same 3x3 system.  It stops where the epoch's own stop test
// is met (or at the cap) and leaves (point, cost, iteration count) in the record; the resume
// launch then needs one pass to finish the solve.  The sums are associated differently here and
// in the one-thread solver: parked epochs (1.5 in a thousand) see a rounding-level different --
// equally valid -- iteration sequence.  With variant IGNORE_N the re-solve on the reduced set can
// be slow as well: the first resume launch parks those into the second queue, which gets the same
// treatment.  Measured (16 anchors, variant 1): 1 Mi epochs 19.8 -> 11.9 ms, 4 Mi 24.4 -> 21.0 ms.
#include <algorithm>

#include "kfpos_kernels.cuh"
#include "kfpos_solve.cuh"

namespace kfpos {

constexpr int ML_BLOCK = 128;
// Newton iterations an epoch gets in the main launch before it is parked.  Measured (16 anchors, variant 1,
// 1 Mi / 4 Mi epochs): 32 -> 10.4 / 21.5 ms, 64 -> 10.5 / 17.8, 80 -> 9.6 / 18.1, 128 -> 9.9 / 18.9, 256 -> 10.6 / 20.8:
// about half of the epochs that pass 32 iterations end before 80, and every record less shortens the advance
// of a long queue, which is bound by throughput as much as by latency.
#ifndef ML_FIRST_CAP_N
#define ML_FIRST_CAP_N 80
#endif
constexpr unsigned ML_FIRST_CAP = ML_FIRST_CAP_N;
// queue length up to which the one-lane-per-anchor form is used (measured on 16 anchors: 6.4 ms flat up
// to ~2700 records, then 2.3 us per record; the 4-lanes-per-record form: 11 ms at 7800 records)
#ifndef KF_COOP_WIDE_MAX
#define KF_COOP_WIDE_MAX 6000
#endif
constexpr int COOP_WIDE_MAX = KF_COOP_WIDE_MAX;

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

// Group sums through the warp's shared-memory tile.  A warp holds 32 / L groups of L lanes; lane
// `lane` contributes K values; afterwards every lane holds the K sums over ITS group.  Lane gl of
// a group adds up rows gl, gl + L, ... of its group's L columns (16-byte loads), the totals go back
// through the tile.  (Shuffles cost more here: in this data-dependent loop every *_sync shuffle
// is wrapped in a WARPSYNC.COLLECTIVE sequence, ~20 cycles each.)  Two tiles alternate between
// iterations so that one __syncwarp per hand-over is enough.
constexpr int COOP_STRIDE = 34;                       // doubles per row: rows 16-byte aligned, four banks apart
constexpr int COOP_TOT = 12 * COOP_STRIDE;            // totals: [group][12]
constexpr int COOP_TILE = COOP_TOT + 16 * 12;         // doubles per tile (up to 16 groups per warp)
template <int K, int L>
KF_DEV void group_sum(double *tile, int lane, double (&v)[K]) {
    const int grp = lane / L, gl = lane % L;
#pragma unroll
    for (int j = 0; j < K; ++j) tile[j * COOP_STRIDE + lane] = v[j];
    __syncwarp();
#pragma unroll
    for (int j0 = 0; j0 < K; j0 += L) {
        const int j = j0 + gl;
        if (j < K) {
            const double2 *row = reinterpret_cast<const double2 *>(tile + j * COOP_STRIDE + grp * L);
            if (L == 2) {
                const double2 lo = row[0];
                tile[COOP_TOT + grp * 12 + j] = lo.x + lo.y;
            } else {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int i = 0; i < L / 4; ++i) {
                    const double2 lo = row[2 * i], hi = row[2 * i + 1];
                    a0 += lo.x; a1 += lo.y; a2 += hi.x; a3 += hi.y;
                }
                tile[COOP_TOT + grp * 12 + j] = (a0 + a1) + (a2 + a3);
            }
        }
    }
    __syncwarp();
    const double2 *tot = reinterpret_cast<const double2 *>(tile + COOP_TOT + grp * 12);
#pragma unroll
    for (int j = 0; j < K / 2; ++j) {
        const double2 t = tot[j];
        v[2 * j] = t.x; v[2 * j + 1] = t.y;
    }
    if (K & 1) v[K - 1] = tile[COOP_TOT + grp * 12 + K - 1];
}

// L lanes per parked record (32 / L records per warp), A anchor slots per lane (slot = gl + L a):
// advances each record's 3-D Newton solve (the loop of ml_solve3, ML.cpp:165-222) to its stop
// test or the iteration cap.  Warps stride over the queue; a warp leaves when all its records are done.
template <bool PME, int L, int A>
__global__ void __launch_bounds__(ML_BLOCK) ml_coop_kernel(const __grid_constant__ MlParams p) {
    __shared__ __align__(16) double tiles[ML_BLOCK / 32][2 * COOP_TILE];
    constexpr int G = 32 / L;
    const int lane = threadIdx.x & 31, grp = lane / L, gl = lane % L;
    double *tile = tiles[threadIdx.x >> 5];
    const int count = min(*p.q_in_count, p.queue_cap);
    // two instantiations are launched back to back and the queue length picks one: many lanes per
    // record give the shortest iteration (few records: latency-bound), few lanes per record the
    // least work per record (many records: the wide form would be throughput-bound)
    if (count < p.coop_min || count >= p.coop_max) return;
    const int n_warps = (int)(gridDim.x * (ML_BLOCK / 32));
    const int64_t N = p.N;
    const int m = p.rs.m_slots;
    const double inv_e0 = fast_rcp(p.rs.err_scalar);
    for (int wbase = (int)(blockIdx.x * (ML_BLOCK / 32) + (threadIdx.x >> 5)) * G; wbase < count; wbase += n_warps * G) {
        const bool live = wbase + grp < count;
        MlParked *rec = reinterpret_cast<MlParked *>(const_cast<void *>(p.q_in)) + (live ? wbase + grp : wbase);
        const int64_t f = rec->idx;
        const unsigned used = live ? rec->used : 0u;
        double ax[A], ay[A], az[A], r[A], w[A];
        bool on[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const int slot = gl + L * a;
            on[a] = slot < m && ((used >> slot) & 1u);
            const int li = slot < m ? slot : 0;
            ax[a] = p.anchors.x[li]; ay[a] = p.anchors.y[li]; az[a] = p.anchors.z[li];
            r[a] = 0.0; w[a] = 1.0;
            if (on[a]) {
                const int64_t at = (int64_t)li * N + f;
                if (p.rs.fmt == 0) r[a] = reinterpret_cast<const double *>(p.rs.ranges)[at];
                else if (p.rs.fmt == 1) r[a] = mm_to_m((double)reinterpret_cast<const int32_t *>(p.rs.ranges)[at]);
                else r[a] = mm_to_m((double)reinterpret_cast<const uint16_t *>(p.rs.ranges)[at]);
                if (PME) w[a] = fast_rcp(p.rs.err[at]);
            }
        }
        const double nvalid = (double)__popc(used);
        double px = rec->p[0], py = rec->p[1], pz = rec->p[2], cost = rec->cost;
        unsigned iter = rec->iter;
        bool done = !live;
        for (unsigned round = 0;; ++round) {
            constexpr int K = PME ? 11 : 10;
            double v[K];
#pragma unroll
            for (int j = 0; j < K; ++j) v[j] = 0.0;
#pragma unroll
            for (int a = 0; a < A; ++a) {
                const double dx = ax[a] - px, dy = ay[a] - py, dz = az[a] - pz;
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const double invd = on[a] ? fast_rsqrt(d2) : 0.0;
                const double res = fma(-d2, invd, r[a]);
                const double rid = r[a] * invd;
                const double t = res * invd * w[a];
                const double c2 = rid * invd * invd * w[a];
                const double cx = c2 * dx, cy = c2 * dy;
                v[0] = fma(res * res, w[a], v[0]);
                v[1] = fma(t, dx, v[1]); v[2] = fma(t, dy, v[2]); v[3] = fma(t, dz, v[3]);
                v[4] += PME ? (on[a] ? (1.0 - rid) * w[a] : 0.0) : rid;
                v[5] = fma(cx, dx, v[5]); v[6] = fma(cx, dy, v[6]); v[7] = fma(cy, dy, v[7]);
                v[8] = fma(cx, dz, v[8]); v[9] = fma(cy, dz, v[9]);
                if (PME) v[K - 1] = fma(c2 * dz, dz, v[K - 1]);
            }
            group_sum<K, L>(tile + (round & 1u) * COOP_TILE, lane, v);
            {
                // straight-line: every lane evaluates the stop test and the 3x3 solve (same operations and order
                // as rel_change_gt / solve_sym3), the results are committed with selects -- the iteration is one
                // serial dependency chain, and every divergent region on it cost a reconvergence barrier
                const double wcost = v[0];
                double c1s = v[4], h5;
                if (PME) {
                    h5 = v[K - 1];
                } else { // unit direction vectors: see ml_pass3
                    h5 = (c1s - v[5]) - v[7];
                    c1s = nvalid - c1s;
                }
                const double newCost = PME ? wcost : wcost * inv_e0;
                const double q = fabs(cost - newCost), thr = 1e-3 * cost;
                bool changed = q > thr * 1.000000001;
                if (!(cost > 0.0) || !(changed || q < thr * 0.999999999)) changed = q / cost > 1e-3; // within 1e-9 of the threshold
                const double Ha = v[5] + c1s, Hb = v[6], Hd = v[7] + c1s, Hc = v[8], He = v[9], Hf = h5 + c1s;
                const double c00 = Hd * Hf - He * He, c01 = Hc * He - Hb * Hf, c02 = Hb * He - Hc * Hd;
                const double det = Ha * c00 + Hb * c01 + Hc * c02;
                const double c11 = Ha * Hf - Hc * Hc, c12 = Hb * Hc - Ha * He, c22 = Ha * Hd - Hb * Hb;
                const double id = fast_rcp(det);
                const double s0 = (c00 * v[1] + c01 * v[2] + c02 * v[3]) * id;
                const double s1 = (c01 * v[1] + c11 * v[2] + c12 * v[3]) * id;
                const double s2 = (c02 * v[1] + c12 * v[2] + c22 * v[3]) * id;
                // a singular system is left to the one-thread solver, which reports it
                const bool step = !done && changed && iter < 10000u && usable_det(det);
                done = !step;
                iter += step ? 1u : 0u;
                cost = step ? newCost : cost;
                px = step ? px - s0 : px; py = step ? py - s1 : py; pz = step ? pz - s2 : pz;
            }
            // the warp vote costs ~100 cycles of a ~1200-cycle iteration: taken every fourth iteration (a finished
            // record just idles in between; nothing it holds changes)
            if ((round & 3u) == 3u && __all_sync(0xffffffffu, done)) break;
        }
        if (gl == 0 && live) {
            rec->p[0] = px; rec->p[1] = py; rec->p[2] = pz;
            rec->cost = cost;
            rec->iter = iter;
        }
        __syncwarp();
    }
}

// RESUME = false: thread f owns epoch f;  true: thread q owns parked record q.
template <bool PME, int MT, bool RESUME>
__global__ void __launch_bounds__(ML_BLOCK) ml_solve_kernel(const __grid_constant__ MlParams p) {
    extern __shared__ double smem[];
    const int64_t t = (int64_t)blockIdx.x * ML_BLOCK + threadIdx.x;
    const MlParked *queue_in = reinterpret_cast<const MlParked *>(p.q_in);
    MlParked *queue = reinterpret_cast<MlParked *>(p.q_out);
    const bool active = RESUME ? t < min(*p.q_in_count, p.queue_cap) : t < p.N;
    unsigned iters = 0, bad = 0, done = 0;
    if (active) {
        const int64_t N = p.N;
        const int m = MT > 0 ? MT : p.rs.m_slots;
        MlParked rec;
        if (RESUME) rec = queue_in[t];
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
        unsigned cap = p.variant == 2 ? 10000u : p.first_cap; // BEST does not park
        int rc;
        bool parked = false, near_tie = false;
        for (;;) {
            const bool last = !(p.variant == 1 && phase == 0);
            rc = ml_any<PME, MT>(p.anchors, ep, used, use2d, p.zero_tz != 0, pos, last ? cov : nullptr, sse, iters, cap,
                                 &rs);
            if (rc == ML_MORE) {
                const int slot = queue ? atomicAdd(p.q_out_count, 1) : p.queue_cap;
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
                used = drop_worst<PME, MT>(p.anchors, ep, used, pos, drop, p.xq ? &near_tie : nullptr);
                if (near_tie) { // too close to call with this arithmetic: the exact-order solver decides
                    const int slot = atomicAdd(p.xq_count, 1);
                    if (slot < p.xq_cap) {
                        p.xq[slot] = (int32_t)f;
                        parked = true; // nothing is written or counted here
                        break;
                    }
                }
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

// ---- NORMAL-mode 3-D epochs as a stream (BASELINE config 2).  In ml_solve_kernel a thread owns one epoch and
// a warp lasts as long as its slowest lane: the Newton trip counts from the fixed start point spread between 5
// and 12 (mean 7.7 at 8 anchors), and ncu shows 20 of 32 lanes active on average.  Here warps are persistent and
// every LANE is a state machine: one Newton pass per trip of a warp-uniform loop; a lane whose solve has ended
// writes its epoch (covariance, outputs) and takes the next one from a global counter when ML_STREAM_FIN lanes
// wait or nobody iterates, its raw rangings already in registers (fetched one epoch ahead).  Per epoch the
// arithmetic is ml_solve3's, operation for operation: outputs are bit-identical to ml_solve_kernel's.  An epoch
// that needs more than `first_cap` iterations is parked in the straggler queue exactly as there.
#ifndef ML_STREAM_FIN_N
#define ML_STREAM_FIN_N 24
#endif
constexpr int ML_STREAM_FIN = ML_STREAM_FIN_N;
template <int MT>
__global__ void __launch_bounds__(ML_BLOCK, 4) ml_stream3_kernel(const __grid_constant__ MlParams p, int *next_epoch) {
    const int64_t N = p.N;
    const int fmt = p.rs.fmt; // 1: int32 mm, 2: uint16 mm
    MlParked *queue = reinterpret_cast<MlParked *>(p.q_out);
    enum { FETCH = 0, RUN = 1, FIN = 2, DONE = 3 };
    int phase = FETCH;
    int64_t f = -1, f_next = -1;
    unsigned raw_next[MT];
    EpochT<false, MT> ep;
    ep.e0 = p.rs.err_scalar;
    ep.m_slots = MT;
    ep.valid = 0u;
    auto load_raw = [&](int64_t fe) {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int64_t at = (int64_t)i * N + fe;
            raw_next[i] = fmt == 1 ? (unsigned)__ldg(reinterpret_cast<const int32_t *>(p.rs.ranges) + at)
                                   : (unsigned)__ldg(reinterpret_cast<const uint16_t *>(p.rs.ranges) + at);
        }
    };
    {
        const int64_t t = atomicAdd(next_epoch, 1);
        if (t < N) { f_next = t; load_raw(t); }
    }
    double pos[3] = {0, 0, 0}, cost = 1e20;
    unsigned iter = 0, nvalid = 0, iters_sum = 0, n_done = 0, n_bad = 0;
    bool first = true;
    int rc = ML_OK;
    MlPass3 ps;
    for (;;) {
        const unsigned waiting = __ballot_sync(0xffffffffu, phase == FIN || phase == FETCH);
        const unsigned running = __ballot_sync(0xffffffffu, phase == RUN);
        if (waiting == 0u && running == 0u) break;
        if (waiting != 0u && (running == 0u || __popc(waiting) >= ML_STREAM_FIN)) {
            if (phase == FIN) {
                double cv[6] = {0, 0, 0, 0, 0, 0};
                bool parked = false, have_cov = false;
                if (rc == ML_MORE) { // park; a full queue: finish in place
                    const int slot = queue ? atomicAdd(p.q_out_count, 1) : p.queue_cap;
                    if (slot < p.queue_cap) {
                        MlParked rec;
                        rec.idx = (int32_t)f; rec.phase = 0; rec.used = ep.valid; rec.iter = iter; rec.iters = 0u;
                        rec._pad = 0u; rec.cost = cost; rec.p[0] = pos[0]; rec.p[1] = pos[1]; rec.p[2] = pos[2];
                        rec._pad2 = 0.0;
                        queue[slot] = rec;
                        parked = true;
                    } else {
                        MlResume rs = {cost, iter};
                        double sse_r;
                        unsigned it_r = 0;
                        rc = ml_solve3<false, MT>(p.anchors, ep, ep.valid, pos, sse_r, it_r, cv, nullptr, 10000u, &rs);
                        iter = it_r;
                        have_cov = true;
                    }
                }
                if (!parked) {
                    // estimatePosition ends with the covariance (ML.cpp:229-254); a singular J^T W^-1 J fails the epoch
                    if (rc == ML_OK && !have_cov) rc = ml_cov3<false, MT>(p.anchors, ep, ep.valid, pos, ps.sse, cv);
                    if (p.cov) {
                        const bool ok = rc == ML_OK;
                        const double c00 = ok ? cv[0] : 0.0, c01 = ok ? cv[1] : 0.0, c11 = ok ? cv[2] : 0.0;
                        const double c02 = ok ? cv[3] : 0.0, c12 = ok ? cv[4] : 0.0, c22 = ok ? cv[5] : 0.0;
                        p.cov[0 * N + f] = c00; p.cov[1 * N + f] = c01; p.cov[2 * N + f] = c02;
                        p.cov[3 * N + f] = c01; p.cov[4 * N + f] = c11; p.cov[5 * N + f] = c12;
                        p.cov[6 * N + f] = c02; p.cov[7 * N + f] = c12; p.cov[8 * N + f] = c22;
                    }
                    if (p.pos) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) p.pos[(int64_t)q * N + f] = pos[q];
                    }
                    if (p.iters) p.iters[f] = (int32_t)iter;
                    if (p.sel) {
                        p.sel[f] = (int32_t)ep.valid;
                        p.sel[N + f] = -1;
                    }
                    int stv = rc == ML_OK ? 0 : (rc == ML_FEW ? 2 : 4);
                    if (p.max_z > p.min_z && (pos[2] < p.min_z || pos[2] > p.max_z)) stv |= 128;
                    if (p.status) p.status[f] = stv;
                    n_bad += (stv & ~128) != 0;
                    n_done += 1;
                    iters_sum += iter;
                }
                phase = FETCH;
            }
            if (phase == FETCH) {
                if (f_next < 0) {
                    phase = DONE;
                } else {
                    f = f_next;
                    unsigned valid = 0u;
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        const double r = mm_to_m(fmt == 1 ? (double)(int)raw_next[i] : (double)raw_next[i]);
                        ep.zr[i] = r;
                        if (r > 0) valid |= 1u << i;
                    }
                    ep.valid = valid;
                    nvalid = __popc(valid);
                    const int64_t t = atomicAdd(next_epoch, 1);
                    f_next = -1;
                    if (t < N) { f_next = t; load_raw(t); }
                    pos[0] = p.start[0]; pos[1] = p.start[1]; pos[2] = p.start[2];
                    cost = 1e20; iter = 0; first = true; rc = ML_OK;
                    phase = RUN;
                }
            }
        }
        if (phase == RUN) { // one trip of ml_solve3's loop
            ml_pass3<false, MT>(p.anchors, ep, ep.valid, (int)nvalid, pos, ps);
            double newCost = 1.0;
            bool stop = false;
            if (first) {
                first = false;
                if (nvalid < 4) { rc = ML_FEW; stop = true; }
                else if (ep.e0 == 0.0) { rc = ML_SINGULAR; stop = true; }
            } else {
                newCost = ps.wcost;
            }
            if (!stop) {
                if (!(rel_change_gt(cost, newCost) && iter < 10000u)) {
                    stop = true;
                } else if (iter >= p.first_cap) {
                    rc = ML_MORE;
                    stop = true;
                } else {
                    iter += 1;
                    cost = newCost;
                    double sv[3];
                    if (!solve_sym3(ps.H, ps.g, sv)) {
                        rc = ML_SINGULAR;
                        stop = true;
                    } else {
                        pos[0] -= sv[0]; pos[1] -= sv[1]; pos[2] -= sv[2];
                    }
                }
            }
            if (stop) phase = FIN;
        }
    }
    warp_accumulate(p.counters + CNT_UPDATES, n_done);
    warp_accumulate(p.counters + CNT_ML_ITERS, iters_sum);
    warp_accumulate(p.counters + CNT_BAD, n_bad);
}

template <bool PME, int MT>
static cudaError_t launch_k(const MlParams &p0, cudaStream_t s) {
    MlParams p = p0;
    const int m = p.rs.m_slots;
    const size_t smem = (size_t)((MT > 0 ? 0 : m) + (PME ? m : 0) + raw_rows(p.rs.fmt, m)) * ML_BLOCK * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(ml_solve_kernel<PME, MT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ml_solve_kernel<PME, MT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(p.queue_count, 0, 2 * sizeof(int), s);
    if (e != cudaSuccess) return e;
    const bool parks = !p.use2d && p.variant != 2;
    p.q_in = nullptr; p.q_in_count = nullptr;
    p.q_out = parks ? p.queue[0] : nullptr; p.q_out_count = p.queue_count;
    p.first_cap = parks ? ML_FIRST_CAP : 10000u;
    bool streamed = false;
    if constexpr (!PME && MT > 0) {
        // NORMAL mode, 3-D, integer wire formats: the persistent per-lane state machines (ml_stream3_kernel)
#ifndef ML_NO_STREAM
        if (p.variant == 0 && !p.use2d && p.rs.fmt != 0 && p.stream_counter != nullptr && p.N < 0x7fffffff) {
            e = cudaMemsetAsync(p.stream_counter, 0, sizeof(int), s);
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int64_t want = (p.N + ML_BLOCK - 1) / ML_BLOCK;
            const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * 4);
            ml_stream3_kernel<MT><<<grid, ML_BLOCK, 0, s>>>(p, p.stream_counter);
            streamed = true;
        }
#endif
    }
    if (!streamed) ml_solve_kernel<PME, MT, false><<<(unsigned)((p.N + ML_BLOCK - 1) / ML_BLOCK), ML_BLOCK, smem, s>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess || !parks) return e;
    // the parked epochs: grids sized for the whole queue (warps / blocks beyond the count exit at once)
    const unsigned coop_grid = (unsigned)std::min<int64_t>(((int64_t)p.queue_cap * 32 + ML_BLOCK - 1) / ML_BLOCK, 148 * 8);
    const unsigned res_grid = (unsigned)((p.queue_cap + ML_BLOCK - 1) / ML_BLOCK);
    const int rounds = p.variant == 1 ? 2 : 1; // IGNORE_N solves twice
    for (int q = 0; q < rounds; ++q) {
        p.q_in = p.queue[q]; p.q_in_count = p.queue_count + q;
#ifdef KF_COOP_FORCE_L // profiling builds: one form for every queue length (0 = no cooperative advance)
        p.coop_min = 0; p.coop_max = 0x7fffffff;
#if KF_COOP_FORCE_L > 0
        ml_coop_kernel<PME, KF_COOP_FORCE_L, 16 / KF_COOP_FORCE_L><<<coop_grid, ML_BLOCK, 0, s>>>(p);
#endif
        if (false) {
#else
        if (m <= 16) {
#endif
            p.coop_min = 0; p.coop_max = COOP_WIDE_MAX;
            // 8 lanes x 2 anchors measured 8 % faster than 16 x 1 on the 16-anchor batch (12.4 -> 11.5 ms at 1 Mi
            // epochs; 4 x 4: 15.2 ms): the iteration is a serial chain, a shorter group sum buys more than
            // the second anchor per lane costs
            ml_coop_kernel<PME, 8, 2><<<coop_grid, ML_BLOCK, 0, s>>>(p);
            p.coop_min = COOP_WIDE_MAX; p.coop_max = 0x7fffffff;
            ml_coop_kernel<PME, 4, 4><<<coop_grid, ML_BLOCK, 0, s>>>(p);
        } else {
            p.coop_min = 0; p.coop_max = COOP_WIDE_MAX;
            ml_coop_kernel<PME, 32, 1><<<coop_grid, ML_BLOCK, 0, s>>>(p);
            p.coop_min = COOP_WIDE_MAX; p.coop_max = 0x7fffffff;
            ml_coop_kernel<PME, 8, 4><<<coop_grid, ML_BLOCK, 0, s>>>(p);
        }
        const bool last = q + 1 == rounds;
        p.q_out = last ? nullptr : p.queue[q + 1]; p.q_out_count = p.queue_count + q + 1;
        p.first_cap = last ? 10000u : ML_FIRST_CAP;
        ml_solve_kernel<PME, MT, true><<<res_grid, ML_BLOCK, smem, s>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static cudaError_t launch_fast(const MlParams &p, cudaStream_t s) {
    if (p.rs.err != nullptr) return launch_k<true, 0>(p, s);
    // BEST solves k-anchor subsets: the branch-skipping rolled loops do k, not m, anchors of work
    if (p.variant != 2 && !p.zero_tz && p.rs.m_slots == 8) return launch_k<false, 8>(p, s);
    if (p.variant != 2 && !p.zero_tz && p.rs.m_slots == 16) return launch_k<false, 16>(p, s);
    return launch_k<false, 0>(p, s);
}

// kfpos_config.ml_exact_order: 1 = every epoch in exact order; 0 = BestGroup in exact order, IgnoreN with
// the fast solver + an exact re-decision of the near-ties, NORMAL with the fast solver; -1 = fast only
cudaError_t launch_ml_solve(const MlParams &p0, cudaStream_t s) {
    if (p0.N <= 0) return cudaSuccess;
    MlParams p = p0;
    if (p.exact_mode > 0 || (p.exact_mode == 0 && p.variant == 2)) return launch_ml_exact(p, false, s);
    const bool recheck = p.exact_mode == 0 && p.variant == 1 && p.xq != nullptr && p.xq_cap > 0;
    if (!recheck) {
        p.xq = nullptr;
        return launch_fast(p, s);
    }
    cudaError_t e = cudaMemsetAsync(p.xq_count, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    e = launch_fast(p, s);
    if (e != cudaSuccess) return e;
    return launch_ml_exact(p, true, s);
}

} // namespace kfpos
