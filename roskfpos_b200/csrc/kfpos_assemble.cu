// kfpos_assemble.cu -- the "epoch assembler" (SURVEY.md §8f-1): turns raw time-sorted UWB ranging
// logs into the SoA epoch tensors the replay kernels stream, the batched form of PosGenerator's
// ranging aggregation (Posgenerator.cpp:143-281, 476-507): rangings are grouped by their sequence
// number in a 256-row table; a row is sent when the next sequence number starts or when the
// one-shot 50 ms timer (Posgenerator.h:77) expires after the last ranging.
//
// One thread per log, the log index fastest in every tensor, so the 32 lanes of a warp read 32
// consecutive bytes/words per message field and write 32 consecutive words per output slot.  The
// work is byte/integer shuffling: the kernel is HBM-bound (roofline = measured copy bandwidth).
// The 256-row table lives in global memory in the same layout ([row][slot][log]); with the row
// clearing FIXED a single row is enough.
#include "kfpos_kernels.cuh"

namespace kfpos {

#ifndef ASM_D
#define ASM_D 4
#endif
#ifndef ASM_MINB
#define ASM_MINB 1
#endif
__global__ void __launch_bounds__(128, ASM_MINB) assemble_kernel(const AssembleParams p) {
    const int64_t f = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (f >= p.N) return;
    const int64_t N = p.N;
    const int M = p.M;
    int32_t *tr = p.tbl_r + f;
    double *te = p.tbl_e + f;
    int range_seq = -1; // initialiseTagList (Posgenerator.cpp:499-507); the table is pre-set to -1 / 0
    int64_t n_ep = 0;
    bool have_last = false, armed = false, flushed_once = false;
    double t_last = 0.0, t_prev = 0.0;

    auto flush = [&](double tnow) { // sendRangingMeasurementIfAvailable (:155-198)
        if (range_seq == -1) return;
        if (n_ep < p.max_epochs) {
            const int64_t row = (int64_t)(p.fix_b12 ? 0 : range_seq) * M;
            for (int a = 0; a < M; ++a) {
                p.ranges_out[(n_ep * M + a) * N + f] = tr[(row + a) * N];
                if (p.err_out) p.err_out[(n_ep * M + a) * N + f] = te[(row + a) * N];
            }
            p.dt_out[n_ep * N + f] = flushed_once ? tnow - t_prev : p.first_dt;
            if (p.t_out) p.t_out[n_ep * N + f] = tnow; // the time newTOAMeasurement is called (stream merger)
        }
        n_ep += 1;
        t_prev = tnow;
        flushed_once = true;
        armed = false; // timerRanging.stop() (:174)
    };

    // software pipeline: the fields of the NEXT D messages are in flight while D messages are processed
    // (one message ahead left the kernel latency-bound at 55 % of the copy bandwidth)
    constexpr int D = ASM_D;
    int a_n[D], s_n[D];
    int32_t r_n[D];
    double e_n[D], t_n[D];
    auto load = [&](int64_t i0) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int64_t i = i0 + d;
            const bool in = i < p.L;
            a_n[d] = in ? p.anchor[i * N + f] : 0xff;
            s_n[d] = in ? p.seq[i * N + f] : 0;
            r_n[d] = in ? p.range_mm[i * N + f] : 0;
            e_n[d] = (in && p.err) ? p.err[i * N + f] : 0.0;
            t_n[d] = in ? p.t[i * N + f] : 0.0;
        }
    };
    auto message = [&](int a, int s, int32_t r, double e, double ti) {
        if (a == 0xff || a >= M) return; // padding of a ragged log
        // the one-shot timer fires 0.05 s after the last ranging if nothing arrived before (:143-152)
        if (have_last && armed && ti - t_last > 0.05) flush(t_last + 0.05);
        if (range_seq == s) { // a ranging of the current sequence number (:229-239)
            const int64_t row = (int64_t)(p.fix_b12 ? 0 : s) * M;
            tr[(row + a) * N] = r;
            if (e > 0.0) te[(row + a) * N] = e; // withErrorEstimation (:94,236)
        } else { // a new sequence number: the previous report is sent first (:240-267)
            flush(ti);
            const int64_t row = (int64_t)(p.fix_b12 ? 0 : s) * M;
            if (p.fix_b12) {
                for (int k = 0; k < M; ++k) {
                    tr[(row + k) * N] = -1;
                    te[(row + k) * N] = 0.0;
                }
            } else { // as written: slot 0 only, 64 times (:251-255)
                tr[row * N] = -1;
                te[row * N] = 0.0;
            }
            range_seq = s;
            tr[(row + a) * N] = r;
            te[(row + a) * N] = e;
        }
        t_last = ti; // timerRanging.stop(); timerRanging.start() (:270-273)
        have_last = true;
        armed = true;
    };
    if (p.L > 0) load(0);
    for (int64_t i0 = 0; i0 < p.L; i0 += D) {
        int a_c[D], s_c[D];
        int32_t r_c[D];
        double e_c[D], t_c[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            a_c[d] = a_n[d]; s_c[d] = s_n[d]; r_c[d] = r_n[d]; e_c[d] = e_n[d]; t_c[d] = t_n[d];
        }
        if (i0 + D < p.L) load(i0 + D);
#pragma unroll
        for (int d = 0; d < D; ++d) message(a_c[d], s_c[d], r_c[d], e_c[d], t_c[d]);
    }
    if (have_last && armed) flush(t_last + 0.05); // the timer after the last ranging of the log
    // epochs this log did not produce: no ranging, dt < 0 = "no step" for the replay
    for (int64_t k = n_ep; k < p.max_epochs; ++k) {
        for (int a = 0; a < M; ++a) {
            p.ranges_out[(k * M + a) * N + f] = -1;
            if (p.err_out) p.err_out[(k * M + a) * N + f] = 0.0;
        }
        p.dt_out[k * N + f] = -1.0;
        if (p.t_out) p.t_out[k * N + f] = -1.0;
    }
    if (p.n_epochs) p.n_epochs[f] = (int32_t)(n_ep > 0x7fffffff ? 0x7fffffff : n_ep);
}

cudaError_t launch_assemble(const AssembleParams &p, cudaStream_t s) {
    if (p.N <= 0) return cudaSuccess;
    assemble_kernel<<<(unsigned)((p.N + 127) / 128), 128, 0, s>>>(p);
    return cudaGetLastError();
}

// ---- stream merger: the arrival-order interleaving of one tag's ranging reports and sensor samples
// (PosGenerator's single callback thread, Posgenerator.cpp:92-140), N tags at once, laid onto ONE schedule of
// slots for the ragged replay (kfpos_batch_replay_events_ragged).  One thread per tag does a 5-way merge by
// time stamp over its streams (heads in registers, SoA loads and stores with the tag index fastest): its next
// event takes the next slot of its kind, dt = time since its previous event (first_dt for the first one,
// KF.cpp:232-243), slots passed over stay "no event" (dt = -1).  Byte / integer shuffling: HBM-bound.
__global__ void __launch_bounds__(128) merge_kernel(const MergeParams p) {
    const int64_t f = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (f >= p.N) return;
    const int64_t N = p.N;
    const int M = p.M;
    int64_t ptr[5] = {0, 0, 0, 0, 0};
    double head[5];
    auto load_head = [&](int q) { // negative / NaN = the stream has ended
        head[q] = (p.t_src[q] && ptr[q] < p.L[q]) ? p.t_src[q][ptr[q] * N + f] : -1.0;
    };
#pragma unroll
    for (int q = 0; q < 5; ++q) load_head(q);
    int slot = 0, next_unmarked = 0, dropped = 0;
    bool first = true;
    double t_prev = 0.0;
    for (;;) {
        int k = -1;
        double tk = 0.0;
#pragma unroll
        for (int qq = 1; qq <= 5; ++qq) { // equal time stamps: sensor samples in kind order, then the report
            const int q = qq % 5;
            const double tq = head[q];
            if (!(tq >= 0.0)) continue;
            if (k < 0 || tq < tk) { k = q; tk = tq; }
        }
        if (k < 0) break;
        int s = slot;
        while (s < p.n_slots && p.slot_kind[s] != k) ++s;
        const int64_t j = ptr[k];
        // advance stream k (a select chain keeps ptr / head in registers)
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (q == k) {
                ptr[q] += 1;
                load_head(q);
            }
        if (s >= p.n_slots) { dropped += 1; continue; }
        for (; next_unmarked < s; ++next_unmarked) p.dt_f[(int64_t)next_unmarked * N + f] = -1.0;
        p.dt_f[(int64_t)s * N + f] = first ? p.first_dt : tk - t_prev;
        next_unmarked = s + 1;
        first = false;
        t_prev = tk;
        const int64_t row = p.slot_row[s];
        if (k == 0) {
            for (int a = 0; a < M; ++a) {
                p.ranges_out[(row + a) * N + f] = p.ranges[(j * M + a) * N + f];
                if (p.err_out) p.err_out[(row + a) * N + f] = p.err_src ? p.err_src[(j * M + a) * N + f] : 0.0;
            }
        } else {
            const int rows = k == 1 ? 5 : (k == 2 ? 3 : (k == 3 ? 2 : 1));
            for (int r = 0; r < rows; ++r) p.sensors_out[(row + r) * N + f] = p.src[k][(j * rows + r) * N + f];
        }
        slot = s + 1;
    }
    for (; next_unmarked < p.n_slots; ++next_unmarked) p.dt_f[(int64_t)next_unmarked * N + f] = -1.0;
    if (p.n_dropped) p.n_dropped[f] = dropped;
}

cudaError_t launch_merge(const MergeParams &p, cudaStream_t s) {
    if (p.N <= 0) return cudaSuccess;
    merge_kernel<<<(unsigned)((p.N + 127) / 128), 128, 0, s>>>(p);
    return cudaGetLastError();
}

} // namespace kfpos
