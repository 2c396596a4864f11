/*
 * ko_assemble.c -- CPU ORACLE (test infrastructure, NOT product code): the ranging
 * aggregation of PosGenerator, restated for offline logs.
 *
 * Reference: /root/reference/src/kfpos/publishers/Posgenerator.cpp
 *   processRangingNow                    :201-281   (the live path of newTOAMeasurement :92-97)
 *   sendRangingMeasurementIfAvailable    :155-198
 *   timerRangingCallback                 :143-152   (one-shot timer, MAX_TIME_TO_SEND_RANGING = 0.05 s,
 *                                                    Posgenerator.h:77, re-armed by every ranging :270-273)
 *   initialiseTagList                    :499-507   (table = -1, errorEstimation = 0)
 *   calculateTagLocationWithRangings     :476-496   (a slot is used when its value is > 0)
 *
 * PINNING: pinned against the reference's own source.  Posgenerator.cpp is compiled UNMODIFIED into
 * oracle/_ref/libkfref.so against stand-ins for ROS, tf and the message packages (oracle/shim/ros,
 * kfshim_msgs.h); oracle/shim/posgen_harness.cpp plays ROS's event loop on the fake clock (the
 * one-shot timer fires when it is due before the next ranging arrives) and records every epoch
 * handed to newTOAMeasurement.  Evidence: tests/golden/posgen.npz (tests/golden/make_golden_posgen.py)
 * replayed by tests/test_oracle_golden.py -- epochs bit-equal, timeLag to 1e-12, and the whole
 * log -> report chain -- and tests/test_oracle_vs_ref.py on fresh random logs; plus the hand-computed
 * cases of tests/test_oracle_assemble.py.  The wall clock of the reference (steady_clock at the
 * moment a report is sent) is the arrival time stamp of the log here; gtec_msgs/Ranging itself is
 * an un-vendored dependency (field list restated in kfshim_msgs.h).
 *
 * One call = one tag's time-sorted stream.  A message with anchor == 0xFF (or >= M) is padding.
 * The reference keeps 256 rows indexed by seq and, when a new seq starts, clears only slot 0 of the
 * row 64 times (:251-255): values written 256 sequence numbers earlier survive
 * in the other slots.  fix_b12 = 0 reproduces that; fix_b12 = 1 clears the whole row.
 */
#include <stdlib.h>
#include <string.h>

#include "kfpos_oracle.h"

int64_t ko_assemble(int64_t L, int M, int64_t stride, const uint8_t *anchor, const uint8_t *seq,
                    const int32_t *range_mm, const double *err, const double *t, int64_t max_epochs,
                    int fix_b12, double first_dt, int64_t out_stride, int32_t *ranges_out, double *err_out,
                    double *dt_out, double *t_out /* NULL, or the time of every report */) {
    int32_t(*val)[KO_MAX_ANCHORS] = malloc(sizeof(int32_t) * 256 * KO_MAX_ANCHORS);
    double(*ee)[KO_MAX_ANCHORS] = malloc(sizeof(double) * 256 * KO_MAX_ANCHORS);
    memset(val, 0xff, sizeof(int32_t) * 256 * KO_MAX_ANCHORS); /* :505 */
    memset(ee, 0, sizeof(double) * 256 * KO_MAX_ANCHORS);      /* :501 */
    int range_seq = -1;                                          /* :504 */
    int64_t n_ep = 0;
    int have_last = 0, timer_armed = 0;
    double t_last = 0.0, t_prev_flush = 0.0;
    int flushed_once = 0;

#define FLUSH(tnow)                                                                                  \
    do { /* sendRangingMeasurementIfAvailable: rangeSeq != -1 and rangeCount >= 1 (:162-166) */      \
        if (range_seq != -1) {                                                                       \
            if (n_ep < max_epochs) {                                                                 \
                for (int a = 0; a < M; ++a) {                                                        \
                    ranges_out[(n_ep * M + a) * out_stride] = val[range_seq][a];                     \
                    if (err_out) err_out[(n_ep * M + a) * out_stride] = ee[range_seq][a];            \
                }                                                                                    \
                dt_out[n_ep * out_stride] = flushed_once ? (tnow) - t_prev_flush : first_dt;         \
                if (t_out) t_out[n_ep * out_stride] = (tnow);                                        \
            }                                                                                        \
            n_ep += 1;                                                                               \
            t_prev_flush = (tnow);                                                                   \
            flushed_once = 1;                                                                        \
            timer_armed = 0; /* timerRanging.stop() (:174) */                                        \
        }                                                                                            \
    } while (0)

    for (int64_t i = 0; i < L; ++i) {
        const int a = anchor[i * stride];
        if (a == 0xff || a >= M) continue;
        const int s = seq[i * stride];
        const double ti = t[i * stride];
        /* the one-shot timer fires 0.05 s after the last ranging if nothing arrived before */
        if (have_last && timer_armed && ti - t_last > 0.05) FLUSH(t_last + 0.05);
        const int32_t r = range_mm[i * stride]; /* floor(rawrange) of an integer wire value (:210) */
        const double e = err ? err[i * stride] : 0.0;
        if (range_seq == s) { /* :229-239 */
            val[s][a] = r;
            if (e > 0.0) ee[s][a] = e; /* withErrorEstimation = errorEstimation > 0 (:94) */
        } else { /* :240-267 */
            FLUSH(ti);
            if (fix_b12) {
                for (int k = 0; k < KO_MAX_ANCHORS; ++k) { val[s][k] = -1; ee[s][k] = 0.0; }
            } else {
                val[s][0] = -1; /* as written: slot 0 only (:251-255) */
                ee[s][0] = 0.0;
            }
            range_seq = s;
            val[s][a] = r;
            ee[s][a] = e;
        }
        t_last = ti; /* timer stop + start (:270-273) */
        have_last = 1;
        timer_armed = 1;
    }
    if (have_last && timer_armed) FLUSH(t_last + 0.05); /* the timer after the last ranging */
#undef FLUSH
    free(val);
    free(ee);
    return n_ep;
}

/* N logs, SoA with the log index fastest ([L][N] inputs, [T][M][N] / [T][N] outputs) */
void ko_assemble_batch(int64_t N, int64_t L, int M, const uint8_t *anchor, const uint8_t *seq,
                       const int32_t *range_mm, const double *err, const double *t, int64_t max_epochs,
                       int fix_b12, double first_dt, int32_t *ranges_out, double *err_out, double *dt_out,
                       int32_t *n_epochs, double *t_out /* [max_epochs][N] or NULL */) {
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < N; ++f) {
        for (int64_t k = 0; k < max_epochs; ++k) {
            for (int a = 0; a < M; ++a) {
                ranges_out[(k * M + a) * N + f] = -1;
                if (err_out) err_out[(k * M + a) * N + f] = 0.0;
            }
            dt_out[k * N + f] = -1.0;
            if (t_out) t_out[k * N + f] = -1.0;
        }
        n_epochs[f] = (int32_t)ko_assemble(L, M, N, anchor + f, seq + f, range_mm + f, err ? err + f : 0, t + f,
                                           max_epochs, fix_b12, first_dt, N, ranges_out + f,
                                           err_out ? err_out + f : 0, dt_out + f, t_out ? t_out + f : 0);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Stream merger: what PosGenerator's single callback thread does with one tag's messages when they
 * come from several topics (Posgenerator.cpp:92-140): the ranging reports (ko_assemble, at their report
 * times) and the PX4Flow / IMU / magnetometer / compass samples reach the filter in ARRIVAL ORDER, and
 * the filter takes the time since its previous callback as dt (0.1 s for the first one, KF.cpp:232-243).
 * For a batch, the N per-tag sequences are laid onto ONE schedule of S slots (slot s has kind
 * slot_kind[s]; KO event kinds 0 TOA, 1 PX4, 2 IMU, 3 MAG, 4 COMPASS): a tag's next event takes the next
 * slot of its kind, slots it passes over are marked "no event" (dt = -1) -- the ragged-replay
 * convention of kfpos_batch_replay_events_ragged.  Equal time stamps: sensor samples in kind order,
 * then the ranging report (a timer that is due exactly when a message arrives fires after it).
 *   t_src[k]   : [L_k][N] arrival times of stream k (k = 0: report times from ko_assemble); a value that
 *                is negative or NaN ends the stream
 *   src[k]     : k = 0: int32 ranges [L_0][M][N] (+ err_src [L_0][M][N] or NULL); k > 0: f64 [L_k][rows_k][N]
 *   slot_row[s]: first output row of slot s (rows of ranges_out for TOA slots, of sensors_out otherwise)
 *   outputs    : dt_f [S][N]; ranges_out / err_out [*][N]; sensors_out [*][N]; n_dropped [N] = events that
 *                found no slot left
 */
static const int ko_event_rows[5] = {0, 5, 3, 2, 1};
void ko_merge_batch(int64_t N, int M, const int64_t L[5], const double *const t_src[5], const int32_t *ranges,
                    const double *err_src, const double *const src[5], int S, const int32_t *slot_kind,
                    const int64_t *slot_row, double first_dt, double *dt_f, int32_t *ranges_out, double *err_out,
                    double *sensors_out, int32_t *n_dropped) {
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < N; ++f) {
        int64_t ptr[5] = {0, 0, 0, 0, 0};
        int slot = 0, dropped = 0, first = 1;
        double t_prev = 0.0;
        for (int s = 0; s < S; ++s) dt_f[(int64_t)s * N + f] = -1.0;
        for (;;) {
            int k = -1;
            double tk = 0.0;
            for (int qq = 1; qq <= 5; ++qq) { /* equal time stamps: sensor samples in kind order, then the report */
                const int q = qq % 5;
                if (!t_src[q] || ptr[q] >= L[q]) continue;
                const double tq = t_src[q][ptr[q] * N + f];
                if (!(tq >= 0.0)) continue; /* negative or NaN: the stream has ended */
                if (k < 0 || tq < tk) { k = q; tk = tq; }
            }
            if (k < 0) break;
            int s = slot;
            while (s < S && slot_kind[s] != k) ++s;
            const int64_t j = ptr[k]++;
            if (s >= S) { dropped += 1; continue; }
            dt_f[(int64_t)s * N + f] = first ? first_dt : tk - t_prev;
            first = 0;
            t_prev = tk;
            if (k == 0) {
                for (int a = 0; a < M; ++a) {
                    ranges_out[(slot_row[s] + a) * N + f] = ranges[(j * M + a) * N + f];
                    if (err_out) err_out[(slot_row[s] + a) * N + f] = err_src ? err_src[(j * M + a) * N + f] : 0.0;
                }
            } else {
                for (int r = 0; r < ko_event_rows[k]; ++r)
                    sensors_out[(slot_row[s] + r) * N + f] = src[k][(j * ko_event_rows[k] + r) * N + f];
            }
            slot = s + 1;
        }
        if (n_dropped) n_dropped[f] = dropped;
    }
}
