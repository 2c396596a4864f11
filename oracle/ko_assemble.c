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
                    double *dt_out) {
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
                       int32_t *n_epochs) {
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < N; ++f) {
        for (int64_t k = 0; k < max_epochs; ++k) {
            for (int a = 0; a < M; ++a) {
                ranges_out[(k * M + a) * N + f] = -1;
                if (err_out) err_out[(k * M + a) * N + f] = 0.0;
            }
            dt_out[k * N + f] = -1.0;
        }
        n_epochs[f] = (int32_t)ko_assemble(L, M, N, anchor + f, seq + f, range_mm + f, err ? err + f : 0, t + f,
                                           max_epochs, fix_b12, first_dt, N, ranges_out + f,
                                           err_out ? err_out + f : 0, dt_out + f);
    }
}
