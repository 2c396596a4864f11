#!/usr/bin/env python
"""Generates tests/golden/posgen.npz from the REFERENCE ITSELF: publishers/Posgenerator.cpp compiled
unmodified into oracle/_ref/libkfref.so (oracle/shim/posgen_harness.cpp plays the ROS event loop on a
fake clock).  For a handful of ranging logs it stores
  * the log (anchor index, seq, range_mm, arrival time, errorEstimation; padded with anchor = 255),
  * every epoch PosGenerator handed to newTOAMeasurement (ranges in metres by beacon index, 0 = slot
    not in the epoch; error estimates; timeLag),
  * the report the node publishes (PoseWithCovarianceStamped + Odometry twist) when polled at
    several lags after the last epoch, with KalmanFilterTOA and with KalmanFilterTOAIMU behind it,
    and with MLLocation (variants 0 and 1) behind it.
Run:  python tests/golden/make_golden_posgen.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_py as R  # noqa: E402
from roskfpos_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LAGS = [0.0, 0.02, 0.1]
M = 8


def make_log(rng, truth, n_seq, seq0, p_drop, p_gap, rate_gap, p_err, junk):
    """One tag's arrival-ordered ranging log on a 1 ms time grid."""
    anc = synth.anchors_for(M)
    a, s, r, t, e = [], [], [], [], []
    tt = 0.0
    for q in range(n_seq):
        if rng.random() < p_gap:
            tt += 0.08  # silence > MAX_TIME_TO_SEND_RANGING inside the stream
        p = truth[q + 1][:, 0]
        for k in rng.permutation(M):
            if rng.random() < p_drop:
                continue
            tt += 0.001 * int(rng.integers(1, 4))
            d = np.linalg.norm(anc[k] - p) + rng.normal(0, 0.05)
            rv = int(d * 1000)
            if junk and rng.random() < 0.05:
                rv = int(rng.integers(-3, 1))  # the device's "no ranging" values
            a.append(k); s.append((seq0 + q) % 256); r.append(rv); t.append(round(tt, 3))
            e.append(float(np.float32(0.01 + rng.random() * 0.05)) if rng.random() < p_err else 0.0)
        tt += rate_gap[q % len(rate_gap)]
    return [np.array(v) for v in (a, s, r, t, e)]


def main():
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    anc = synth.anchors_for(M)
    rng = np.random.default_rng(77)
    cases = [  # n_seq, seq0, p_drop, p_gap, rate_gap, p_err, junk
        (40, 0, 0.0, 0.0, (0.03,), 1.0, False),         # clean 20 Hz stream, every sequence complete
        (60, 250, 0.15, 0.1, (0.03, 0.07), 1.0, False),  # drops, silences, every other epoch sent twice
        (300, 200, 0.3, 0.05, (0.02,), 1.0, True),       # > 256 sequences: stale slots of the table (App. B-12)
        (25, 7, 0.5, 0.3, (0.01,), 1.0, True),           # sparse
        (50, 100, 0.1, 0.0, (0.03,), 0.7, False),        # some rangings without an error estimate
        (1, 9, 0.0, 0.0, (0.03,), 1.0, False),           # a single sequence: only the timer sends it
    ]
    logs, L = [], 0
    for c in cases:
        truth = synth.truth_lissajous(1, c[0], 0.1, seed=int(rng.integers(1 << 30)))
        logs.append((truth[0][:, 0], make_log(rng, truth, *c)))
        L = max(L, len(logs[-1][1][0]))
    N = len(logs)
    out = dict(anchors=anc, lags=np.array(LAGS),
               anchor=np.full((L, N), 255, np.uint8), seq=np.zeros((L, N), np.uint8),
               range_mm=np.zeros((L, N), np.int32), t=np.zeros((L, N)), err=np.zeros((L, N)),
               x0=np.zeros((3, N)), n_epochs=np.zeros(N, np.int32))
    eps = []
    for j, (p0, (a, s, r, t, e)) in enumerate(logs):
        n = len(a)
        out["anchor"][:n, j], out["seq"][:n, j], out["range_mm"][:n, j] = a, s, r
        out["t"][:n, j], out["err"][:n, j] = t, e
        out["t"][n:, j] = t[-1]
        out["x0"][:, j] = p0
        pg = R.RefPosGenerator(anc)
        ne = pg.feed(a, r, s, t, err=e)
        eps.append(pg.epochs(ne))
        out["n_epochs"][j] = ne
        for alg, name in ((5, "t6"), (6, "t9")):
            # KF_TOA takes the start position when useStartPosition is FALSE (Posgenerator.cpp:512-516)
            pg = R.RefPosGenerator(anc, algorithm=alg, start=p0, use_start=(alg == 6))
            assert pg.feed(a, r, s, t, err=e) == ne
            for k, lag in enumerate(LAGS):
                rc, pose, cov = pg.report(t[-1] + 0.05 + lag)
                assert rc == 0
                out.setdefault(name + "_pose", np.zeros((N, len(LAGS), 13)))[j, k] = pose
                out.setdefault(name + "_cov", np.zeros((N, len(LAGS), 36)))[j, k] = cov
            out.setdefault(name + "_failed", np.zeros(N, np.int32))[j] = pg.errors()
        # ALGORITHM_ML: the report is MLLocation::getPose on the LAST epoch (stateless, start (1,1,4));
        # variant 0 and variant 1 ignoring 2.  rc != 0: getPose threw (an error estimate of 0)
        for v, (variant, n_ign) in enumerate(((0, 0), (1, 2))):
            pg = R.RefPosGenerator(anc, algorithm=2, variant=variant, n_ignore=n_ign)
            assert pg.feed(a, r, s, t, err=e) == ne
            rc, pose, cov = pg.report(t[-1] + 0.06)
            out.setdefault("ml_rc", np.zeros((N, 2), np.int32))[j, v] = rc
            out.setdefault("ml_pose", np.zeros((N, 2, 13)))[j, v] = pose
            out.setdefault("ml_cov", np.zeros((N, 2, 36)))[j, v] = cov
    T = int(out["n_epochs"].max())
    out["ep_ranges"] = np.zeros((T, M, N)); out["ep_err"] = np.zeros((T, M, N)); out["ep_lag"] = np.full((T, N), -1.0)
    for j, ep in enumerate(eps):
        out["ep_ranges"][:ep["n"], :, j] = ep["ranges"]
        out["ep_err"][:ep["n"], :, j] = ep["err"]
        out["ep_lag"][:ep["n"], j] = ep["time_lag"]
    np.savez_compressed(os.path.join(OUT, "posgen.npz"), **out)
    print("posgen.npz: logs", N, "messages", L, "epochs", out["n_epochs"], "failed updates t6/t9",
          out["t6_failed"], out["t9_failed"])


if __name__ == "__main__":
    main()
