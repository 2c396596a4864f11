"""GPU parity, log -> report: a ranging log through kfpos_assemble_epochs, the replay kernels and
kfpos_batch_get_pose_msg against what the REFERENCE NODE publishes for the same log
(tests/golden/posgen.npz: publishers/Posgenerator.cpp + the filter classes, compiled unmodified,
driven by tests/golden/make_golden_posgen.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "posgen.npz")


def relP(P, ref):
    return np.abs(P - ref).max() / max(np.abs(ref).max(), 1e-300)


def _assembled(g):
    from roskfpos_b200.batch import assemble_epochs
    M, T = g["anchors"].shape[0], int(g["n_epochs"].max())
    return assemble_epochs(g["anchor"], g["seq"], g["range_mm"], g["t"], M, T, err=g["err"])


def test_assembled_epochs_match_the_reference_node(kflib):
    g = np.load(GOLD)
    o = _assembled(g)
    assert np.array_equal(o["n_epochs"], g["n_epochs"])
    for j in range(len(g["n_epochs"])):
        n = int(g["n_epochs"][j])
        raw = o["ranges"][:n, :, j]
        used = raw > 0
        assert np.array_equal(np.where(used, raw / 1000.0, 0.0), g["ep_ranges"][:n, :, j]), j
        assert np.array_equal(np.where(used, o["err"][:n, :, j], 0.0), g["ep_err"][:n, :, j]), j
        assert np.abs(o["dt"][1:n, j] - g["ep_lag"][1:n, j]).max(initial=0.0) < 1e-12, j
        assert np.all(o["dt"][n:, j] == -1.0)


def test_t6_log_to_report(kflib):
    """All six logs as one batch: per-filter time steps, per-ranging error estimates, logs of
    different length (dt < 0 = no epoch).  Log 2 (311 sparse epochs) to 1e-8, the rest to 1e-9."""
    from roskfpos_b200.batch import Batch
    g = np.load(GOLD)
    N = len(g["n_epochs"])
    o = _assembled(g)
    x0 = np.zeros((6, N)); x0[:3] = g["x0"]
    with Batch(kflib.MODEL_T6, N, anchors=g["anchors"], accel_noise=0.5) as b:
        b.set_state(x0)
        b.replay_epochs(o["dt"], o["ranges"], err=o["err"])
        for i, lag in enumerate(g["lags"]):
            pose, cov = b.get_pose_msg(float(lag))
            for j in range(N):
                tol = 1e-8 if j == 2 else 1e-9
                assert np.abs(pose[:, j] - g["t6_pose"][j, i]).max() < tol, (j, lag)
                assert relP(cov[:, j], g["t6_cov"][j, i]) < tol, (j, lag)


def test_t9_log_to_report(kflib):
    """As test_t6_log_to_report for the 9-state filter (one batch, per-filter time steps).  Log 2 is
    left out (its 96-iteration Newton run is a rounding amplifier, see test_oracle_golden); log 4
    has rangings without an error estimate, whose epochs the reference rejects (singular)."""
    from roskfpos_b200.batch import Batch
    g = np.load(GOLD)
    N = len(g["n_epochs"])
    o = _assembled(g)
    x0 = np.zeros((9, N)); x0[:3] = g["x0"]
    with Batch(kflib.MODEL_T9, N, anchors=g["anchors"], accel_noise=0.5, jolt=0.5) as b:
        b.set_state(x0)
        traj = b.replay_epochs(o["dt"], o["ranges"], err=o["err"], want_traj=True)
        x, P, st = b.get_state()
        assert np.array_equal(traj[-1], x[:3])  # shorter logs repeat their last position
        for i, lag in enumerate(g["lags"]):
            pose, cov = b.get_pose_msg(float(lag))
            for j in (0, 1, 3, 4, 5):
                assert np.abs(pose[:, j] - g["t9_pose"][j, i]).max() < 1e-9, (j, lag)
                assert relP(cov[:, j], g["t9_cov"][j, i]) < 1e-9, (j, lag)


def test_k8_t9_ragged_epochs_match_per_filter_replays(kflib, oracle):
    """kfpos_batch_replay_epochs for K8 and T9: every filter has its own time steps and its own
    number of epochs; the result equals the oracle run filter by filter on its own schedule."""
    from roskfpos_b200 import synth
    from roskfpos_b200.batch import Batch
    from tests.util import rel_err_cov, rel_err_state
    N, T, m = 70, 12, 8
    anc = synth.anchors_for(m)
    truth = synth.truth_lissajous(N, T, 0.1, seed=91)
    r = synth.ranges_mm(truth[1:], anc, seed=92, p_missing=0.1)
    rng = np.random.default_rng(93)
    dt = rng.integers(20, 200, size=(T, N)) / 1000.0
    dt[rng.random((T, N)) < 0.25] = -1.0  # this filter has no such epoch
    dt[0] = 0.1
    e = rng.uniform(0.005, 0.05, size=(T, m, N))
    for model, n in ((kflib.MODEL_T9, 9), (kflib.MODEL_K8, 8)):
        x0 = np.zeros((n, N))
        if n == 9:
            x0[:3] = truth[0]
        else:
            x0[:2] = truth[0][:2]; x0[6] = 0.3
        kw = dict(xml=synth.K8_XML) if n == 8 else {}
        with Batch(model, N, anchors=anc, accel_noise=0.5, jolt=0.5, **kw) as b:
            b.set_state(x0)
            b.replay_epochs(dt, r, err=e)
            x, P, st = b.get_state()
            cnt = b.counters()
        assert cnt["updates"] == (dt >= 0).sum()
        xr = np.zeros_like(x); Pr = np.zeros_like(P)
        for f in range(N):
            keep = np.nonzero(dt[:, f] >= 0)[0]
            ev = [(synth.EV_TOA, float(dt[t, f]), k * m, None) for k, t in enumerate(keep)]
            rf = np.ascontiguousarray(r[keep][:, :, f:f + 1]); ef = np.ascontiguousarray(e[keep][:, :, f:f + 1])
            if n == 9:
                ref = oracle.t9_events(x0[:, f:f + 1], None, ev, rf, None, anc, ef)
            else:
                ref = oracle.k8_replay(x0[:, f:f + 1], None, ev, rf, None, anc, ef,
                                       oracle.k8_cfg(0.5, 0.5, **synth.K8_ORACLE_CFG))
            xr[:, f] = ref["x"][:, 0]; Pr[:, f] = ref["P"][:, 0]
        assert rel_err_state(x, xr) < 1e-9 and rel_err_cov(P, Pr) < 1e-9, n


def test_ml_log_to_report(kflib):
    """ALGORITHM_ML behind the reference node: its report is MLLocation::getPose on the last epoch of
    the log.  All logs as one batch of stateless epochs (each log's last epoch), variants 0 and 1."""
    from roskfpos_b200.batch import Batch
    g = np.load(GOLD)
    N, M = len(g["n_epochs"]), g["anchors"].shape[0]
    o = _assembled(g)
    last = g["n_epochs"] - 1
    r = np.ascontiguousarray(np.stack([o["ranges"][last[j], :, j] for j in range(N)], axis=1))  # [M][N]
    e = np.ascontiguousarray(np.stack([o["err"][last[j], :, j] for j in range(N)], axis=1))
    for v, (variant, n_ign) in enumerate(((0, 0), (1, 2))):
        with Batch(kflib.MODEL_ML, N, anchors=g["anchors"], variant=variant, num_ignored_rangings=n_ign,
                   ml_start=[1.0, 1.0, 4.0]) as b:
            out = b.ml_solve(r, err=e)
        for j in range(N):
            if g["ml_rc"][j, v] != 0:
                assert out["status"][j] & 4, (j, v)  # the reference throws, the library flags SINGULAR
                continue
            assert out["status"][j] == 0
            assert np.abs(out["pos"][:, j] - g["ml_pose"][j, v, :3]).max() < 1e-9, (j, v)
            assert relP(out["cov"][:, j].reshape(3, 3), g["ml_cov"][j, v].reshape(6, 6)[:3, :3]) < 1e-9, (j, v)
