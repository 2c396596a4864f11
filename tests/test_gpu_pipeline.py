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


def test_k8_multi_sensor_logs_to_report(kflib, oracle):
    """The whole multi-sensor chain through the C ABI: six tags' raw ranging logs and PX4Flow / IMU / magnetometer /
    compass message logs (each tag in its own order) -> kfpos_assemble_epochs_t -> kfpos_merge_streams ->
    kfpos_batch_replay_events_ragged -> kfpos_batch_get_pose_msg, against the report the reference NODE published
    for each tag (tests/golden/node_k8.npz, produced by PosGenerator + KalmanFilter fed message by message)."""
    from roskfpos_b200.batch import Batch, assemble_epochs, merge_streams
    g = np.load(os.path.join(os.path.dirname(GOLD), "node_k8.npz"))
    k = np.load(os.path.join(os.path.dirname(GOLD), "k8_multi.npz"))
    M, N = 8, g["anchor"].shape[1]
    ep = assemble_epochs(g["anchor"], g["seq"], g["range_mm"], g["t"], M, 64, err=g["err"])
    ref_ep = oracle.assemble(g["anchor"], g["seq"], g["range_mm"], g["t"], M, 64, err=g["err"])
    for key in ("ranges", "err", "dt", "n_epochs", "t"):
        assert np.array_equal(ep[key], ref_ep[key]), key
    # the node was polled before the 50 ms timer sent each log's last epoch
    ep["t"][ep["t"] > g["t_report"][None, :]] = -1.0
    sensors = {1: (g["px4_t"], g["px4"]), 2: (g["imu_t"], g["imu"]), 3: (g["mag_t"], g["mag"]),
               4: (g["compass_t"], g["compass"])}
    slots = [2, 0, 1, 2, 3, 4, 0, 2] * 70
    mg = merge_streams(ep["t"], ep["ranges"], ep["err"], sensors, slots, imu_aux=g["imu_aux"])
    ref_mg = oracle.merge_streams(ep["t"], ep["ranges"], ep["err"], sensors, slots)
    for key in ("dt", "ranges", "err", "sensors", "n_dropped"):
        assert np.array_equal(mg[key], ref_mg[key]), key
    assert mg["n_dropped"].max() == 0
    xml = [str(k[n]) for n in ("xml_pos", "xml_px4", "xml_tag", "xml_imu", "xml_mag")]
    x0 = np.zeros((8, N)); x0[:2] = g["x0"]; x0[6] = g["ang0"]
    n_rng = mg["ranges"].shape[0] // M
    with Batch(kflib.MODEL_K8, N, anchors=g["anchors"], xml=xml, accel_noise=float(g["accel_noise"]),
               jolt=float(g["jolt"]), ml2d_zero_tentative_z=1) as b:
        b.set_state(x0)
        b.replay_events(mg["events"], ranges=mg["ranges"].reshape(n_rng, M, N), sensors=mg["sensors"], err=mg["err"],
                        dt_per_filter=mg["dt"])
        for f in range(N):
            last = [ep["t"][:, f].max(), g["imu_t"][:, f].max(), g["mag_t"][:, f].max(), g["compass_t"][:, f].max()]
            ok = (g["px4_t"][:, f] >= 0) & (g["px4"][:, 4, f] != 0)
            if ok.any():
                last.append(g["px4_t"][ok, f].max())
            pose, cov = b.get_pose_msg(float(g["t_report"][f]) - max(last))
            assert np.abs(pose[:, f] - g["pose"][:, f]).max() < 1e-9, f
            assert np.abs(cov[:, f] - g["cov"][:, f]).max() <= 1e-9 * np.abs(g["cov"][:, f]).max(), f


def test_stream_merger_bit_exact_at_scale(kflib, oracle):
    """4000 tags with ragged, differently ordered streams (missing sensors, ended streams, too few slots for some):
    the merge kernel equals the oracle merger bit for bit."""
    from roskfpos_b200.batch import merge_streams
    rng = np.random.default_rng(11)
    N, M, T = 4000, 8, 12
    def stream(Lk, rate):
        t = np.cumsum(rng.uniform(0.2, 1.8, (Lk, N)) * rate, axis=0).round(4)
        n = rng.integers(0, Lk + 1, N)
        t[np.arange(Lk)[:, None] >= n[None, :]] = -1.0
        return t
    t_epoch = stream(T, 0.1)
    ranges = rng.integers(-1, 20000, (T, M, N)).astype(np.int32)
    err = rng.uniform(0.01, 0.03, (T, M, N))
    sensors = {2: (stream(40, 0.03), rng.normal(size=(40, 3, N))), 1: (stream(15, 0.08), rng.normal(size=(15, 5, N))),
               4: (stream(9, 0.13), rng.normal(size=(9, 1, N)))}
    slots = [2, 2, 1, 2, 0, 4, 3] * 14  # the magnetometer slots stay empty; some tags run out of slots
    got = merge_streams(t_epoch, ranges, err, sensors, slots)
    ref = oracle.merge_streams(t_epoch, ranges, err, sensors, slots)
    for key in ("dt", "ranges", "err", "sensors", "n_dropped"):
        assert np.array_equal(got[key], ref[key]), key
    assert ref["n_dropped"].max() > 0 and (ref["n_dropped"] == 0).mean() > 0.3
