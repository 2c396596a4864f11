"""GPU: the ML-initialisation branches of KalmanFilter / KalmanFilterTOAIMU (KF.cpp:244-285, TOAIMU.cpp:118-162;
kfpos_config.ml_initial_position = 1, x0 = NaN), (1) against golden vectors produced by the REFERENCE ITSELF
through its constructors without initialPosition (tests/golden/mlinit.npz, no oracle in between), state and
covariance after every callback, and (2) batches of filters that initialise at different epochs against the oracle."""
import os

import numpy as np
import pytest

from roskfpos_b200 import synth
from tests.util import assert_parity, to_metres, ulp_perturbations

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9


def relP(P, ref):
    return np.abs(P - ref).max() / max(np.abs(ref).max(), 1e-300)


def nan_eq(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    m = ~np.isnan(a)
    return bool(np.all(np.abs(a[m] - b[m]) <= tol * np.maximum(1.0, np.abs(b[m]))))


@pytest.mark.parametrize("case", ["k8_fh0_normal", "k8_fh0_few", "k8_fh1_normal", "k8_fh1_few"])
def test_k8_reference_vectors(kflib, case):
    from roskfpos_b200.batch import Batch
    g = np.load(os.path.join(GOLD, "mlinit.npz"))
    k = np.load(os.path.join(GOLD, "k8_multi.npz"))
    fh = int(g[case + "/fh"])
    xml = [str(k[n]) for n in ("xml_pos", "xml_px4", "xml_imu", "xml_mag")] + [str(g["xml_tag_fh1"]) if fh else str(k["xml_tag"])]
    x0 = np.zeros((8, 1)); x0[:2, 0] = np.nan; x0[6, 0] = float(g["init_angle"])
    col = lambda v: np.array([[float(v)]])
    n_uninit = 0
    with Batch(kflib.MODEL_K8, 1, anchors=g[case + "/anchors"], xml=xml, accel_noise=float(g["accel_noise"]),
               jolt=float(g["jolt"]), ml2d_zero_tentative_z=1, ml_initial_position=1) as b:
        assert b.cfg.use_fixed_height == fh
        b.set_state(x0)
        for i, (kind, dt, pl) in enumerate(zip(g[case + "/kinds"], g[case + "/dts"], g[case + "/payload"])):
            kind = str(kind)
            if kind == "imu":
                b.step_imu(dt, pl[0:3].reshape(3, 1), pl[12:15].reshape(3, 1), cov_ang_vel=pl[3:12], cov_acc=pl[15:24])
            elif kind == "px4":
                b.step_px4(dt, col(pl[0]), col(pl[1]), col(pl[2]), col(pl[3]), np.array([[int(pl[4])]], dtype=np.int32))
            elif kind == "compass":
                b.step_compass(dt, col(pl[0]))
            elif kind == "mag":
                b.step_mag(dt, pl[:3].reshape(3, 1))
            else:
                b.step_toa(dt, np.ascontiguousarray(pl[:8].reshape(8, 1)), err=float(g["err"]))
            x, P, st = b.get_state()
            n_uninit += bool(np.isnan(x[0, 0]))
            assert nan_eq(x[:, 0], g[case + "/x"][i], TOL), (i, kind, x[:, 0], g[case + "/x"][i])
            assert relP(P[:, 0].reshape(8, 8), g[case + "/P"][i]) < TOL, (i, kind)
            if not np.isnan(x[0, 0]):  # the report's z is mUWBtagZ: the ML estimate's z in 3-D mode (KF.cpp:257,328-332)
                pose, _ = b.get_pose_msg(0.0)
                assert abs(pose[2, 0] - g[case + "/tagz"][i]) <= 1e-9, (i, kind)
        assert st[0] & kflib.ST_UNINIT
        assert bool(st[0] & kflib.ST_ML_FEW) == case.endswith("few")
    assert n_uninit >= 5


@pytest.mark.parametrize("case", ["t9_normal", "t9_few"])
def test_t9_reference_vectors(kflib, case):
    from roskfpos_b200.batch import Batch
    g = np.load(os.path.join(GOLD, "mlinit.npz"))
    x0 = np.zeros((9, 1)); x0[:3, 0] = np.nan
    with Batch(kflib.MODEL_T9, 1, anchors=g[case + "/anchors"], accel_noise=float(g["accel_noise"]),
               jolt=float(g["jolt"]), ml_initial_position=1) as b:
        b.set_state(x0)
        for t, r in enumerate(g[case + "/ranges"]):
            b.step_toa(0.1, np.ascontiguousarray(r.reshape(8, 1)), err=float(g["err"]))
            x, P, st = b.get_state()
            assert nan_eq(x[:, 0], g[case + "/x"][t], TOL), (t, x[:, 0], g[case + "/x"][t])
            assert relP(P[:, 0].reshape(9, 9), g[case + "/P"][t]) < TOL, t


@pytest.mark.parametrize("fh", [0, 1])
def test_k8_batch_initialises_at_different_epochs(kflib, oracle, fh):
    """2000 filters, each missing the rangings of its first 0-3 epochs: one replay launch carries filters that are
    still uninitialised next to running ones; the library switches to the tuned kernel only once a launch has
    ended with every filter initialised (2-D mode; 3-D mode keeps the per-filter tag height)."""
    from roskfpos_b200.batch import Batch
    N, n_macro = 2000, 6
    anc = synth.anchors_for(8)
    w = synth.k8_workload(N, n_macro, anc, seed=synth.SEED + 77, full=True)
    rng = np.random.default_rng(5)
    late = rng.choice([0, 0, 0, 0, 1, 2, 3], N)  # an epoch without rangings "initialises" at the start point, as written
    toa_rows = [ev[2] for ev in w["events"] if ev[0] == synth.EV_TOA]
    ranges = w["ranges"].reshape(-1, N).copy()
    for k, row in enumerate(toa_rows):
        ranges[row:row + 8, late > k] = 0      # no rangings at all in that epoch
    one_few = rng.random(N) < 0.1              # ... or only two of them in the filter's first epoch
    for f in np.flatnonzero(one_few):
        ranges[toa_rows[late[f]] + 2:toa_rows[late[f]] + 8, f] = 0
    ranges = ranges.reshape(w["ranges"].shape)
    x0 = w["x0"].copy(); x0[:2] = np.nan
    cfgo = dict(synth.K8_ORACLE_CFG, use_fixed_height=fh, ml_init=1)
    ocfg = oracle.k8_cfg(0.5, 0.5, **cfgo)
    ref = oracle.k8_replay(x0, None, w["events"], ranges, w["sensors"], anc, 0.01, ocfg)
    # "stable" = the oracle's own result survives rounding-level perturbations of the ranges (tests/util.py):
    # the 3-D start (1, 1, 4) sends a few epochs into Newton runs whose end point is decided by rounding
    per = [oracle.k8_replay(x0, None, w["events"], rp, w["sensors"], anc, 0.01, ocfg)
           for rp in ulp_perturbations(to_metres(ranges), n_random=2)]
    half = len(w["events"]) // 2
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5, use_fixed_height=fh,
               ml_initial_position=1) as b:
        b.set_state(x0)
        ev = w["events"]
        b.replay_events(ev[:half], ranges=ranges, sensors=w["sensors"], err=0.01)
        b.replay_events(ev[half:], ranges=ranges, sensors=w["sensors"], err=0.01)
        x, P, st = b.get_state()
        pose, _ = b.get_pose_msg(0.0)
    assert not np.isnan(x[:2]).any() and (st & kflib.ST_UNINIT).all()
    assert np.array_equal((st & kflib.ST_ML_FEW) != 0, (ref["status"] & 2) != 0)
    got = dict(x=x, P=P)
    assert_parity(got, ref, per, float_keys=("x",), cov_keys=("P",), min_stable=0.85, max_tie_frac=1e-3,
                  what=f"K8 ML initialisation fh={fh}")
    assert np.abs(pose[2] - ref["tagz"]).max() < 1e-9


def test_t9_batch_initialises_at_different_epochs(kflib, oracle):
    from roskfpos_b200.batch import Batch
    N, T = 2000, 12
    anc = synth.anchors_for(8)
    truth = synth.truth_lissajous(N, T, 0.1, seed=21)
    r = synth.ranges_mm(truth[1:], anc, seed=22)
    rng = np.random.default_rng(6)
    late = rng.integers(0, 4, N)
    for k in range(3):
        r[k][:, late > k] = 0
    x0 = np.zeros((9, N)); x0[:3] = np.nan
    ev = [(synth.EV_TOA, 0.1, 8 * t) for t in range(T)]
    ref = oracle.t9_events(x0, None, ev, r, None, anc, 0.01, ml_init=1)
    per = [oracle.t9_events(x0, None, ev, rp, None, anc, 0.01, ml_init=1) for rp in ulp_perturbations(to_metres(r), n_random=2)]
    with Batch(kflib.MODEL_T9, N, anchors=anc, accel_noise=0.5, jolt=0.5, ml_initial_position=1) as b:
        b.set_state(x0)
        b.replay_events(ev[:5], ranges=r, err=0.01)
        b.replay_events(ev[5:], ranges=r, err=0.01)
        x, P, st = b.get_state()
    assert not np.isnan(x).any()
    assert_parity(dict(x=x, P=P), ref, per, float_keys=("x",), cov_keys=("P",), min_stable=0.98,
                  max_tie_frac=1e-3, what="T9 ML initialisation")
