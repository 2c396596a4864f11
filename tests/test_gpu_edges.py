"""GPU: edge cases of the C ABI -- error behaviour, maximum anchor count, tiny and odd batch sizes."""
import ctypes as C

import numpy as np
import pytest

from roskfpos_b200 import synth
from tests.util import REL_TOL, rel_err_cov, rel_err_state

pytestmark = pytest.mark.gpu


def test_error_codes(kflib):
    from roskfpos_b200.batch import Batch
    anc = synth.anchors_for(8)
    r = np.full((2, 8, 4), 3000, dtype=np.int32)
    with Batch(kflib.MODEL_T6, 4, accel_noise=0.5) as b:  # anchors not set yet
        b.set_state(np.zeros((6, 4)))
        with pytest.raises(kflib.KfposError) as e:
            b.replay_toa(0.1, r)
        assert e.value.code == -4  # KFPOS_ERR_NOT_READY
        b.set_anchors(anc)
        b.replay_toa(0.1, r)
        with pytest.raises(kflib.KfposError) as e:  # ML call on an EKF batch
            b.ml_solve(r[0])
        assert e.value.code == -1
        with pytest.raises(kflib.KfposError) as e:  # sensor step of a model that has none
            b.step_compass(0.1, np.zeros(4))
        assert e.value.code in (-1, -5)
    with pytest.raises(kflib.KfposError) as e:  # more anchors than KFPOS_MAX_ANCHORS
        Batch(kflib.MODEL_T6, 4, accel_noise=0.5, anchors=np.zeros((33, 3)))
    assert e.value.code == -1
    with pytest.raises(kflib.KfposError) as e:  # the two T6 heuristics are alternatives
        with Batch(kflib.MODEL_T6, 4, anchors=anc, accel_noise=0.5, ignore_worst_anchor=1, variant=1) as b:
            b.set_state(np.zeros((6, 4)))
            b.replay_toa(0.1, r)
    assert e.value.code in (-1, -2, -5)
    h = C.c_void_p()
    cfg = kflib.KfposConfig()
    kflib.lib().kfpos_config_default(C.byref(cfg))
    assert kflib.lib().kfpos_batch_create(C.byref(h), 99, kflib.MODEL_T6, 4, C.byref(cfg)) == -2  # no such device
    assert kflib.lib().kfpos_batch_create(C.byref(h), 0, 7, 4, C.byref(cfg)) == -1                 # no such model
    assert kflib.lib().kfpos_batch_create(C.byref(h), 0, kflib.MODEL_T6, 0, C.byref(cfg)) == -1    # empty batch


@pytest.mark.parametrize("N", [1, 31, 33, 129])
def test_maximum_anchor_count_and_odd_batch_sizes(kflib, oracle, N):
    """32 anchor slots (KFPOS_MAX_ANCHORS), batches that do not fill a warp or a block."""
    from roskfpos_b200.batch import Batch
    rng = np.random.default_rng(N)
    m, T = 32, 6
    anc = np.column_stack([rng.uniform(0, 10, m), rng.uniform(0, 10, m), rng.uniform(0.3, 3.0, m)])
    truth = synth.truth_lissajous(N, T, 0.1, seed=3 + N)
    r = synth.ranges_mm(truth[1:], anc, seed=4 + N, p_missing=0.2)
    ref = oracle.t6_replay(truth[0], None, r, anc, 0.1, 0.01)
    with Batch(kflib.MODEL_T6, N, anchors=anc, accel_noise=0.5) as b:
        b.set_state(truth[0])
        b.replay_toa(0.1, r, err=0.01)
        x, P, st = b.get_state()
        cnt = b.counters()
    assert rel_err_state(x[:3], ref["x"]) < REL_TOL and rel_err_cov(P, ref["P"]) < REL_TOL
    assert cnt["updates"] == N * T
    with Batch(kflib.MODEL_ML, N, anchors=anc) as b:
        got = b.ml_solve(r[0], err=0.01)
    mref = oracle.ml_batch(r[0], anc, 0.01, [1.0, 1.0, 4.0])
    ok = mref["iters"] < 100  # epochs whose Newton iteration wanders are compared in test_gpu_ml.py
    assert np.abs(got["pos"][:, ok] - mref["pos"][:, ok]).max() < 1e-9
    assert np.array_equal(got["iters"][ok], mref["iters"][ok])


@pytest.mark.parametrize("N", [1, 33, 130])
def test_event_stream_models_at_odd_batch_sizes(kflib, oracle, N):
    """K8 (full multi-sensor schedule) and T9 (rangings + accelerometer) with batches that do not fill a
    warp / a block: the partially filled warp takes part in the warp-level re-convergence."""
    from roskfpos_b200.batch import Batch
    anc = synth.anchors_for(8)
    w = synth.k8_workload(N, 3, anc, seed=70 + N, full=True)
    cfg = oracle.k8_cfg(0.5, 0.5, **synth.K8_ORACLE_CFG)
    ref = oracle.k8_replay(w["x0"], None, w["events"], w["ranges"], w["sensors"], anc, 0.01, cfg)
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(w["x0"])
        b.replay_events(w["events"], ranges=w["ranges"], sensors=w["sensors"], err=0.01)
        x, P, st = b.get_state()
    assert rel_err_state(x, ref["x"]) < REL_TOL and rel_err_cov(P, ref["P"]) < REL_TOL
    # T9: TOA epochs interleaved with accelerometer samples
    T = 6
    truth = synth.truth_lissajous(N, T, 0.1, seed=80 + N)
    r = synth.ranges_mm(truth[1:], anc, seed=81 + N)
    rng = np.random.default_rng(N)
    acc = rng.normal(0, 0.2, size=(3 * T, N))
    cov = [0.01, 0.001, 0.0, 0.001, 0.02, 0.0, 0.0, 0.0, 0.03]
    events = []
    for t in range(T):
        events.append((kflib.EV_IMU, 0.04, 3 * t, cov))
        events.append((kflib.EV_TOA, 0.06, t * 8, None))
    x0 = np.zeros((9, N)); x0[:3] = truth[0]
    ref9 = oracle.t9_events(x0, None, events, r, acc, anc, 0.01)
    with Batch(kflib.MODEL_T9, N, anchors=anc, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(x0)
        b.replay_events(events, ranges=r, sensors=acc, err=0.01)
        x9, P9, st9 = b.get_state()
    assert rel_err_state(x9, ref9["x"]) < REL_TOL and rel_err_cov(P9, ref9["P"]) < REL_TOL


def test_rangings_without_error_estimate(kflib, oracle):
    """Per-measurement mode with errorEstimation == 0 on some rangings (what PosGenerator stores for
    a ranging that carried none, Posgenerator.cpp:94,236): the ML weights 1/e are infinite, the
    reference's dense solver rejects the normal matrix and the epoch is skipped (covariance stays
    predicted, status SINGULAR).  Epochs of the same filter with estimates everywhere update normally."""
    from roskfpos_b200.batch import Batch
    N, T, m = 1024, 8, 8
    anc = synth.anchors_for(m)
    truth = synth.truth_lissajous(N, T, 0.1, seed=71)
    r = synth.ranges_mm(truth[1:], anc, seed=72, p_missing=0.1)
    rng = np.random.default_rng(73)
    e = rng.uniform(0.005, 0.05, size=(T, m, N))
    hit = rng.random((T, 1, N)) < 0.4  # 40 % of the epochs have some rangings without an estimate
    e[np.broadcast_to(hit, e.shape) & (rng.random(e.shape) < 0.3)] = 0.0
    ref = oracle.t6_replay(truth[0], None, r, anc, 0.1, e)
    with Batch(kflib.MODEL_T6, N, anchors=anc, accel_noise=0.5) as b:
        b.set_state(truth[0])
        b.replay_toa(0.1, r, err=e)
        x, P, st = b.get_state()
    assert np.isfinite(x).all() and np.isfinite(P).all()
    from tests.util import assert_parity, to_metres, ulp_perturbations
    per = [oracle.t6_replay(truth[0], None, pr, anc, 0.1, e) for pr in ulp_perturbations(to_metres(r))]
    for d in [ref] + per:
        d["status"] = d["status"] & 4
    # units that are not stable under 1-ulp input changes (an ill-conditioned 6-ranging epoch) are exempt
    assert_parity(dict(x=x[:3], P=P, status=st & 4), ref, per, float_keys=("x",), cov_keys=("P",),
                  int_keys=("status",), min_stable=0.99, what="T6 with missing error estimates")
    assert (st & 4).any() and not (st & 4).all()
    x0 = np.zeros((9, N)); x0[:3] = truth[0]
    ref9 = oracle.t9_replay(x0, None, r, anc, 0.1, e)
    with Batch(kflib.MODEL_T9, N, anchors=anc, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(x0)
        b.replay_toa(0.1, r, err=e)
        x, P, st = b.get_state()
    assert np.isfinite(x).all() and np.isfinite(P).all()
    per = [oracle.t9_replay(x0, None, pr, anc, 0.1, e) for pr in ulp_perturbations(to_metres(r))]
    for d in [ref9] + per:
        d["status"] = d["status"] & 4
    assert_parity(dict(x=x, P=P, status=st & 4), ref9, per, float_keys=("x",), cov_keys=("P",),
                  int_keys=("status",), min_stable=0.99, what="T9 with missing error estimates")
    assert (st & 4).any()
    # K8: compass events between the ranging epochs, 2-D inner ML (solve_sym2)
    comp = rng.uniform(-3, 3, size=(T, N))
    events = []
    for t in range(T):
        events.append((synth.EV_COMPASS, 0.03, t, None))
        events.append((synth.EV_TOA, 0.07, t * m, None))
    x8 = np.zeros((8, N)); x8[:2] = truth[0][:2]; x8[6] = comp[0]
    cfg = oracle.k8_cfg(0.5, 0.5, **synth.K8_ORACLE_CFG)
    ref8 = oracle.k8_replay(x8, None, events, r, comp, anc, e, cfg)
    per = [oracle.k8_replay(x8, None, events, pr, comp, anc, e, cfg) for pr in ulp_perturbations(to_metres(r))]
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(x8)
        b.replay_events(events, ranges=r, sensors=comp, err=e)
        x, P, st = b.get_state()
    assert np.isfinite(x).all() and np.isfinite(P).all()
    for d in [ref8] + per:
        d["status"] = d["status"] & 4
    assert_parity(dict(x=x, P=P, status=st & 4), ref8, per, float_keys=("x",), cov_keys=("P",),
                  int_keys=("status",), min_stable=0.99, what="K8 with missing error estimates")
    assert (st & 4).any()
    for use2d in (False, True):
        mref = oracle.ml_batch(r[0], anc, e[0], [1.0, 1.0, 4.0], use2d=use2d)
        with Batch(kflib.MODEL_ML, N, anchors=anc, use2d=int(use2d)) as b:
            got = b.ml_solve(r[0], err=e[0])
        assert np.array_equal(got["status"] & 4, mref["status"] & 4) and (mref["status"] & 4).any()
        ok = (mref["status"] == 0) & (mref["iters"] < 100)
        assert ok.sum() > N // 3
        assert np.abs(got["pos"][:, ok] - mref["pos"][:, ok]).max() < 1e-9


@pytest.mark.parametrize("err", [0.0, 0.01])
def test_empty_epochs_and_zero_scalar_error_estimate(kflib, oracle, err):
    """Epochs without any ranging (PosGenerator forwards them when every range is <= 0) leave the
    predicted state and covariance untouched, also with errorEstimation == 0 (what the C++ mirror of
    KalmanFilterTOA passes for an empty call).  With rangings AND a scalar errorEstimation of exactly 0
    the reference's Newton solve divides by it and its solver throws: update skipped, status SINGULAR
    (with fewer than 4 / 3 rangings the Newton solve is never entered and the update is a normal one)."""
    from roskfpos_b200.batch import Batch
    from tests.util import assert_parity, to_metres, ulp_perturbations
    N, T, m = 512, 6, 8
    anc = synth.anchors_for(m)
    truth = synth.truth_lissajous(N, T, 0.1, seed=91)
    r = synth.ranges_mm(truth[1:], anc, seed=92, p_missing=0.15)
    r[1, :, : N // 2] = 0   # an empty epoch for half of the filters
    r[3, :, ::3] = 0
    r[2, 3:, 5::7] = 0      # three rangings only
    pers = ulp_perturbations(to_metres(r))
    keys = dict(float_keys=("x",), cov_keys=("P",), int_keys=("status",), min_stable=0.9)

    def st5(d):
        return dict(x=d["x"], P=d["P"], status=d["status"] & 5)

    ref = st5(oracle.t6_replay(truth[0], None, r, anc, 0.1, err))
    per = [st5(oracle.t6_replay(truth[0], None, q, anc, 0.1, err)) for q in pers]
    with Batch(kflib.MODEL_T6, N, anchors=anc, accel_noise=0.5) as b:
        b.set_state(truth[0])
        b.replay_toa(0.1, r, err=err)
        x, P, st = b.get_state()
    assert np.isfinite(x).all() and np.isfinite(P).all()
    assert (ref["status"] & 1).any() and ((ref["status"] & 4) != 0).any() == (err == 0.0)
    assert_parity(dict(x=x[:3], P=P, status=st & 5), ref, per, what=f"T6 empty epochs err={err}", **keys)
    # step API, one filter, an empty epoch first (the path of the C++ mirror)
    with Batch(kflib.MODEL_T6, 1, anchors=anc, accel_noise=0.5) as b:
        b.set_state(truth[0][:, :1])
        b.step_toa(0.1, np.zeros((m, 1)), err=0.0)
        b.step_toa(0.1, np.ascontiguousarray(r[0][:, :1]), err=0.01)
        x1, P1, st1 = b.get_state()
    r2 = np.stack([np.zeros((m, 1), dtype=r.dtype), r[0][:, :1]])
    ref1 = oracle.t6_replay(truth[0][:, :1], None, r2, anc, 0.1, 0.01)
    assert np.isfinite(x1).all() and rel_err_state(x1[:3], ref1["x"]) < REL_TOL and rel_err_cov(P1, ref1["P"]) < REL_TOL
    # T9 and K8 (ranging-only schedules) and the ML solver
    x0 = np.zeros((9, N)); x0[:3] = truth[0]
    ref9 = st5(oracle.t9_replay(x0, None, r, anc, 0.1, err))
    per = [st5(oracle.t9_replay(x0, None, q, anc, 0.1, err)) for q in pers]
    with Batch(kflib.MODEL_T9, N, anchors=anc, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(x0)
        b.replay_toa(0.1, r, err=err)
        x, P, st = b.get_state()
    assert np.isfinite(x).all()
    assert_parity(dict(x=x, P=P, status=st & 5), ref9, per, what=f"T9 empty epochs err={err}", **keys)
    events = [(synth.EV_TOA, 0.1, t * m, None) for t in range(T)]
    x8 = np.zeros((8, N)); x8[:2] = truth[0][:2]
    cfg = oracle.k8_cfg(0.5, 0.5, **synth.K8_ORACLE_CFG)
    ref8 = st5(oracle.k8_replay(x8, None, events, r, np.zeros((1, N)), anc, err, cfg))
    per = [st5(oracle.k8_replay(x8, None, events, q, np.zeros((1, N)), anc, err, cfg)) for q in pers]
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(x8)
        b.replay_events(events, ranges=r, sensors=np.zeros((1, N)), err=err)
        x, P, st = b.get_state()
    assert np.isfinite(x).all()
    assert_parity(dict(x=x, P=P, status=st & 5), ref8, per, what=f"K8 empty epochs err={err}", **keys)
    for use2d in (False, True):
        mref = oracle.ml_batch(r[0], anc, err, [1.0, 1.0, 4.0], use2d=use2d)
        with Batch(kflib.MODEL_ML, N, anchors=anc, use2d=int(use2d)) as b:
            got = b.ml_solve(r[0], err=err)
        assert np.array_equal(got["status"] & 6, mref["status"] & 6)
