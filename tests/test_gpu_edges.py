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
