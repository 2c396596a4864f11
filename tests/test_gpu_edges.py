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
