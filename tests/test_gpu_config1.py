"""GPU parity on BASELINE config 1: a single tag, 4 UWB anchors, synthetic TOA rangings at 10 Hz for
10 000 steps (kfpos_toa.launch).  1a = KalmanFilterTOA (what the launch file really selects), 1b =
KalmanFilter ranging-only at a fixed tag height ("fixed height 2D")."""
import numpy as np
import pytest

from roskfpos_b200 import synth
from tests.util import REL_TOL, rel_err_cov, rel_err_state

pytestmark = pytest.mark.gpu
T = 10000


def config1_inputs(seed):
    anc = synth.anchors_for(4)
    truth = synth.truth_lissajous(1, T, 0.1, seed=seed, z=1.0)
    r = synth.ranges_mm(truth[1:], anc, seed=seed + 1)
    return anc, truth, r


def test_config1b_k8_ten_thousand_steps_in_one_replay(kflib, oracle):
    """The 2-D inner ML has one ranging to spare with 4 anchors, so the whole 10 000-step trajectory of
    one tag (and of 15 more with other noise draws) must agree with the oracle at the end and on the
    way: one persistent launch, state never leaves the chip."""
    from roskfpos_b200.batch import Batch
    anc = synth.anchors_for(4)
    N = 16
    truth = synth.truth_lissajous(N, T, 0.1, seed=7, z=1.0)
    r = synth.ranges_mm(truth[1:], anc, seed=8)
    x0 = np.zeros((8, N)); x0[:2] = truth[0][:2]
    cfg = oracle.k8_cfg(0.5, 0.5, tag_z=1.0)
    ev = [(0, 0.1, t * 4, None) for t in range(T)]
    ref = oracle.k8_replay(x0, None, ev, r, None, anc, 0.01, cfg, want_traj=True)
    with Batch(kflib.MODEL_K8, N, anchors=anc, accel_noise=0.5, jolt=0.5, fixed_height=1.0) as b:
        b.set_state(x0)
        traj, _ = b.replay_toa(0.1, r, err=0.01, want_traj=True)
        x, P, st = b.get_state()
        cnt = b.counters()
    assert cnt["updates"] == N * T and cnt["bad"] == 0
    assert rel_err_state(x, ref["x"]) < REL_TOL and rel_err_cov(P, ref["P"]) < REL_TOL
    assert rel_err_state(traj[:, :2], ref["traj"][:, :2]) < REL_TOL
    assert [cnt["ml_iters"], cnt["cost_evals"], cnt["gain_evals"]] == list(ref["counters"][:3])
    err = np.sqrt(((x[:2] - truth[-1][:2]) ** 2).sum(axis=0))
    assert err.max() < 0.5  # still tracking after 1000 s


def test_config1a_t6_ten_thousand_steps_step_by_step(kflib, oracle):
    """With 4 anchors the 3-D inner ML is exactly determined and now and then wanders to its iteration cap,
    where the reference itself is chaotic (tests/util.py).  So the 10 000 steps are checked one at a
    time: every step must match the oracle to 1e-9 unless the oracle's own inner solve went wild on
    that step, in which case the GPU filter is re-synchronised (the tests/test_oracle_golden.py rule).
    The bar is the north star's 1e-9 (round 1 needed 1e-8 here; measured now: no wild step, worst regular
    step 3.7e-11 over the 10 000 steps, printed)."""
    tol = 1e-9
    from roskfpos_b200.batch import Batch
    anc, truth, r = config1_inputs(seed=11)
    o = oracle.T6(0.5, False, 0.0, truth[0][:, 0])
    n_wild, worst = 0, 0.0
    with Batch(kflib.MODEL_T6, 1, anchors=anc, accel_noise=0.5) as b:
        b.set_state(truth[0])
        for t in range(T):
            info = o.new_toa(0.1, r[t, :, 0].astype(np.float64) / 1000, anc, 0.01)
            b.step_toa(0.1, r[t], err=0.01)
            if t % 10 and info.ml_iters <= 100:
                continue  # read back every 10th step (and every wild one)
            x, P, st = b.get_state()
            ex = np.abs(x[:3, 0] - o.pos).max()
            eP = np.abs(P[:, 0].reshape(6, 6) - o.P).max() / max(np.abs(o.P).max(), 1e-300)
            if info.ml_iters <= 100:
                worst = max(worst, ex, eP)
            if ex > tol or eP > tol:
                assert info.ml_iters > 100, (t, ex, eP, info.ml_iters)
                n_wild += 1
                full = np.zeros((6, 1)); full[:3, 0] = o.pos
                b.set_state(full, o.P.reshape(36, 1))
        cnt = b.counters()
    print("config 1a: wild steps", n_wild, "worst regular step", worst)
    assert cnt["updates"] == T
    assert n_wild <= 10, n_wild
