"""GPU parity: kfpos_batch_get_pose_msg (getPose in the publisher's layout) against the oracle's
restatement, which tests/test_oracle_golden.py pins to the reference's own classes."""
import numpy as np
import pytest

from roskfpos_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["t6", "k8", "t9"])
def test_pose_msg_parity(kflib, oracle, model):
    from roskfpos_b200.batch import Batch
    N, T, m = 700, 6, 8
    anc = synth.anchors_for(m)
    truth = synth.truth_lissajous(N, T, 0.1, seed=21)
    r = synth.ranges_mm(truth[1:], anc, seed=22)
    mid, n = {"t6": (kflib.MODEL_T6, 6), "k8": (kflib.MODEL_K8, 8), "t9": (kflib.MODEL_T9, 9)}[model]
    kw = dict(accel_noise=0.5, jolt=0.5)
    if model == "k8":
        kw["xml"] = synth.K8_XML
    x0 = np.zeros((n, N))
    x0[:2] = truth[0][:2]
    if model != "k8":
        x0[2] = truth[0][2]
    else:
        x0[6] = np.random.default_rng(1).uniform(-3, 3, N)  # headings all around the circle
        x0[7] = 0.2
    with Batch(mid, N, anchors=anc, **kw) as b:
        b.set_state(x0)
        with pytest.raises(kflib.KfposError) as ei:  # getPose is false before the first measurement
            b.get_pose_msg(0.05)
        assert ei.value.code == -4
        b.replay_toa(0.1, r, err=0.01)
        for lag in (0.0, 0.05, 0.4):
            xp, Pp = b.get_pose(lag)
            pose, cov = b.get_pose_msg(lag)
            ref_pose, ref_cov = oracle.pose_msg({"t6": 1, "k8": 2, "t9": 3}[model], xp, Pp, tag_z=1.049)
            assert np.array_equal(pose[[0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 12]], ref_pose[[0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 12]])
            assert np.abs(pose[5:7] - ref_pose[5:7]).max() < 1e-15  # sin / cos of theta / 2
            assert np.array_equal(cov, ref_cov)
            if model == "k8":
                assert np.allclose(pose[5] ** 2 + pose[6] ** 2, 1.0) and np.all(pose[2] == 1.049)
