#!/usr/bin/env python
"""Generates tests/golden/pose_msg.npz from the REFERENCE ITSELF (oracle/_ref/libkfref.so, see
make_golden.py): for each of the three filters a short replay, then getPose polled at several
time lags and read the way PosGenerator::publishPositionReport reads a report
(Posgenerator.cpp:385-470).  Stored with the filter state (x, P) at the time of the poll so that
the packing can be checked on its own.  Run:  python tests/golden/make_golden_pose.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_py as R  # noqa: E402
from roskfpos_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LAGS = [0.0, 0.013, 0.1, 0.37]


def main():
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    m, T = 8, 12
    anc = synth.anchors_for(m)
    truth = synth.truth_lissajous(1, T, 0.1, seed=3)
    r = synth.ranges_mm(truth[1:], anc, seed=4).astype(np.float64) / 1000
    rng = np.random.default_rng(5)
    compass = rng.uniform(-3, 3, size=T)  # K8: one compass sample before every third TOA epoch
    out = dict(anchors=anc, x0=truth[0][:, 0], ranges=r[:, :, 0], lags=np.array(LAGS), dt=0.1, compass=compass,
               compass_dt=0.01, k8_init_angle=0.3)
    for name, f in (("t6", R.RefT6(0.5, False, 0.5, truth[0][:, 0])),
                    ("k8", R.RefK8(0.5, 0.3, 0.5, truth[0][:, 0])),
                    ("t9", R.RefT9(0.5, 0.5, truth[0][:, 0]))):
        rc0, _, _ = R.pose_msg(f, 0.1)
        assert rc0 == 4  # getPose is false before the first measurement
        xs, Ps, poses, covs = [], [], [], []
        for t in range(T):
            if name == "k8" and t % 3 == 0:  # give the heading and the rates something to show
                assert f.new_compass(0.01, float(compass[t])) == 0
            assert f.new_toa(0.1, r[t, :, 0], anc, np.full(m, 0.01)) == 0
            if t % 4 == 3:
                x, P = f.state()
                for lag in LAGS:
                    rc, pose, cov = R.pose_msg(f, lag)
                    assert rc == 0
                    xs.append(np.asarray(x).ravel()); Ps.append(np.asarray(P).ravel())
                    poses.append(pose); covs.append(cov)
        out[name + "_x"] = np.array(xs); out[name + "_P"] = np.array(Ps)
        out[name + "_pose"] = np.array(poses); out[name + "_cov"] = np.array(covs)
    np.savez_compressed(os.path.join(OUT, "pose_msg.npz"), **out)
    print("wrote pose_msg.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
