#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE ITSELF: the reference's own .cpp files
compiled against the shim headers (oracle/_ref/libkfref.so, `make -C oracle ref`; needs
/root/reference, so it only runs in the build container).  Each file holds seeded inputs and
the outputs the reference code produced for them; tests/test_oracle_golden.py replays the
inputs through the CPU oracle.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_py as R  # noqa: E402
from roskfpos_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
XML = R.XML_DEFAULT


def k8_events(T, seed, m=8):
    """A deterministic multi-sensor event script: list of (kind, dt, payload)."""
    rng = np.random.default_rng(seed)
    anc = synth.anchors_for(m)
    truth = synth.truth_lissajous(1, T, 0.1, seed=seed + 1)
    r = synth.ranges_mm(truth[1:], anc, seed=seed + 2, p_missing=0.1).astype(np.float64) / 1000
    ev = []
    for t in range(T):
        for _ in range(3):
            w = [0.0, 0.0, rng.normal(0.1, 0.05)]
            a = [rng.normal(0, 0.3), rng.normal(0, 0.3), 9.8]
            cav = np.diag([1e-3, 1e-3, 2e-3]).ravel()
            cac = np.array([[4e-3, 1e-4, 0], [1e-4, 5e-3, 0], [0, 0, 6e-3]]).ravel()
            ev.append(("imu", 0.02, np.concatenate([w, cav, a, cac])))
        q = int(rng.integers(1, 3) * 100)
        ev.append(("px4", 0.01, np.array([rng.normal(0, 0.002), rng.normal(0, 0.002), rng.normal(0, 0.001),
                                          33333.0, q])))
        if t % 2:
            ev.append(("compass", 0.01, np.array([rng.uniform(-4, 4)])))
        else:
            ev.append(("mag", 0.01, np.array([np.cos(0.3 + 0.01 * t), np.sin(0.3 + 0.01 * t), 0.1])))
        ev.append(("toa", 0.02, r[t, :, 0]))
    ev[0] = (ev[0][0], 0.1, ev[0][2])  # the reference uses 0.1 s for its first update (KF.cpp:238)
    return anc, truth[0][:, 0], ev


def main():
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    # ---------------------------------------------------------------- T6
    for name, m, T, kw in (("t6_m8", 8, 120, {}), ("t6_m16_missing", 16, 60, dict(p_missing=0.3)),
                           ("t6_m8_loo", 8, 80, dict(p_nlos=0.2))):
        anc = synth.anchors_for(m)
        truth = synth.truth_lissajous(1, T, 0.1, seed=len(name))
        r = synth.ranges_mm(truth[1:], anc, seed=7 * len(name), **kw).astype(np.float64) / 1000
        loo = name.endswith("loo")
        f = R.RefT6(0.5, loo, 0.5, truth[0][:, 0])
        xs, Ps = [], []
        for t in range(T):
            assert f.new_toa(0.1, r[t, :, 0], anc, np.full(m, 0.01)) == 0
            p, P = f.state()
            xs.append(p); Ps.append(P)
        rc, pp, Pp = f.get_pose(0.05)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), anchors=anc, x0=truth[0][:, 0], ranges=r[:, :, 0],
                            dt=0.1, err=0.01, accel_noise=0.5, loo=int(loo), thr=0.5, x=np.array(xs),
                            P=np.array(Ps), pose_dt=0.05, pose_x=pp, pose_P=Pp)
    # ---------------------------------------------------------------- T9 (ranging path)
    anc = synth.anchors_for(8)
    truth = synth.truth_lissajous(1, 100, 0.1, seed=33)
    r = synth.ranges_mm(truth[1:], anc, seed=34).astype(np.float64) / 1000
    f = R.RefT9(0.5, 0.5, truth[0][:, 0])
    xs, Ps = [], []
    for t in range(100):
        assert f.new_toa(0.1, r[t, :, 0], anc, np.full(8, 0.01)) == 0
        x, P = f.state()
        xs.append(x); Ps.append(P)
    np.savez_compressed(os.path.join(OUT, "t9_m8.npz"), anchors=anc, x0=truth[0][:, 0], ranges=r[:, :, 0], dt=0.1,
                        err=0.01, accel_noise=0.5, jolt=0.5, x=np.array(xs), P=np.array(Ps))
    # ---------------------------------------------------------------- K8 (all five callbacks)
    anc, p0, ev = k8_events(60, seed=50)
    f = R.RefK8(0.5, 0.3, 0.5, [p0[0], p0[1], 0.0])
    xs, Ps = [], []
    for kind, dt, pl in ev:
        if kind == "imu":
            rc = f.new_imu(dt, pl[0:3], pl[3:12], pl[12:15], pl[15:24])
        elif kind == "px4":
            rc = f.new_px4(dt, pl[0], pl[1], pl[2], pl[3], int(pl[4]))
        elif kind == "compass":
            rc = f.new_compass(dt, pl[0])
        elif kind == "mag":
            rc = f.new_mag(dt, pl)
        else:
            rc = f.new_toa(dt, pl, anc, np.full(len(pl), 0.01))
        assert rc == 0
        x, P = f.state()
        xs.append(x); Ps.append(P)
    kinds = np.array([e[0] for e in ev])
    dts = np.array([e[1] for e in ev])
    payload = np.zeros((len(ev), 24))
    for i, e in enumerate(ev):
        payload[i, :len(e[2])] = e[2]
    np.savez_compressed(os.path.join(OUT, "k8_multi.npz"), anchors=anc, x0=p0, kinds=kinds, dts=dts, payload=payload,
                        accel_noise=0.5, init_angle=0.3, jolt=0.5, err=0.01, x=np.array(xs), P=np.array(Ps),
                        xml_pos=XML["kfpos_pos"], xml_px4=XML["kfpos_px4"], xml_tag=XML["kfpos_tag"],
                        xml_imu=XML["kfpos_imu"], xml_mag=XML["kfpos_mag"])
    # ---------------------------------------------------------------- ML
    rng = np.random.default_rng(77)
    cases = []
    for use2d, variant, n_ign, m in ((0, 0, 0, 8), (1, 0, 0, 8), (0, 0, 0, 16), (0, 1, 2, 8), (1, 1, 2, 8),
                                     (0, 2, 0, 5), (0, 0, 0, 3), (1, 0, 0, 2)):
        anc = synth.anchors_for(8 if m < 8 else m)[:m]
        start = [1.0, 1.0, 1.0 if use2d else 4.0]
        ml = R.RefML(use2d, variant, n_ign, start)
        for _ in range(40):
            tp = np.array([rng.uniform(1, 9), rng.uniform(1, 9), 1.0])
            rr = np.floor((np.sqrt(((anc - tp) ** 2).sum(1)) + rng.normal(0, 0.1, m)) * 1000) / 1000
            out = ml.solve(rr, anc, np.full(m, 0.01), mode=1)
            assert out["rc"] == 0
            cov = np.zeros((3, 3))
            d = out["cov"].shape[0]
            cov[:d, :d] = out["cov"]
            cases.append(dict(use2d=use2d, variant=variant, n_ignore=n_ign, m=m, anchors=np.pad(anc, ((0, 16 - m), (0, 0))),
                              ranges=np.pad(rr, (0, 16 - m)), start=start, pos=out["pos"], cov=cov, cov_dim=d))
    keys = cases[0].keys()
    np.savez_compressed(os.path.join(OUT, "ml_cases.npz"), **{k: np.array([c[k] for c in cases]) for k in keys})
    print("golden files written to", OUT)


if __name__ == "__main__":
    main()
