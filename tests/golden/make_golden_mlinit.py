#!/usr/bin/env python
"""Golden vectors for the ML-INITIALISATION branches (KalmanFilter.cpp:244-285, KalmanFilterTOAIMU.cpp:118-162),
produced by the REFERENCE ITSELF (oracle/_ref/libkfref.so: the reference's own .cpp files) through the
constructors WITHOUT initialPosition -- what PosGenerator builds when useStartPosition is 0.

tests/golden/mlinit.npz holds, per case `c`: c/kinds, c/dts, c/payload (the event script), c/x, c/P (state and
covariance after EVERY callback), c/rc (0 ok, 2 = the reference threw std::logic_error), c/tagz (K8: mUWBtagZ
after every callback).  Cases:
  k8_fh{0,1}_{normal,few}: K8 with useFixedHeight 0 / 1; `normal` = IMU, compass and PX4 samples arrive before
      the first epoch; `few` = the first epoch has only two valid rangings (the filter is then "initialised" at
      the solver's start point with P = 0 and the reference throws on the empty covariance matrix)
  t9_{normal,few}: T9, ranging epochs only (its IMU rows throw)
Run:  python tests/golden/make_golden_mlinit.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_py as R  # noqa: E402
from roskfpos_b200 import synth  # noqa: E402
from tests.golden.make_golden import k8_events  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
XML_FH1 = '<config><uwb useFixedHeight="1" fixedHeight="1.049" tagId="0"/></config>'


def run_k8(fh, few, seed):
    anc, p0, ev = k8_events(12, seed=seed)
    if few:  # the first epoch keeps two rangings
        i0 = next(i for i, e in enumerate(ev) if e[0] == "toa")
        r = ev[i0][2].copy()
        keep = np.flatnonzero(r > 0)[:2]
        r2 = np.zeros_like(r)
        r2[keep] = r[keep]
        ev[i0] = ("toa", ev[i0][1], r2)
    f = R.RefK8(0.5, 0.3, 0.5, None, xml={"kfpos_tag": XML_FH1} if fh else None)
    xs, Ps, rcs, tz = [], [], [], []
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
        x, P = f.state()
        xs.append(x); Ps.append(P); rcs.append(rc); tz.append(f.tag_z())
    payload = np.zeros((len(ev), 24))
    for i, e in enumerate(ev):
        payload[i, :len(e[2])] = e[2]
    return dict(anchors=anc, kinds=np.array([e[0] for e in ev]), dts=np.array([e[1] for e in ev]), payload=payload,
                x=np.array(xs), P=np.array(Ps), rc=np.array(rcs), tagz=np.array(tz), fh=fh)


def run_t9(few, seed):
    anc = synth.anchors_for(8)
    T = 25
    truth = synth.truth_lissajous(1, T, 0.1, seed=seed)
    r = synth.ranges_mm(truth[1:], anc, seed=seed + 1, p_missing=0.1).astype(np.float64)[:, :, 0] / 1000
    if few:
        r[0, 3:] = 0
    f = R.RefT9(0.5, 0.5, None)
    xs, Ps, rcs = [], [], []
    for t in range(T):
        rcs.append(f.new_toa(0.1, r[t], anc, np.full(8, 0.01)))
        x, P = f.state()
        xs.append(x); Ps.append(P)
    return dict(anchors=anc, ranges=r, x=np.array(xs), P=np.array(Ps), rc=np.array(rcs))


def main():
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    for fh in (0, 1):
        for few in (0, 1):
            name = f"k8_fh{fh}_{'few' if few else 'normal'}"
            for k, v in run_k8(fh, few, seed=200 + 10 * fh + few).items():
                out[f"{name}/{k}"] = v
    for few in (0, 1):
        name = f"t9_{'few' if few else 'normal'}"
        for k, v in run_t9(few, seed=300 + few).items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "mlinit.npz"), accel_noise=0.5, init_angle=0.3, jolt=0.5, err=0.01,
                        xml_tag_fh1=XML_FH1, **out)
    for k in sorted(out):
        if k.endswith("/rc"):
            print(k, "throws at", np.flatnonzero(out[k]).tolist())
    print("written", os.path.join(OUT, "mlinit.npz"))


if __name__ == "__main__":
    main()
