#!/usr/bin/env python
"""Golden vectors for the WHOLE multi-sensor chain of the reference node, produced by the REFERENCE ITSELF
(oracle/_ref/libkfref.so: PosGenerator + KalmanFilter, unmodified, on the harness's fake clock):
several tags, each with its own raw ranging log AND its own PX4Flow / IMU / magnetometer / compass
messages arriving in its own order, fed to the node message by message (Posgenerator.cpp:92-140); the
report the node would publish at the end (publishPositionReport) is recorded.

tests/golden/node_k8.npz: SoA message logs padded to a common length --
  ranging log: anchor u8 / seq u8 / range_mm i32 / t f64 / err f64, [L][N] (anchor 0xFF = padding)
  sensor logs: px4_t [L1][N], px4 [L1][5][N]; imu_t, imu [L2][3][N] (angular velocity z, acceleration x, y);
               mag_t, mag [L3][2][N]; compass_t, compass [L4][1][N]   (time -1 = padding)
  per tag: x0 [2][N] + start angle, t_report [N], pose [13][N], cov [36][N], n_epochs [N]
The tests push these logs through assembler -> stream merger -> ragged replay -> getPose report, on the CPU
oracle (tests/test_oracle_golden.py) and through the C ABI on the GPU (tests/test_gpu_pipeline.py).
Run:  python tests/golden/make_golden_node.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_py as R  # noqa: E402
from roskfpos_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
M = 8
CAV = np.diag([1e-3, 1e-3, 2e-3]).ravel()
CAC = np.array([[4e-3, 1e-4, 0], [1e-4, 5e-3, 0], [0, 0, 6e-3]]).ravel()


def one_tag(seed, T):
    """Returns the tag's logs and the node's final report.  Message times sit on a 0.1 ms grid with
    kind-specific offsets, so that no two callbacks (timer included) share a time stamp."""
    rng = np.random.default_rng(seed)
    anc = synth.anchors_for(M)
    truth = synth.truth_lissajous(1, T, 0.1, seed=seed + 1)
    p0 = truth[0][:, 0]
    ang0 = 0.3
    pg = R.RefPosGenerator(anc, algorithm=4, start=[p0[0], p0[1], 0.0], use_start=True, start_angle=ang0)
    log = dict(a=[], s=[], r=[], t=[], e=[])
    sens = {1: ([], []), 2: ([], []), 3: ([], []), 4: ([], [])}
    tt = 0.0
    for q in range(T):
        n_imu = int(rng.integers(2, 5))
        for _ in range(n_imu):
            tt = round(tt + float(rng.choice([0.011, 0.017, 0.021])) + 0.0003, 4)
            w = [0.0, 0.0, rng.normal(0.1, 0.05)]
            a = [rng.normal(0, 0.3), rng.normal(0, 0.3), 9.8]
            assert pg.sensor(2, tt, np.concatenate([w, CAV, a, CAC])) == 0
            sens[2][0].append(tt); sens[2][1].append([w[2], a[0], a[1]])
        if rng.random() < 0.8:  # a PX4Flow frame, sometimes of quality 0 (skipped before the clock is read)
            tt = round(tt + 0.0071, 4)
            qual = 0.0 if rng.random() < 0.25 else 200.0
            px = np.array([np.float32(rng.normal(0, 0.002)), np.float32(rng.normal(0, 0.002)),
                           np.float32(rng.normal(0, 0.001)), 33333.0, qual], dtype=np.float64)
            assert pg.sensor(1, tt, px) == 0
            sens[1][0].append(tt); sens[1][1].append(px)
        u = rng.random()
        if u < 0.45:
            tt = round(tt + 0.0052, 4)
            c = float(rng.uniform(-3, 3))
            assert pg.sensor(4, tt, [c]) == 0
            sens[4][0].append(tt); sens[4][1].append([c])
        elif u < 0.9:
            tt = round(tt + 0.0052, 4)
            mg = np.array([np.cos(0.3 + 0.01 * q), np.sin(0.3 + 0.01 * q), 0.1])
            assert pg.sensor(3, tt, np.concatenate([mg, np.zeros(9)])) == 0
            sens[3][0].append(tt); sens[3][1].append(mg[:2])
        for k in range(M):  # the ranging burst of sequence number q
            if rng.random() < 0.12:
                continue
            tt = round(tt + 0.001, 4)
            d = np.linalg.norm(anc[k] - truth[q + 1][:, 0]) + rng.normal(0, 0.05)
            msg = (k, q % 256, int(d * 1000), tt, float(np.float32(0.01 + 0.02 * rng.random())))
            for key, v in zip("asrte", msg):
                log[key].append(v)
            pg.feed([msg[0]], [msg[2]], [msg[1]], [msg[3]], err=[msg[4]], flush_tail=False)
        if q + 1 < T and rng.random() < 0.3:  # a silence longer than the 50 ms timer before the next cycle
            tt = round(tt + 0.0613, 4)
    assert pg.errors() == 0
    t_rep = tt + 0.0137
    rc, pose, cov = pg.report(t_rep)
    assert rc == 0
    n_ep = len(pg.epoch_times(8 * T))
    return dict(log=log, sens=sens, x0=p0[:2], ang0=ang0, t_report=t_rep, pose=pose, cov=cov, n_epochs=n_ep)


def main():
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    tags = [one_tag(500 + 7 * i, 22 + 3 * (i % 3)) for i in range(6)]
    N = len(tags)
    L = max(len(t["log"]["a"]) for t in tags)
    out = dict(anchors=synth.anchors_for(M), anchor=np.full((L, N), 0xFF, np.uint8), seq=np.zeros((L, N), np.uint8),
               range_mm=np.zeros((L, N), np.int32), t=np.zeros((L, N)), err=np.zeros((L, N)))
    for f, tg in enumerate(tags):
        n = len(tg["log"]["a"])
        out["anchor"][:n, f] = tg["log"]["a"]; out["seq"][:n, f] = tg["log"]["s"]
        out["range_mm"][:n, f] = tg["log"]["r"]; out["t"][:n, f] = tg["log"]["t"]; out["err"][:n, f] = tg["log"]["e"]
        out["t"][n:, f] = tg["log"]["t"][-1]
    for kind, name, rows in ((1, "px4", 5), (2, "imu", 3), (3, "mag", 2), (4, "compass", 1)):
        Lk = max(1, max(len(t["sens"][kind][0]) for t in tags))
        tk = np.full((Lk, N), -1.0); pk = np.zeros((Lk, rows, N))
        for f, tg in enumerate(tags):
            n = len(tg["sens"][kind][0])
            if n:
                tk[:n, f] = tg["sens"][kind][0]
                pk[:n, :, f] = np.array(tg["sens"][kind][1])
        out[name + "_t"] = tk; out[name] = pk
    out["x0"] = np.array([t["x0"] for t in tags]).T
    out["ang0"] = np.array([t["ang0"] for t in tags])
    out["t_report"] = np.array([t["t_report"] for t in tags])
    out["pose"] = np.array([t["pose"] for t in tags]).T
    out["cov"] = np.array([t["cov"] for t in tags]).T
    out["n_epochs"] = np.array([t["n_epochs"] for t in tags])
    out["imu_aux"] = np.array([CAC[0], CAC[1], CAC[3], CAC[4], CAV[8]])
    np.savez_compressed(os.path.join(OUT, "node_k8.npz"), accel_noise=0.5, jolt=0.5, **out)
    print("written", os.path.join(OUT, "node_k8.npz"), "epochs per tag", out["n_epochs"],
          "px4 frames of quality 0:", int((out["px4"][:, 4, :] == 0).sum() - (out["px4_t"] < 0).sum()))


if __name__ == "__main__":
    main()
