#!/usr/bin/env python
"""T9 (KalmanFilterTOAIMU) replay time per event kind: the kfpos_toa_imu pattern (10 accelerometer samples per
ranging epoch), 1 Mi filters.   python profiles/t9_split.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roskfpos_b200 import lib as L, synth  # noqa: E402
from roskfpos_b200.batch import Batch  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()
N, n_macro, M = int(os.environ.get("KF_N", 1 << 20)), 5, 8
anc = synth.anchors_for(M)
ranges, x0, _ = synth.device_ranges_mm(N, n_macro, anc, 0.1, dev, seed=synth.SEED + 3)
g = torch.Generator(device=dev)
g.manual_seed(5)
acc = 0.05 * torch.randn((n_macro * 10 * 3, N), generator=g, device=dev, dtype=torch.float64)
x0f = torch.zeros((9, N), device=dev, dtype=torch.float64)
x0f[:3] = x0
cov = list(np.diag([4e-3, 5e-3, 6e-3]).ravel())
events = []
for q in range(n_macro):
    for k in range(10):
        events.append((synth.EV_IMU, 0.009, (q * 10 + k) * 3, cov))
    events.append((synth.EV_TOA, 0.01, q * M, None))
scheds = {"all": events, "imu only": [e for e in events if e[0] == synth.EV_IMU],
          "toa only (no imu latched)": [e for e in events if e[0] == synth.EV_TOA]}
with Batch(L.MODEL_T9, N, device=0, anchors=anc, accel_noise=0.5, jolt=0.5) as b:
    for name, ev in scheds.items():
        ms = []
        for k in range(4):
            b.set_state(x0f, None, stream=stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            b.replay_events(ev, ranges=ranges, sensors=acc, err=0.01, stream=stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = min(ms[1:])
        print(f"t9 {name:28s} {len(ev):4d} events {t:8.3f} ms {t / len(ev) * 1e3:8.1f} us/event {N * len(ev) / t / 1e6:8.3f} G events/s", flush=True)
