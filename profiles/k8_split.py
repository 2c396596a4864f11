#!/usr/bin/env python
"""Where the K8 replay spends its time: the config-3 / config-5 schedules and sub-schedules with one event kind
only (kernel time per event, CUDA events, 1 Mi filters).   python profiles/k8_split.py [full]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roskfpos_b200 import lib as L, synth  # noqa: E402
from roskfpos_b200.batch import Batch  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()
N, n_macro = int(os.environ.get("KF_N", 1 << 20)), 5
anc = synth.anchors_for(8)
NAMES = {synth.EV_TOA: "toa", synth.EV_IMU: "imu", synth.EV_PX4: "px4", synth.EV_COMPASS: "compass"}
for full in (False, True):
    w = synth.k8_workload(N, n_macro, anc, seed=synth.SEED + 8, full=full, xp=torch, device=dev)
    scheds = {"all": w["events"]}
    for k, nm in NAMES.items():
        sub = [e for e in w["events"] if e[0] == k]
        if sub:
            scheds[nm + " only"] = sub
    scheds["light (imu+px4)"] = [e for e in w["events"] if e[0] in (synth.EV_IMU, synth.EV_PX4)]
    with Batch(L.MODEL_K8, N, device=0, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        for name, ev in scheds.items():
            ms = []
            for k in range(4):
                b.set_state(w["x0"], None, stream=stream)
                # latch every sensor first so that "toa only" / "compass only" fuse what they fuse in the full schedule
                b.replay_events(w["events"][:len(w["events"]) // n_macro], ranges=w["ranges"], sensors=w["sensors"], err=0.01, stream=stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                b.replay_events(ev, ranges=w["ranges"], sensors=w["sensors"], err=0.01, stream=stream)
                e1.record(stream)
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t = min(ms[1:])
            print(f"{'config5' if full else 'config3'} {name:18s} {len(ev):4d} events {t:8.3f} ms  {t / len(ev) * 1e3:8.1f} us/event  "
                  f"{N * len(ev) / t / 1e6:8.3f} G events/s", flush=True)
