#!/usr/bin/env python
"""Launches each product kernel a few times at a profiling-friendly size (used under
`ncu -k regex:<name>`; prints its own CUDA-event times so the plain run is a record too).

    python profiles/run_kernels.py [t6] [k8] [k8full] [t9] [ml3] [ml2] [mlign] [loo] [asm]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from roskfpos_b200 import lib as L, synth  # noqa: E402
from roskfpos_b200.batch import Batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
which = sys.argv[1:] or ["t6", "k8", "k8full", "t9", "ml3", "ml2", "mlign", "loo", "asm"]
REPS = 3


def timed(name, fn, units):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(REPS):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / REPS
    print(f"{name:8s} {ms:9.3f} ms  {units / ms / 1e6:9.3f} G units/s", flush=True)


anc8, anc16 = synth.anchors_for(8), synth.anchors_for(16)
N, T = int(os.environ.get("KF_N", 1 << 18)), int(os.environ.get("KF_T", 20))
if "t6" in which or "t9" in which:
    r, x0, _ = synth.device_ranges_mm(N, T, anc8, 0.1, dev, seed=synth.SEED)
if "t6" in which:
    x0f = torch.zeros((6, N), device=dev, dtype=torch.float64)
    x0f[:3] = x0
    with Batch(L.MODEL_T6, N, device=0, anchors=anc8, accel_noise=0.5) as b:
        def run():
            b.set_state(x0f, None, stream=stream)
            b.replay_toa(0.1, r, err=0.01, stream=stream)
        timed("t6", run, N * T)
if "t9" in which:
    x0f = torch.zeros((9, N), device=dev, dtype=torch.float64)
    x0f[:3] = x0
    ev = [(L.EV_TOA, 0.1, t * 8, None) for t in range(T)]
    with Batch(L.MODEL_T9, N, device=0, anchors=anc8, accel_noise=0.5, jolt=0.5) as b:
        def run():
            b.set_state(x0f, None, stream=stream)
            b.replay_events(ev, ranges=r, sensors=None, err=0.01, stream=stream)
        timed("t9", run, N * T)
for name, full in (("k8", False), ("k8full", True)):
    if name not in which:
        continue
    w = synth.k8_workload(N, 3, anc8, seed=synth.SEED + 8, full=full, xp=torch, device=dev)
    with Batch(L.MODEL_K8, N, device=0, anchors=anc8, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        def run():
            b.set_state(w["x0"], None, stream=stream)
            b.replay_events(w["events"], ranges=w["ranges"], sensors=w["sensors"], err=0.01, stream=stream)
        timed(name, run, N * w["n_events"])
    del w
if "ml3" in which or "ml2" in which:
    Nm = 1 << 20
    rm, _, _ = synth.device_ranges_mm(Nm, 1, anc8, 0.1, dev, seed=synth.SEED + 5)
    rm = rm[0].contiguous()
    outs = dict(pos=torch.empty((3, Nm), device=dev, dtype=torch.float64),
                cov=torch.empty((9, Nm), device=dev, dtype=torch.float64),
                iters=torch.empty(Nm, device=dev, dtype=torch.int32),
                sel=torch.empty((2, Nm), device=dev, dtype=torch.int32),
                status=torch.empty(Nm, device=dev, dtype=torch.int32))
    for use2d, nm in ((0, "ml3"), (1, "ml2")):
        if nm in which:
            with Batch(L.MODEL_ML, Nm, device=0, anchors=anc8, use2d=use2d,
                       ml_start=[1.0, 1.0, 1.0 if use2d else 4.0]) as b:
                timed(nm, lambda: b.ml_solve(rm, err=0.01, out=outs, stream=stream), Nm)
if "mlign" in which:
    Nm = 1 << 20
    rm, _, _ = synth.device_ranges_mm(Nm, 1, anc16, 0.1, dev, seed=synth.SEED + 6)
    rm = rm[0].contiguous()
    o4 = dict(pos=torch.empty((3, Nm), device=dev, dtype=torch.float64), cov=None, iters=None,
              sel=torch.empty((2, Nm), device=dev, dtype=torch.int32), status=None)
    with Batch(L.MODEL_ML, Nm, device=0, anchors=anc16, use2d=0, variant=1, num_ignored_rangings=2) as b:
        timed("mlign", lambda: b.ml_solve(rm, err=0.01, out=o4, stream=stream), Nm)
if "loo" in which:
    Nl, Tl = 1 << 17, 10
    rl, x0l, _ = synth.device_ranges_mm(Nl, Tl, anc16, 0.1, dev, seed=synth.SEED + 7)
    x0f = torch.zeros((6, Nl), device=dev, dtype=torch.float64)
    x0f[:3] = x0l
    with Batch(L.MODEL_T6, Nl, device=0, anchors=anc16, accel_noise=0.5, ignore_worst_anchor=1,
               ignore_cost_threshold=0.5) as b:
        def run():
            b.set_state(x0f, None, stream=stream)
            b.replay_toa(0.1, rl, err=0.01, stream=stream)
        timed("loo", run, Nl * Tl)
if "asm" in which:
    from roskfpos_b200.batch import assemble_epochs
    Na, n_seq, Ma = 1 << 19, 32, 8
    La = n_seq * Ma
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    a_idx = torch.arange(La, device=dev).remainder(Ma).to(torch.uint8)[:, None].expand(La, Na).contiguous()
    sq = (torch.arange(La, device=dev) // Ma).to(torch.uint8)[:, None].expand(La, Na).contiguous()
    rmm = torch.randint(500, 15000, (La, Na), generator=g, device=dev, dtype=torch.int32)
    tt = (torch.arange(La, device=dev, dtype=torch.float64) * 0.002
          + (torch.arange(La, device=dev) // Ma) * 0.084)[:, None].expand(La, Na).contiguous()
    Ta = n_seq + 2
    oa = dict(ranges=torch.empty((Ta, Ma, Na), dtype=torch.int32, device=dev), err=None,
              dt=torch.empty((Ta, Na), dtype=torch.float64, device=dev),
              n_epochs=torch.empty(Na, dtype=torch.int32, device=dev))
    timed("asm", lambda: assemble_epochs(a_idx, sq, rmm, tt, Ma, Ta, fix_row_clear=True, out=oa, stream=stream), La * Na)
