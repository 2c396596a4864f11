#!/usr/bin/env python
"""Times the MLLocation variants at the BASELINE config-4 sizes (16 anchors): IgnoreN at 1 Mi and 4 Mi epochs,
BestGroup 2-D (best 3 of 16) and 3-D (best 4 of 16) on the exact 4 x 4 grid (where most epochs end at the first
collinear triple, as the reference's exception does) and on a jittered grid (every epoch enumerates all subsets).

    [KFPOS_B200_SO=build/variants/libX.so] python profiles/ml_variants.py [ign] [best2] [best3]
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
which = sys.argv[1:] or ["ign", "best2", "best3"]
REPS = 3


def timed(name, fn, units, reps=REPS):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{name:28s} {ms:9.3f} ms  {units / ms / 1e3:11.3f} M epochs/s", flush=True)


anc16 = synth.anchors_for(16)
jit = anc16.copy()
jit[:, :2] += np.random.default_rng(3).uniform(-0.3, 0.3, size=(16, 2))
if "ign" in which:
    for Nm in (1 << 20, 1 << 22):
        rm, _, _ = synth.device_ranges_mm(Nm, 1, anc16, 0.1, dev, seed=synth.SEED + 6)
        rm = rm[0].contiguous()
        o4 = dict(pos=torch.empty((3, Nm), device=dev, dtype=torch.float64), cov=None, iters=None,
                  sel=torch.empty((2, Nm), device=dev, dtype=torch.int32), status=None)
        with Batch(L.MODEL_ML, Nm, device=0, anchors=anc16, use2d=0, variant=1, num_ignored_rangings=2) as b:
            timed(f"ignore2 3-D {Nm >> 20} Mi", lambda: b.ml_solve(rm, err=0.01, out=o4, stream=stream), Nm)
        del rm, o4
if "best2" in which or "best3" in which:
    Nb = int(os.environ.get("KF_NB", 1 << 17))
    for nm, anc in (("grid", anc16), ("jittered", jit)):
        rb, _, _ = synth.device_ranges_mm(Nb, 1, anc, 0.1, dev, seed=synth.SEED + 6)
        rb = rb[0].contiguous()
        ob = dict(pos=torch.empty((3, Nb), device=dev, dtype=torch.float64), cov=None,
                  iters=torch.empty(Nb, device=dev, dtype=torch.int32),
                  sel=torch.empty((2, Nb), device=dev, dtype=torch.int32),
                  status=torch.empty(Nb, device=dev, dtype=torch.int32))
        for use2d in [u for u, w in ((1, "best2"), (0, "best3")) if w in which]:
            with Batch(L.MODEL_ML, Nb, device=0, anchors=anc, use2d=use2d, variant=2,
                       ml_start=[1.0, 1.0, 1.0 if use2d else 4.0]) as b:
                timed(f"best {'3 2-D' if use2d else '4 3-D'} {nm}", lambda: b.ml_solve(rb, err=0.01, out=ob, stream=stream),
                      Nb, reps=1)
                it = ob["iters"].double()
                print(f"    mean Newton iterations per epoch {it.mean().item():.1f}, status!=0: "
                      f"{(ob['status'] != 0).double().mean().item():.3f}, checksum sel {int(ob['sel'].sum().item())} "
                      f"iters {int(ob['iters'].sum().item())}", flush=True)
