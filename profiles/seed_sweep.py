#!/usr/bin/env python
"""Input sensitivity of the headline kernel: the T6 replay of bench.py for the eight per-rank seeds
(SEED + 1000 r) on ONE GPU, kernel time per seed and the count of inner Newton solves that ran to the
reference's 10000-iteration cap (ML.cpp:165) -- the launch-tail stragglers of VERDICT r01 item 2.

    python profiles/seed_sweep.py [--tsteps 100] [--filters 1048576] [--reps 5] > profiles/r02_seed_sweep.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roskfpos_b200 import lib as L, synth  # noqa: E402
from roskfpos_b200.batch import Batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tsteps", type=int, default=100)
    ap.add_argument("--filters", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--seeds", type=int, default=8)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    N, T = a.filters, a.tsteps
    anc = synth.anchors_for(8)
    stream = torch.cuda.current_stream()
    rows = []
    for r in range(a.seeds):
        ranges, x0, truth_end = synth.device_ranges_mm(N, T, anc, 0.1, dev, seed=synth.SEED + 1000 * r)
        x0_full = torch.zeros((6, N), device=dev, dtype=torch.float64)
        x0_full[:3] = x0
        with Batch(L.MODEL_T6, N, device=0, anchors=anc, accel_noise=0.5) as b:
            ms = []
            for k in range(a.reps + 2):
                b.set_state(x0_full, None, stream=stream)
                if k == 2:
                    b.counters(reset=True)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                b.replay_toa(0.1, ranges, err=0.01, stream=stream)
                e1.record(stream)
                torch.cuda.synchronize()
                if k >= 2:
                    ms.append(e0.elapsed_time(e1))
            c = b.counters(reset=True)
        rows.append({"seed": synth.SEED + 1000 * r, "kernel_ms_min": float(np.min(ms)), "kernel_ms_mean": float(np.mean(ms)),
                     "ml_capped_per_launch": c["ml_capped"] / a.reps, "ml_cycles_skipped_per_launch": c["ml_cycles"] / a.reps,
                     "mean_ml_iters": c["ml_iters"] / max(c["updates"], 1)})
        del ranges
        torch.cuda.empty_cache()
    t = np.array([x["kernel_ms_mean"] for x in rows])
    print(json.dumps({"filters": N, "epochs": T, "reps": a.reps, "rows": rows,
                      "spread": float((t.max() - t.min()) / t.min())}, indent=1))


if __name__ == "__main__":
    main()
