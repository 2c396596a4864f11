import os, sys, torch
sys.path.insert(0, os.getcwd())
from roskfpos_b200 import lib as L, synth
from roskfpos_b200.batch import Batch
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
N, n_macro = 1 << 20, 5
anc = synth.anchors_for(8)
for full in (False, True):
    w = synth.k8_workload(N, n_macro, anc, seed=synth.SEED + 8, full=full, xp=torch, device=dev)
    for name, kw in (("tuned", {}), ("general (ml_initial_position, 3-D)", dict(ml_initial_position=1))):
        with Batch(L.MODEL_K8, N, device=0, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5, **kw) as b:
            ms = []
            for k in range(4):
                b.set_state(w["x0"], None, stream=stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                b.replay_events(w["events"], ranges=w["ranges"], sensors=w["sensors"], err=0.01, stream=stream)
                e1.record(stream); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
            t = min(ms[1:])
            print(f"{'config5' if full else 'config3'} {name:40s} {t:8.3f} ms {N*len(w['events'])/t/1e6:8.3f} G events/s", flush=True)
# ---- the ragged replay (per-filter time steps) through the general instantiation, same schedule, every filter present
    import numpy as np
    dtf = torch.tensor(np.array([[e[1]] * 1 for e in w["events"]]), device=dev, dtype=torch.float64).repeat(1, N).contiguous()
    with Batch(L.MODEL_K8, N, device=0, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        ms = []
        for k in range(4):
            b.set_state(w["x0"], None, stream=stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            b.replay_events(w["events"], ranges=w["ranges"], sensors=w["sensors"], err=0.01, stream=stream, dt_per_filter=dtf)
            e1.record(stream); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        t = min(ms[1:])
        print(f"{'config5' if full else 'config3'} {'ragged replay (general instantiation)':40s} {t:8.3f} ms {N*len(w['events'])/t/1e6:8.3f} G events/s", flush=True)
