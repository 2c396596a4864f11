import numpy as np, sys, time
sys.path.insert(0, '.')
from roskfpos_b200 import synth, lib as kflib
from roskfpos_b200.batch import Batch
from oracle import oracle_py as oracle
N, T, m = 200000, 20, 8
anc = synth.anchors_for(m)
truth = synth.truth_lissajous(N, T, 0.1, seed=5)
for name, kw_o, kw_g, nlos in (("plain", {}, {}, 0.0), ("v1", dict(variant=1, n_ignore=2), dict(variant=1, num_ignored_rangings=2), 0.15)):
    r = synth.ranges_mm(truth[1:], anc, seed=6, p_missing=0.1, p_nlos=nlos)
    ref = oracle.t6_replay(truth[0], None, r, anc, 0.1, 0.01, want_traj=True, **kw_o)
    with Batch(kflib.MODEL_T6, N, anchors=anc, accel_noise=0.5, **kw_g) as b:
        b.set_state(truth[0])
        traj, _ = b.replay_toa(0.1, r, err=0.01, want_traj=True)
    d = np.abs(traj - ref["traj"]).max(axis=1)   # [T][N]
    first_bad = (d > 1e-9)
    units_bad = first_bad.any(axis=0).sum()
    print(name, "filters with any step > 1e-9:", int(units_bad), "of", N, " max err", float(d.max()),
          " median err", float(np.median(d)), " p99.9", float(np.quantile(d, 0.999)))
