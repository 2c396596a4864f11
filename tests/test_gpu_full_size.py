"""GPU parity at the BASELINE.json FULL sizes.  The filters of a batch are independent, so a random
sample of them, replayed through the oracle on the same inputs, checks the whole-size run directly
(on top of the size-independent properties: every filter updated exactly T times, no status bits,
statistics identical for every launch geometry)."""
import numpy as np
import pytest

from roskfpos_b200 import synth
from tests.util import REL_TOL, rel_err_cov, rel_err_state

pytestmark = pytest.mark.gpu


def test_t6_bench_size_sampled_against_the_oracle(kflib, oracle):
    """The bench workload: 1,048,576 filters x 100 epochs x 8 anchors (3.4 GB of int32 rangings)."""
    import torch
    from roskfpos_b200.batch import Batch
    dev = torch.device("cuda", 0)
    N, T, M = 1 << 20, 100, 8
    anc = synth.anchors_for(M)
    ranges, x0, truth_end = synth.device_ranges_mm(N, T, anc, 0.1, dev, seed=synth.SEED)
    x0f = torch.zeros((6, N), device=dev, dtype=torch.float64)
    x0f[:3] = x0
    idx = np.sort(np.random.default_rng(0).choice(N, 1536, replace=False))
    tidx = torch.as_tensor(idx, device=dev)
    with Batch(kflib.MODEL_T6, N, anchors=anc, accel_noise=0.5) as b:
        b.set_state(x0f)
        b.replay_toa(0.1, ranges, err=0.01)
        xg = torch.empty((6, N), device=dev, dtype=torch.float64)
        Pg = torch.empty((36, N), device=dev, dtype=torch.float64)
        st = torch.empty(N, device=dev, dtype=torch.int32)
        b.get_state_into(xg, Pg, st)
        cnt = b.counters()
        stats = b.error_stats(truth_end)
    assert cnt["updates"] == N * T and cnt["bad"] == 0
    assert int((st & ~32).abs().sum().item()) == 0
    assert stats[2] == N and np.sqrt(stats[0] / N) < 0.5  # RMSE of a working filter (0.1 m ranging noise)
    ref = oracle.t6_replay(x0[:, tidx].cpu().numpy(), None, ranges[:, :, tidx].cpu().numpy(), anc, 0.1, 0.01)
    assert rel_err_state(xg[:3, tidx].cpu().numpy(), ref["x"]) < REL_TOL
    assert rel_err_cov(Pg[:, tidx].cpu().numpy(), ref["P"]) < REL_TOL
    # iteration counters of the sample: identical work
    with Batch(kflib.MODEL_T6, len(idx), anchors=anc, accel_noise=0.5) as b2:
        b2.set_state(x0f[:, tidx].contiguous())
        b2.replay_toa(0.1, ranges[:, :, tidx].contiguous(), err=0.01)
        c2 = b2.counters()
        x2, P2, _ = b2.get_state()
    assert [c2["ml_iters"], c2["cost_evals"], c2["gain_evals"]] == list(ref["counters"][:3])
    # a filter's result does not depend on the batch it is in (bit-identical)
    assert np.array_equal(x2, xg[:, tidx].cpu().numpy()) and np.array_equal(P2, Pg[:, tidx].cpu().numpy())


@pytest.mark.parametrize("full", [False, True])
def test_k8_million_filters_sampled_against_the_oracle(kflib, oracle, full):
    """BASELINE configs 3 / 5 geometry at 1,048,576 filters (5 macro-steps of the event schedule)."""
    import torch
    from roskfpos_b200.batch import Batch
    dev = torch.device("cuda", 0)
    N = 1 << 20
    anc = synth.anchors_for(8)
    w = synth.k8_workload(N, 5, anc, seed=synth.SEED + 8, full=full, xp=torch, device=dev)
    idx = np.sort(np.random.default_rng(1).choice(N, 768, replace=False))
    tidx = torch.as_tensor(idx, device=dev)
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(w["x0"])
        b.replay_events(w["events"], ranges=w["ranges"], sensors=w["sensors"], err=0.01)
        xg = torch.empty((8, N), device=dev, dtype=torch.float64)
        Pg = torch.empty((64, N), device=dev, dtype=torch.float64)
        st = torch.empty(N, device=dev, dtype=torch.int32)
        b.get_state_into(xg, Pg, st)
        cnt = b.counters()
    assert cnt["updates"] == N * w["n_events"] and cnt["bad"] == 0
    assert int((st & ~32).abs().sum().item()) == 0
    cfg = oracle.k8_cfg(0.5, 0.5, **synth.K8_ORACLE_CFG)
    ref = oracle.k8_replay(w["x0"][:, tidx].cpu().numpy(), None, w["events"], w["ranges"][:, :, tidx].cpu().numpy(),
                           w["sensors"][:, tidx].cpu().numpy(), anc, 0.01, cfg)
    assert rel_err_state(xg[:, tidx].cpu().numpy(), ref["x"]) < REL_TOL
    assert rel_err_cov(Pg[:, tidx].cpu().numpy(), ref["P"]) < REL_TOL
