"""GPU parity: the epoch assembler (kfpos_assemble_epochs) against the oracle, bit-exact, and the
assembled logs through the per-filter-dt replay (kfpos_batch_replay_epochs)."""
import numpy as np
import pytest

from roskfpos_b200 import synth

pytestmark = pytest.mark.gpu


def make_logs(N, n_seq, M, seed, p_drop=0.2, p_gap=0.05, p_dup=0.05):
    """N ragged ranging logs: per sequence number a random subset of the anchors reports in random
    order, sometimes twice, 1-3 ms apart; sometimes a pause > 50 ms (timer report) in the middle of a
    sequence; sequence numbers wrap at 256.  Padded with anchor 0xFF to a common length."""
    rng = np.random.default_rng(seed)
    L = n_seq * (M + 2)
    anchor = np.full((L, N), 0xFF, dtype=np.uint8)
    seq = np.zeros((L, N), dtype=np.uint8)
    rmm = np.zeros((L, N), dtype=np.int32)
    err = np.zeros((L, N))
    t = np.zeros((L, N))
    for f in range(N):
        i, now, s0 = 0, float(rng.uniform(0, 1)), int(rng.integers(0, 256))
        for k in range(n_seq):
            order = rng.permutation(M)
            order = order[rng.random(M) >= p_drop]
            if rng.random() < p_dup and len(order):
                order = np.append(order, order[0])
            for a in order:
                now += float(rng.uniform(0.001, 0.003))
                if rng.random() < p_gap:
                    now += 0.06
                anchor[i, f], seq[i, f] = a, (s0 + k) % 256
                rmm[i, f] = int(rng.integers(500, 15000))
                err[i, f] = 0.0 if rng.random() < 0.3 else float(rng.uniform(0.005, 0.05))
                t[i, f] = now
                i += 1
            now += float(rng.uniform(0.02, 0.09))
    return anchor, seq, rmm, err, t


@pytest.mark.parametrize("fix", [False, True])
@pytest.mark.parametrize("N,n_seq,M", [(300, 40, 8), (64, 600, 4), (1000, 12, 16)])
def test_assembler_bit_exact(kflib, oracle, N, n_seq, M, fix):
    from roskfpos_b200.batch import assemble_epochs
    anchor, seq, rmm, err, t = make_logs(N, n_seq, M, seed=N + n_seq)
    T = 4 * n_seq + 8  # in-sequence pauses add timer reports
    ref = oracle.assemble(anchor, seq, rmm, t, M, T, err=err, fix_b12=fix)
    got = assemble_epochs(anchor, seq, rmm, t, M, T, err=err, fix_row_clear=fix)
    assert np.array_equal(got["n_epochs"], ref["n_epochs"]) and ref["n_epochs"].max() <= T
    assert np.array_equal(got["ranges"], ref["ranges"])
    assert np.array_equal(got["err"], ref["err"])
    assert np.array_equal(got["dt"], ref["dt"])
    if n_seq > 256 and not fix:  # the wrap makes stale slots visible (App. B-12)
        fixed = oracle.assemble(anchor, seq, rmm, t, M, T, err=err, fix_b12=True)
        assert not np.array_equal(fixed["ranges"], ref["ranges"])


def test_assembler_truncation_empty_and_device_tensors(kflib, oracle):
    import torch
    from roskfpos_b200.batch import assemble_epochs
    N, n_seq, M = 257, 20, 8
    anchor, seq, rmm, err, t = make_logs(N, n_seq, M, seed=5)
    anchor[:, :3] = 0xFF  # three empty logs
    T = 7                 # fewer slots than reports
    ref = oracle.assemble(anchor, seq, rmm, t, M, T, err=None)
    dev = torch.device("cuda", 0)
    d = [torch.as_tensor(a, device=dev) for a in (anchor, seq, rmm, t)]
    out = dict(ranges=torch.empty((T, M, N), dtype=torch.int32, device=dev), err=None,
               dt=torch.empty((T, N), dtype=torch.float64, device=dev),
               n_epochs=torch.empty(N, dtype=torch.int32, device=dev))
    assemble_epochs(d[0], d[1], d[2], d[3], M, T, out=out)
    assert np.array_equal(out["n_epochs"].cpu().numpy(), ref["n_epochs"])
    assert (ref["n_epochs"][:3] == 0).all() and (ref["n_epochs"][3:] > T).all()
    assert np.array_equal(out["ranges"].cpu().numpy(), ref["ranges"])
    assert np.array_equal(out["dt"].cpu().numpy(), ref["dt"])


def test_assembled_logs_through_the_ragged_replay(kflib, oracle):
    """Raw logs -> assembler -> T6 replay with per-filter dt == the oracle filter stepped through each
    log's own epochs (skipping the epochs a log does not have)."""
    from roskfpos_b200.batch import Batch, assemble_epochs
    N, n_seq, M = 200, 25, 8
    anc = synth.anchors_for(M)
    rng = np.random.default_rng(11)
    anchor, seq, rmm, err, t = make_logs(N, n_seq, M, seed=12, p_drop=0.1)
    # plausible ranges: distance to a fixed point per log + noise
    p0 = np.stack([rng.uniform(2, 8, N), rng.uniform(2, 8, N), np.full(N, 1.0)])
    d = np.sqrt(((p0[None] - anc[:, :, None]) ** 2).sum(axis=1))  # [M][N]
    a_idx = np.where(anchor == 0xFF, 0, anchor).astype(np.int64)
    rmm = np.floor((np.take_along_axis(d, a_idx, axis=0) + 0.05 * rng.normal(size=anchor.shape)) * 1000).astype(np.int32)
    T = 4 * n_seq + 8  # in-sequence pauses add timer reports
    ep = assemble_epochs(anchor, seq, rmm, t, M, T, fix_row_clear=True)
    with Batch(kflib.MODEL_T6, N, anchors=anc, accel_noise=0.5) as b:
        b.set_state(p0)
        traj = b.replay_epochs(ep["dt"], ep["ranges"], err=0.01, want_traj=True)
        x, P, st = b.get_state()
        cnt = b.counters()
    assert cnt["updates"] == ep["n_epochs"].sum()
    for f in range(0, N, 7):
        n = int(ep["n_epochs"][f])
        ref = oracle.t6_replay(p0[:, f:f + 1], None, ep["ranges"][:n, :, f:f + 1], anc, ep["dt"][:n, f], 0.01)
        assert np.abs(x[:3, f] - ref["x"][:, 0]).max() < 1e-9
        assert np.abs(P[:, f] - ref["P"][:, 0]).max() <= 1e-9 * np.abs(ref["P"]).max()
        assert np.array_equal(traj[n:, :, f], np.broadcast_to(traj[n - 1, :, f], (T - n, 3)))
