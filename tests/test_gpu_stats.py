"""GPU: the Monte Carlo error statistics -- fused into the replay kernels, geometry-independent summation
order, and the one collective of a sharded run (kfpos_stats_allreduce) through a real NCCL communicator."""
import numpy as np
import pytest

from roskfpos_b200 import synth

pytestmark = pytest.mark.gpu


def _t6_case(N, T=6, seed=11):
    anc = synth.anchors_for(8)
    truth = synth.truth_lissajous(N, T, 0.1, seed=seed)
    r = synth.ranges_mm(truth[1:], anc, seed=seed + 1)
    return anc, truth, r


@pytest.mark.parametrize("N", [4096, 1000, 129])
def test_fused_partials_equal_the_separate_reduction(kflib, N):
    """kfpos_batch_set_truth: the replay kernel leaves the per-block partials; error_stats(NULL) only folds
    them.  Bit-identical to the un-fused path, for T6, K8 and T9, full and ragged last blocks."""
    from roskfpos_b200.batch import Batch
    anc, truth, r = _t6_case(N)
    with Batch(kflib.MODEL_T6, N, anchors=anc, accel_noise=0.5) as b:
        b.set_state(truth[0])
        b.replay_toa(0.1, r)
        plain = b.error_stats(truth[-1])
        x, _, _ = b.get_state(want_P=False)
        b.set_truth(truth[-1])
        b.set_state(truth[0])
        b.replay_toa(0.1, r)
        fused = b.error_stats()
        again = b.error_stats()            # partials consumed: recomputed from the registered truth
        b.step_toa(0.1, r[0])              # single step: state changed, partials of this launch
        stepped = b.error_stats()
        ref_stepped = b.error_stats(truth[-1])
        alone = b.stats_allreduce(None)
    e2 = ((x[:3] - truth[-1]) ** 2).sum(axis=0)
    assert plain[2] == N and abs(plain[0] - e2.sum()) <= 1e-12 * e2.sum()
    assert np.array_equal(plain, fused) and np.array_equal(plain, again)
    assert np.array_equal(stepped, ref_stepped) and not np.array_equal(stepped, plain)
    assert np.array_equal(alone[:4], stepped) and alone[4] == np.sqrt(stepped[0] / N)
    # K8 (planar: z = tag height) and T9 through their event kernels
    w = synth.k8_workload(N, 2, anc, seed=5, full=True)
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(w["x0"])
        b.replay_events(w["events"], ranges=w["ranges"], sensors=w["sensors"], err=0.01)
        plain = b.error_stats(w["truth_end"])
        b.set_truth(w["truth_end"])
        b.set_state(w["x0"])
        b.replay_events(w["events"], ranges=w["ranges"], sensors=w["sensors"], err=0.01)
        assert np.array_equal(plain, b.error_stats())
    x9 = np.zeros((9, N)); x9[:3] = truth[0]
    with Batch(kflib.MODEL_T9, N, anchors=anc, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(x9)
        b.replay_toa(0.1, r)
        plain = b.error_stats(truth[-1])
        b.set_truth(truth[-1])
        b.set_state(x9)
        b.replay_toa(0.1, r)
        assert np.array_equal(plain, b.error_stats())


def test_aligned_shards_are_subtrees_of_the_global_tree(kflib):
    """The summation order is a pairwise tree over the filter index: shards of N / G filters (a power of two,
    multiple of 128) reduce to the value of the corresponding subtree, so adding the shard results pairwise
    in rank order reproduces the single-batch result BIT FOR BIT (what kfpos_stats_allreduce does)."""
    from roskfpos_b200.batch import Batch
    N = 8192
    anc, truth, r = _t6_case(N, seed=21)

    def stats(lo, hi):
        with Batch(kflib.MODEL_T6, hi - lo, anchors=anc, accel_noise=0.5) as b:
            b.set_truth(np.ascontiguousarray(truth[-1][:, lo:hi]))
            b.set_state(np.ascontiguousarray(truth[0][:, lo:hi]))
            b.replay_toa(0.1, np.ascontiguousarray(r[:, :, lo:hi]))
            return b.error_stats()
    whole = stats(0, N)
    for G in (2, 4, 8):
        parts = [stats(g * N // G, (g + 1) * N // G) for g in range(G)]
        while len(parts) > 1:
            parts = [parts[i] + parts[i + 1] for i in range(0, len(parts), 2)]
        assert np.array_equal(parts[0], whole), G


def test_stats_allreduce_through_nccl(kflib):
    """A real ncclComm_t (one rank: this GPU) through the C ABI: the collective path (ncclAllGather + rank
    tree) returns the batch's own statistics.  The multi-rank case runs in bench.py under torchrun."""
    from roskfpos_b200.batch import Batch
    from roskfpos_b200.shard import nccl_comm
    N = 2048
    anc, truth, r = _t6_case(N, seed=31)
    comm = nccl_comm(0, 1, 0)
    try:
        with Batch(kflib.MODEL_T6, N, anchors=anc, accel_noise=0.5) as b:
            b.set_truth(truth[-1])
            b.set_state(truth[0])
            b.replay_toa(0.1, r)
            b.error_stats(readback=False)       # enqueue only
            got = b.stats_allreduce(comm)        # folds again from the registered truth, then the collective
            ref = b.error_stats(truth[-1])
        assert np.array_equal(got[:4], ref)
        assert got[4] == np.sqrt(ref[0] / ref[2]) and got[5] == np.sqrt(ref[1] / ref[2])
    finally:
        comm.close()
