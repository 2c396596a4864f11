"""GPU parity: MLLocation epochs (variants 0/1/2, 2-D/3-D) through the C ABI vs the oracle."""
import numpy as np
import pytest

from roskfpos_b200 import synth
from tests.util import assert_parity, to_metres, ulp_perturbations

pytestmark = pytest.mark.gpu


def epochs(m, N, seed, **kw):
    anc = synth.anchors_for(m)
    rng = np.random.default_rng(seed)
    truth = np.stack([rng.uniform(1, 9, N), rng.uniform(1, 9, N), np.full(N, 1.0)])
    return anc, truth, synth.ranges_mm(truth, anc, seed=seed + 1, **kw)


def gpu_ml(kflib, anc, r, err=0.01, **cfg):
    from roskfpos_b200.batch import Batch
    with Batch(kflib.MODEL_ML, r.shape[-1], anchors=anc, **cfg) as b:
        return b.ml_solve(r, err=err)


def oracle_ml(oracle, r, anc, err, start, n_random=4, **kw):
    ref = oracle.ml_batch(r, anc, err, start, **kw)
    per = [oracle.ml_batch(p, anc, err, start, **kw) for p in ulp_perturbations(to_metres(r), n_random=n_random)]
    return ref, per


def start_for(use2d):
    return [1.0, 1.0, 1.0 if use2d else 4.0]  # 3-D: the reference's default (1,1,4), PG.cpp:531


@pytest.mark.parametrize("use2d", [0, 1])
@pytest.mark.parametrize("m", [4, 8, 16])
def test_ml_normal(kflib, oracle, m, use2d):
    N = 20000
    anc, truth, r = epochs(m, N, seed=300 + m)
    ref, per = oracle_ml(oracle, r, anc, 0.01, start_for(use2d), use2d=use2d)
    got = gpu_ml(kflib, anc, r, use2d=use2d, ml_start=start_for(use2d))
    rep = assert_parity(got, ref, per, float_keys=("pos",), cov_keys=("cov",), int_keys=("status", "iters"),
                        min_stable=0.98, max_tie_frac=1e-3, what=f"ML m={m} 2d={use2d}")
    print("parity report", m, use2d, rep)


def test_ml_config2_full_size(kflib, oracle):
    """BASELINE config 2: 8 anchors, 1,048,576 epochs, 3-D and 2-D: every epoch within 1e-9."""
    N = 1 << 20
    anc, truth, r = epochs(8, N, seed=777)
    for use2d in (0, 1):
        ref, per = oracle_ml(oracle, r, anc, 0.01, start_for(use2d), use2d=use2d)
        got = gpu_ml(kflib, anc, r, use2d=use2d, ml_start=start_for(use2d))
        rep = assert_parity(got, ref, per, float_keys=("pos",), cov_keys=("cov",), min_stable=0.9999,
                            int_keys=("status", "iters"), max_tie_frac=1e-4, what=f"config2 2d={use2d}")
        print("parity report config2", use2d, rep)


@pytest.mark.parametrize("use2d", [0, 1])
def test_ml_ragged_and_too_few(kflib, oracle, use2d):
    N, m = 8000, 8
    anc, truth, r = epochs(m, N, seed=400, p_missing=0.45)
    r[:, :10] = 0
    err = np.random.default_rng(1).uniform(0.005, 0.05, size=r.shape)
    ref, per = oracle_ml(oracle, r, anc, err, start_for(use2d), use2d=use2d)
    got = gpu_ml(kflib, anc, r, err=err, use2d=use2d, ml_start=start_for(use2d))
    assert (ref["status"] == 2).any()
    assert np.array_equal(got["status"] == 2, ref["status"] == 2)
    few = ref["status"] == 2
    assert np.array_equal(got["pos"][:, few], ref["pos"][:, few])  # start returned untouched
    rep = assert_parity(got, ref, per, float_keys=("pos",), int_keys=("status", "iters"),
                        min_stable=0.85, max_tie_frac=1e-3, what=f"ML ragged 2d={use2d}")
    print("parity report ragged", use2d, rep)


@pytest.mark.parametrize("use2d,n_ignore", [(0, 2), (1, 2), (0, 20)])
def test_ml_ignore_n_selection(kflib, oracle, use2d, n_ignore):
    """Variant 1 (ML.cpp:307-347): the set of dropped anchors is bit-exact."""
    N, m = 20000, 16
    anc, truth, r = epochs(m, N, seed=500, p_nlos=0.15)
    ref, per = oracle_ml(oracle, r, anc, 0.01, start_for(use2d), use2d=use2d, variant=1, n_ignore=n_ignore)
    got = gpu_ml(kflib, anc, r, use2d=use2d, variant=1, num_ignored_rangings=n_ignore,
                 ml_start=start_for(use2d))
    rep = assert_parity(got, ref, per, float_keys=("pos",), int_keys=("status", "sel"),
                        min_stable=0.98, max_tie_frac=0.0, what=f"IgnoreN 2d={use2d} n={n_ignore}")
    print("parity report ignoreN", use2d, n_ignore, rep)


def assert_bit_equal(got, ref, keys, what):
    """EVERY unit, no stability filter: the exact-order solver runs the oracle's IEEE operation sequence."""
    from tests import util
    bad = {}
    for k in keys:
        a, b = np.asarray(got[k]), np.asarray(ref[k])
        eq = (a == b) | ((a != a) & (b != b))
        bad[k] = int((~eq).reshape(-1, eq.shape[-1]).any(axis=0).sum())
    U = int(np.asarray(ref[keys[0]]).shape[-1])
    util.PARITY_REPORT.append(dict(test=util._current_test(), what=what, units=U, stable=U, stable_frac=1.0,
                                   all_units_ok=U - max(bad.values()), stable_violations=max(bad.values()),
                                   stable_int_mismatches=max(bad.values()), unstable_ok=0, worst_stable_float_err=0.0,
                                   tol=0.0, min_stable=1.0, max_tie_frac=0.0, tie_tol=0.0, int_keys=list(keys),
                                   float_keys=[], bit_exact=True, mismatches_per_key=bad))
    assert not any(bad.values()), f"{what}: units that are not bit-identical to the oracle, per output: {bad}"


@pytest.mark.parametrize("use2d,m,best_mode,N", [(1, 8, 0, 3000), (0, 8, 0, 3000), (0, 8, 1, 3000), (1, 5, 0, 3000),
                                                 (1, 16, 0, 1000), (0, 16, 0, 160)])
def test_ml_best_group_selection(kflib, oracle, use2d, m, best_mode, N):
    """Variant 2 (ML.cpp:351-414): subset index in prev_permutation order, slot mask, position,
    covariance and iteration count BIT-IDENTICAL to the oracle on every epoch -- also the chaotic ones
    (3-D subsets of four rangings whose Newton iteration wanders) and on the exact 4 x 4 grid of BASELINE
    config 4, where collinear triples make J^T W^-1 J singular and the scan picks rounding garbage:
    BestGroup epochs are solved by the exact-order kernel (kfpos_exact.cu)."""
    anc, truth, r = epochs(m, N, seed=600 + m, p_nlos=0.15)
    ref = oracle.ml_batch(r, anc, 0.01, start_for(use2d), use2d=use2d, variant=2, best_mode=best_mode)
    got = gpu_ml(kflib, anc, r, use2d=use2d, variant=2, best_mode=best_mode, ml_start=start_for(use2d))
    assert_bit_equal(got, ref, ("sel", "status", "iters", "pos", "cov"), f"BestGroup exact 2d={use2d} m={m} mode={best_mode}")
    assert len(np.unique(ref["sel"][1])) > 5


@pytest.mark.parametrize("use2d,m,best_mode", [(1, 8, 0), (0, 8, 0), (1, 16, 0)])
def test_ml_best_group_fast_formulation(kflib, oracle, use2d, m, best_mode):
    """The same scan with the fast (re-associated) solver, ml_exact_order = -1 -- the code the EKF-side
    variant 2 of the filters runs: stable epochs agree, a rounding-level tie may flip a subset."""
    N = 3000 if m < 16 else 1000
    anc, truth, r = epochs(m, N, seed=600 + m, p_nlos=0.15)
    if m == 16:  # the exact grid is singular for collinear triples (see above): jitter it
        anc = anc + np.random.default_rng(5).uniform(-0.4, 0.4, size=anc.shape) * [1, 1, 0]
        r = synth.ranges_mm(truth, anc, seed=617, p_nlos=0.15)
    ref, per = oracle_ml(oracle, r, anc, 0.01, start_for(use2d), use2d=use2d, variant=2, best_mode=best_mode,
                         n_random=16 if use2d else 48)
    got = gpu_ml(kflib, anc, r, use2d=use2d, variant=2, best_mode=best_mode, ml_start=start_for(use2d),
                 ml_exact_order=-1)
    rep = assert_parity(got, ref, per, float_keys=("pos",), int_keys=("status", "sel"),
                        min_stable=0.8, max_tie_frac=1e-3, what=f"BestGroup fast 2d={use2d} m={m}")
    print("parity report best (fast)", use2d, m, best_mode, rep)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("use2d", [0, 1])
def test_ml_exact_order_every_variant(kflib, oracle, variant, use2d):
    """ml_exact_order = 1: every epoch of every variant through the exact-order kernel -- ragged epochs,
    too few rangings, per-ranging error estimates (some of them 0), f64 / int32 / uint16 wire formats:
    all outputs bit-identical to the oracle, all epochs."""
    N, m = 4000, 8
    anc, truth, r = epochs(m, N, seed=650 + variant, p_missing=0.3, p_nlos=0.1)
    r[:, :10] = 0
    err = np.random.default_rng(2).uniform(0.005, 0.05, size=r.shape)
    err[np.random.default_rng(3).random(r.shape) < 0.01] = 0.0
    kw = dict(use2d=use2d, variant=variant, n_ignore=2)
    keys = ("sel", "status", "iters", "pos", "cov")
    for e, rr in ((0.01, r), (err, r), (0.01, r.astype(np.uint16)), (0.02, r.astype(np.float64) / 1000)):
        ref = oracle.ml_batch(rr, anc, e, start_for(use2d), **kw)
        got = gpu_ml(kflib, anc, rr, err=e, use2d=use2d, variant=variant, num_ignored_rangings=2,
                     ml_start=start_for(use2d), ml_exact_order=1)
        assert_bit_equal(got, ref, keys, f"exact order v{variant} 2d={use2d} {rr.dtype} pme={not np.isscalar(e)}")


@pytest.mark.parametrize("use2d", [0, 1])
def test_ml_ignore_n_near_ties_are_redecided_exactly(kflib, oracle, use2d):
    """Variant 1 with DUPLICATED anchors (two slots at the same place reporting the same range): the
    squared residuals tie exactly, the reference order keeps the lower index (App. B-11).  The fast
    solver flags such epochs (residual order within 1e-6 of a tie) and the exact-order kernel decides
    them: the dropped set is the oracle's on every epoch whose first solve is stable."""
    from roskfpos_b200.batch import Batch
    N, m = 6000, 8
    anc, truth, r = epochs(m, N, seed=680, p_nlos=0.3)
    anc2 = np.concatenate([anc, anc[[1, 4, 6]]])
    r2 = np.concatenate([r, r[[1, 4, 6]]])
    # "stable" is decided on perturbed copies that keep the duplicates identical: the tie stays a tie
    dup = lambda a: np.concatenate([a, a[[1, 4, 6]]])
    kw = dict(use2d=use2d, variant=1, n_ignore=3)
    ref = oracle.ml_batch(r2, anc2, 0.01, start_for(use2d), **kw)
    per = [oracle.ml_batch(dup(q), anc2, 0.01, start_for(use2d), **kw) for q in ulp_perturbations(to_metres(r))]
    with Batch(kflib.MODEL_ML, N, anchors=anc2, use2d=use2d, variant=1, num_ignored_rangings=3,
               ml_start=start_for(use2d)) as b:
        got = b.ml_solve(r2, err=0.01)
        cnt = b.counters()
    assert cnt["updates"] == N and cnt["ml_iters"] == got["iters"].astype(np.int64).sum()
    dropped = ~ref["sel"][0] & ((1 << (m + 3)) - 1)
    tied = ((dropped >> 1) ^ (dropped >> 8)) & 1 | ((dropped >> 4) ^ (dropped >> 9)) & 1 | ((dropped >> 6) ^ (dropped >> 10)) & 1
    assert tied.mean() > 0.05, "the workload no longer produces exact ties at the drop boundary"
    rep = assert_parity(got, ref, per, float_keys=("pos",), int_keys=("status", "sel"), min_stable=0.94,
                        max_tie_frac=0.0, what=f"IgnoreN exact ties 2d={use2d}")
    print("parity report ignoreN ties", use2d, rep, float(tied.mean()))


def test_ml_zero_noise_recovers_truth(kflib):
    N, m = 4096, 8
    anc = synth.anchors_for(m)
    rng = np.random.default_rng(3)
    truth = np.stack([rng.uniform(1, 9, N), rng.uniform(1, 9, N), rng.uniform(0.8, 1.6, N)])
    d = np.sqrt(((truth[None] - anc[:, :, None]) ** 2).sum(axis=1))
    got = gpu_ml(kflib, anc, d, err=0.01)
    assert np.abs(got["pos"] - truth).max() < 1e-5


@pytest.mark.parametrize("variant", [0, 1])
def test_ml_straggler_queue(kflib, oracle, variant):
    """16 anchors, 3-D from (1,1,4): about one epoch in a thousand needs more than the main
    kernel's 32 Newton iterations (most of those run to the reference's 10000 cap) and is parked
    in the straggler queue and resumed by the second launch.  Parked or not, every epoch whose
    oracle result is stable must agree, iteration counts included; every epoch is counted once."""
    from roskfpos_b200.batch import Batch
    N, m = 150000, 16
    anc, truth, r = epochs(m, N, seed=900 + variant)
    ref, per = oracle_ml(oracle, r, anc, 0.01, start_for(0), variant=variant, n_ignore=2)
    with Batch(kflib.MODEL_ML, N, anchors=anc, variant=variant, num_ignored_rangings=2) as b:
        got = b.ml_solve(r, err=0.01)
        cnt = b.counters()
    slow = ref["iters"] > 32 * (2 if variant else 1)
    assert slow.sum() >= 20, "the workload no longer exercises the queue"
    assert cnt["updates"] == N and cnt["ml_iters"] == got["iters"].astype(np.int64).sum()
    keys = dict(float_keys=("pos",), int_keys=("status", "iters", "sel"))
    rep = assert_parity(got, ref, per, min_stable=0.99, max_tie_frac=1e-3, what=f"stragglers v{variant}", **keys)
    # the parked epochs on their own: those the oracle calls stable must all agree
    sub = lambda d: {k: np.asarray(v)[..., slow] for k, v in d.items()}
    rep_slow = assert_parity(sub(got), sub(ref), [sub(p) for p in per], min_stable=0.0, max_tie_frac=0.0,
                             what=f"parked epochs v{variant}", **keys)
    print("parity report stragglers", variant, rep, rep_slow, int(slow.sum()))


def test_ml_z_gate(kflib):
    """minZ / maxZ of config_pos.xml: estimates whose z falls outside are flagged, not altered."""
    N, m = 5000, 8
    anc, truth, r = epochs(m, N, seed=950)
    plain = gpu_ml(kflib, anc, r)
    gated = gpu_ml(kflib, anc, r, min_z=0.9, max_z=1.1)
    assert np.array_equal(plain["pos"], gated["pos"]) and not (plain["status"] & kflib.ST_Z_GATE).any()
    out = (gated["pos"][2] < 0.9) | (gated["pos"][2] > 1.1)
    assert 0.05 < out.mean() < 0.95
    assert np.array_equal((gated["status"] & kflib.ST_Z_GATE) != 0, out)
    assert np.array_equal(gated["status"] & ~kflib.ST_Z_GATE, plain["status"])


def test_ml_best_group_3d_parked_solves_overflow(kflib, oracle, monkeypatch):
    """3-D BestGroup parks long subset solves in task records; when the records run out (forced here with a
    capacity of 40 tasks) the warp finishes its parked solves in place: same bits either way."""
    N, m = 160, 16
    anc, truth, r = epochs(m, N, seed=616, p_nlos=0.15)
    ref = oracle.ml_batch(r, anc, 0.01, start_for(0), use2d=0, variant=2, best_mode=0)
    keys = ("sel", "status", "iters", "pos", "cov")
    got = gpu_ml(kflib, anc, r, use2d=0, variant=2, best_mode=0, ml_start=start_for(0))
    assert_bit_equal(got, ref, keys, "BestGroup 3-D parked")
    assert ref["iters"].max() > 10000, "the workload no longer has subset solves that run to the cap"
    monkeypatch.setenv("KFPOS_XW_TASK_CAP", "40")
    got = gpu_ml(kflib, anc, r, use2d=0, variant=2, best_mode=0, ml_start=start_for(0))
    assert_bit_equal(got, ref, keys, "BestGroup 3-D parked, task records exhausted")


def test_ml_best_group_3d_chunks_ragged_per_ranging_errors(kflib, oracle, monkeypatch):
    """The 3-D scan goes through the batch in chunks (131072 epochs; forced to 48 here so that a 200-epoch batch takes
    five passes with a short last one), with missing rangings, per-ranging error estimates and all three wire formats."""
    N, m = 200, 12
    anc, truth, r = epochs(m, N, seed=640, p_missing=0.25, p_nlos=0.1)
    r[:, :6] = 0
    err = np.random.default_rng(4).uniform(0.005, 0.05, size=r.shape)
    keys = ("sel", "status", "iters", "pos", "cov")
    monkeypatch.setenv("KFPOS_XW_CHUNK", "48")
    for e, rr in ((0.01, r), (err, r), (0.01, r.astype(np.uint16)), (0.02, r.astype(np.float64) / 1000)):
        for best_mode in (0, 1):
            ref = oracle.ml_batch(rr, anc, e, start_for(0), use2d=0, variant=2, best_mode=best_mode)
            got = gpu_ml(kflib, anc, rr, err=e, use2d=0, variant=2, best_mode=best_mode, ml_start=start_for(0))
            assert_bit_equal(got, ref, keys, f"BestGroup 3-D chunks {rr.dtype} pme={not np.isscalar(e)} mode={best_mode}")
