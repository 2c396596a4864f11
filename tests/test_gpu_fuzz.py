"""GPU parity, randomised: T6 replays over randomly drawn configurations (anchor count incl. the
compile-time-specialised 4 / 8 / 16, wire format, missing rangings, per-ranging errors, variable dt,
leave-one-out, EKF-side variants, restored covariance) against the oracle, with the stability rule of
tests/util.py.  Fixed seeds: the draw is the same on every run."""
import os

import numpy as np
import pytest

from roskfpos_b200 import synth
from tests.util import assert_parity, to_metres, ulp_perturbations

pytestmark = pytest.mark.gpu


def draw(seed):
    rng = np.random.default_rng(seed)
    m = int(rng.choice([4, 5, 7, 8, 8, 9, 12, 16, 16, 24, 32]))
    c = dict(m=m, N=int(rng.integers(1, 700)), T=int(rng.integers(1, 14)),
             fmt=[np.float64, np.int32, np.uint16][int(rng.integers(0, 3))],
             p_missing=float(rng.choice([0.0, 0.0, 0.1, 0.4])), pme=bool(rng.integers(0, 2)),
             var_dt=bool(rng.integers(0, 2)), mode=str(rng.choice(["plain", "plain", "loo", "v1", "v2"])),
             with_P0=bool(rng.integers(0, 2)))
    if c["mode"] == "v2" and m > 9:
        c["mode"] = "v1"  # C(m, 4) oracle solves per update: keep the test fast
    if c["mode"] == "v2":
        c["T"] = min(c["T"], 2)  # rounding-level ties of the 4-anchor groups compound along a trajectory
    return c, rng


@pytest.mark.parametrize("seed", range(int(os.environ.get("KF_FUZZ_SEEDS", "24"))))  # more: KF_FUZZ_SEEDS=400
def test_t6_random_configuration(kflib, oracle, seed):
    from roskfpos_b200.batch import Batch
    c, rng = draw(1000 + seed)
    m, N, T = c["m"], c["N"], c["T"]
    anc = synth.anchors_for(m)
    truth = synth.truth_lissajous(N, T, 0.1, seed=seed)
    mm = synth.ranges_mm(truth[1:], anc, seed=seed + 50, p_missing=c["p_missing"],
                         p_nlos=0.15 if c["mode"] != "plain" else 0.0)
    r = (mm.astype(np.float64) / 1000) if c["fmt"] is np.float64 else mm.astype(c["fmt"])
    dt = rng.uniform(0.03, 0.25, size=T) if c["var_dt"] else 0.1
    err = rng.uniform(0.005, 0.05, size=r.shape) if c["pme"] else 0.01
    P0 = None
    if c["with_P0"]:
        A = rng.normal(size=(N, 6, 6)) * 0.1
        P0 = np.ascontiguousarray(np.einsum("nij,nkj->ikn", A, A).reshape(36, N))
    kw_o, kw_g = {}, dict(accel_noise=0.5)
    if c["mode"] == "loo":
        kw_o.update(ignore_worst=True, thr=0.3); kw_g.update(ignore_worst_anchor=1, ignore_cost_threshold=0.3)
    elif c["mode"] in ("v1", "v2"):
        v = 1 if c["mode"] == "v1" else 2
        kw_o.update(variant=v, n_ignore=2); kw_g.update(variant=v, num_ignored_rangings=2)
    run = lambda rr: oracle.t6_replay(truth[0], P0, rr, anc, dt, err, **kw_o)
    ref = run(r)
    per = [run(p) for p in ulp_perturbations(to_metres(r), n_random=32 if c["mode"] in ("v1", "v2") else 8)]
    with Batch(kflib.MODEL_T6, N, anchors=anc, **kw_g) as b:
        x0 = np.zeros((6, N)); x0[:3] = truth[0]
        b.set_state(x0, P0)
        _, sel = b.replay_toa(dt, r, err=err, want_sel=True)
        x, P, st = b.get_state()
    got = dict(x=x[:3], P=P, status=st & ~32, sel=sel)
    for d in [ref] + per:
        d["status"] = d["status"] & ~32
    # more discrete decisions per unit.  Measured on 200 000 filters x 20 epochs against the oracle: 0.08 % of
    # the plain filters and 2.2 % of the variant-1 filters (15 % NLOS rangings) leave the 1e-9 band at some
    # step -- rounding-level ties of an iteration count or of the residual order, whichever way the
    # sums are associated -- so the small batches of this test need room for one or two of them
    few = m <= 5 or c["p_missing"] >= 0.4 or c["mode"] in ("v1", "v2", "loo")
    rep = assert_parity(got, ref, per, float_keys=("x",), cov_keys=("P",), int_keys=("status", "sel"),
                        min_stable=0.4 if few else 0.85, max_tie_frac=1e-3, what=str(c))
    print("fuzz", seed, c, rep)


@pytest.mark.parametrize("seed", range(int(os.environ.get("KF_FUZZ_SEEDS", "16"))))
def test_k8_random_event_schedule(kflib, oracle, seed):
    """K8: random schedules of the five event kinds (random order and dt, PX4 frames of quality 0, raw
    magnetometer samples, missing rangings, per-ranging errors, 4 / 6 / 8 anchors) against the oracle."""
    from roskfpos_b200.batch import Batch
    rng = np.random.default_rng(5000 + seed)
    m = int(rng.choice([4, 6, 8, 8]))
    N, n_ev = int(rng.integers(1, 400)), int(rng.integers(3, 40))
    pme = bool(rng.integers(0, 2))
    anc = synth.anchors_for(m)
    th = rng.uniform(-3, 3, N)
    pos = np.stack([rng.uniform(2, 8, N), rng.uniform(2, 8, N)])
    x0 = np.zeros((8, N)); x0[:2] = pos; x0[2:4] = rng.normal(0, 0.3, (2, N)); x0[6] = th; x0[7] = rng.normal(0, 0.1, N)
    d = np.sqrt((pos[0][None] - anc[:, 0:1]) ** 2 + (pos[1][None] - anc[:, 1:2]) ** 2 + (1.049 - anc[:, 2:3]) ** 2)
    events, rows, rng_rows = [], [], []
    for _ in range(n_ev):
        kind = int(rng.choice([synth.EV_TOA, synth.EV_PX4, synth.EV_IMU, synth.EV_IMU, synth.EV_MAG, synth.EV_COMPASS]))
        dt = float(rng.uniform(0.003, 0.12))
        if kind == synth.EV_TOA:
            r = np.floor((d + 0.1 * rng.normal(size=d.shape)) * 1000)
            r[rng.random(r.shape) < 0.15] = 0
            events.append((kind, dt, len(rng_rows) * m, None)); rng_rows.append(r)
        elif kind == synth.EV_PX4:
            q = np.where(rng.random(N) < 0.25, 0.0, 150.0)
            off = len(rows)
            rows += [rng.normal(0, 2e-3, N), rng.normal(0, 2e-3, N), rng.normal(0, 1e-3, N), np.full(N, 33333.0), q]
            events.append((kind, dt, off, None))
        elif kind == synth.EV_IMU:
            off = len(rows)
            rows += [rng.normal(0.05, 0.3, N), rng.normal(0, 0.3, N), rng.normal(0, 0.3, N)]
            events.append((kind, dt, off, [0.004, 1e-4, 1e-4, 0.005, 0.07]))
        elif kind == synth.EV_MAG:
            off = len(rows)
            a = th + 0.05 * rng.normal(size=N)
            rows += [np.cos(a), np.sin(a)]
            events.append((kind, dt, off, None))
        else:
            off = len(rows)
            rows.append(np.arctan2(np.sin(th + 0.02 * rng.normal(size=N)), np.cos(th)))
            events.append((kind, dt, off, None))
    ranges = np.ascontiguousarray(np.array(rng_rows).reshape(-1, m, N).astype(np.int32)) if rng_rows else \
        np.zeros((1, m, N), dtype=np.int32)
    sensors = np.ascontiguousarray(np.array(rows)) if rows else np.zeros((1, N))
    err = rng.uniform(0.005, 0.05, size=ranges.shape) if pme else 0.01
    xml = list(synth.K8_XML)
    ocfg = dict(synth.K8_ORACLE_CFG)
    if rng.random() < 0.5:  # covariances from the messages instead of the fixed XML values
        xml[2] = '<config><imu useFixedCovarianceAcceleration="0" useFixedCovarianceAngularVelocityZ="0"/></config>'
        ocfg.update(imu_fixed_cov_acc=0, imu_fixed_cov_gyro=0)
    cfg = oracle.k8_cfg(0.5, 0.5, **ocfg)
    ref = oracle.k8_replay(x0, None, events, ranges, sensors, anc, err, cfg)
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=xml, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(x0)
        b.replay_events(events, ranges=ranges, sensors=sensors, err=err)
        x, P, st = b.get_state()
    ok = np.isfinite(ref["x"]).all(axis=0)
    assert ok.mean() > 0.9
    ex = np.abs(x - ref["x"])[:, ok].max() if ok.any() else 0.0
    with np.errstate(invalid="ignore", divide="ignore"):
        eP = (np.abs(P - ref["P"]).max(axis=0) / np.abs(ref["P"]).max(axis=0))[ok].max() if ok.any() else 0.0
    assert ex < 1e-9 and eP < 1e-9, (m, N, n_ev, pme, ex, eP)
    assert np.array_equal(st[ok] & ~(32 | 64), ref["status"][ok] & ~(32 | 64))
