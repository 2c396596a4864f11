"""GPU parity: KalmanFilter (K8: UWB + IMU + magnetometer/compass + PX4Flow) through the C ABI."""
import numpy as np
import pytest

from roskfpos_b200 import synth
from tests.util import REL_TOL, assert_parity, rel_err_cov, rel_err_state

pytestmark = pytest.mark.gpu


def gpu_k8(kflib, w, anc, want_traj=False, **kw):
    from roskfpos_b200.batch import Batch
    N = w["x0"].shape[-1]
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5, **kw) as b:
        b.set_state(w["x0"])
        traj = b.replay_events(w["events"], ranges=w["ranges"], sensors=w["sensors"], err=0.01, want_traj=want_traj)
        x, P, st = b.get_state()
        cnt = b.counters()
        stats = b.error_stats(w["truth_end"])
    return dict(x=x, P=P, status=st & ~32, traj=traj, counters=cnt, stats=stats)


def oracle_k8(oracle, w, anc, want_traj=False):
    cfg = oracle.k8_cfg(0.5, 0.5, **synth.K8_ORACLE_CFG)
    out = oracle.k8_replay(w["x0"], None, w["events"], w["ranges"], w["sensors"], anc, 0.01, cfg, want_traj=want_traj)
    out["status"] = out["status"] & ~32
    return out


@pytest.mark.parametrize("full,N,n_macro", [(False, 4096, 12), (True, 3000, 10), (True, 20000, 3)])
def test_k8_event_stream_parity(kflib, oracle, full, N, n_macro):
    """BASELINE configs 3 (IMU + compass + TOA) and 5 (+ PX4Flow): every filter within 1e-9,
    iteration counters identical."""
    anc = synth.anchors_for(8)
    w = synth.k8_workload(N, n_macro, anc, seed=40 + N, full=full)
    ref = oracle_k8(oracle, w, anc, want_traj=True)
    got = gpu_k8(kflib, w, anc, want_traj=True)
    assert rel_err_state(got["x"], ref["x"]) < REL_TOL
    assert rel_err_cov(got["P"], ref["P"]) < REL_TOL
    assert rel_err_state(got["traj"], ref["traj"]) < REL_TOL
    c = got["counters"]
    assert [c["updates"], c["ml_iters"], c["cost_evals"], c["gain_evals"]] == [ref["counters"][4], *ref["counters"][:3]]
    assert np.array_equal(got["status"], ref["status"])
    e2 = ((got["x"][:2] - w["truth_end"][:2]) ** 2).sum(axis=0)
    assert abs(got["stats"][0] - e2.sum()) <= 1e-9 * e2.sum() and got["stats"][2] == N


def test_k8_step_api_equals_event_replay(kflib, oracle):
    """The five reference callbacks as single ABI calls == the fused event schedule; includes a
    non-diagonal IMU covariance, raw magnetometer events and PX4 frames of quality 0 (skipped,
    their dt carried over: KF.cpp:111-113)."""
    from roskfpos_b200.batch import Batch
    N = 1024
    anc = synth.anchors_for(8)
    w = synth.k8_workload(N, 4, anc, seed=77, full=True)
    rng = np.random.default_rng(1)
    sens = w["sensors"].copy()
    events = []
    extra_rows = []
    for (kind, dt, off, aux) in w["events"]:
        if kind == synth.EV_IMU:
            aux = [0.004, 1e-4, 1e-4, 0.005, 0.07]
        if kind == synth.EV_PX4:
            sens[off + 4, rng.random(N) < 0.3] = 0.0  # quality 0 for 30 % of the filters
        if kind == synth.EV_COMPASS and rng.random() < 0.5:  # replace by a raw magnetometer sample
            ang = sens[off]
            extra_rows += [np.cos(ang), np.sin(ang)]
            kind, off = synth.EV_MAG, len(sens) + len(extra_rows) - 2
        events.append((kind, dt, off, aux))
    sens = np.vstack([sens] + [r[None] for r in extra_rows]) if extra_rows else sens
    w2 = dict(w, events=events, sensors=sens)
    xml = list(synth.K8_XML)
    xml[2] = '<config><imu useFixedCovarianceAcceleration="0" useFixedCovarianceAngularVelocityZ="0"/></config>'
    cfg = oracle.k8_cfg(0.5, 0.5, **dict(synth.K8_ORACLE_CFG, imu_fixed_cov_acc=0, imu_fixed_cov_gyro=0))
    ref = oracle.k8_replay(w["x0"], None, events, w["ranges"], sens, anc, 0.01, cfg)
    outs = []
    for mode in ("events", "steps"):
        with Batch(kflib.MODEL_K8, N, anchors=anc, xml=xml, accel_noise=0.5, jolt=0.5) as b:
            b.set_state(w["x0"])
            if mode == "events":
                b.replay_events(events, ranges=w["ranges"], sensors=sens, err=0.01)
            else:
                M = len(anc)
                for (kind, dt, off, aux) in events:
                    if kind == synth.EV_TOA:
                        b.step_toa(dt, w["ranges"].reshape(-1, N)[off:off + M], err=0.01)
                    elif kind == synth.EV_IMU:
                        av = np.zeros((3, N)); av[2] = sens[off]
                        la = np.zeros((3, N)); la[0] = sens[off + 1]; la[1] = sens[off + 2]
                        cav = np.zeros(9); cav[8] = aux[4]
                        cac = np.zeros(9); cac[0], cac[1], cac[3], cac[4] = aux[0], aux[1], aux[2], aux[3]
                        b.step_imu(dt, av, la, cov_ang_vel=cav, cov_acc=cac)
                    elif kind == synth.EV_PX4:
                        b.step_px4(dt, sens[off], sens[off + 1], sens[off + 2], sens[off + 3], sens[off + 4].astype(np.int32))
                    elif kind == synth.EV_MAG:
                        b.step_mag(dt, np.vstack([sens[off:off + 2], np.zeros((1, N))]))
                    else:
                        b.step_compass(dt, sens[off])
            outs.append(b.get_state())
    for x, P, st in outs:
        assert rel_err_state(x, ref["x"]) < REL_TOL
        assert rel_err_cov(P, ref["P"]) < REL_TOL
    # the step API carries skipped-PX4 time differently only if the carry were lost: must agree
    assert rel_err_state(outs[0][0], outs[1][0]) < 1e-12


def test_k8_toa_only_replay_and_get_pose(kflib, oracle):
    """config 1b: K8 ranging-only ("fixed height 2D"), through replay_toa; getPose is non-mutating."""
    from roskfpos_b200.batch import Batch
    N, T = 2048, 30
    anc = synth.anchors_for(4)
    truth = synth.truth_lissajous(N, T, 0.1, seed=91, z=1.0)
    r = synth.ranges_mm(truth[1:], anc, seed=92)
    x0 = np.zeros((8, N)); x0[:2] = truth[0][:2]; x0[6] = 0.3
    cfg = oracle.k8_cfg(0.5, 0.5, tag_z=1.0)
    ev = [(0, 0.1, t * 4, None) for t in range(T)]
    ref = oracle.k8_replay(x0, None, ev, r, None, anc, 0.01, cfg)
    with Batch(kflib.MODEL_K8, N, anchors=anc, accel_noise=0.5, jolt=0.5, fixed_height=1.0) as b:
        b.set_state(x0)
        b.replay_toa(0.1, r, err=0.01)
        x, P, st = b.get_state()
        xp, Pp = b.get_pose(0.04)
        x2, P2, _ = b.get_state()
    assert rel_err_state(x, ref["x"]) < REL_TOL and rel_err_cov(P, ref["P"]) < REL_TOL
    assert np.array_equal(x, x2) and np.array_equal(P, P2)
    for f in (0, 100, 2047):
        o = oracle.K8(0.5, 0.3, 0.5, x[:2, f], tag_z=1.0)
        o.f.vel[0], o.f.vel[1], o.f.angle, o.f.omega = x[2, f], x[3, f], x[6, f], x[7, f]
        for k in range(64):
            o.f.P[k] = P[k, f]
        xr, Pr = o.get_pose(0.04)
        assert np.abs(xp[:, f] - xr).max() < 1e-12
        assert np.abs(Pp[:, f].reshape(8, 8) - Pr).max() <= 1e-12 * np.abs(Pr).max()


@pytest.mark.parametrize("variant,n_ignore", [(1, 2), (2, 0)])
def test_k8_ekf_side_nlos_variants(kflib, oracle, variant, n_ignore):
    """EKF-side variants for the planar filter: variant 1 drops the N worst rangings, variant 2 keeps the
    best THREE anchors (the 2-D best group), selected by the 2-D ML estimator from the predicted position;
    compass events in between so that the latched sensor rows take part in the update."""
    from roskfpos_b200.batch import Batch
    from tests.util import to_metres, ulp_perturbations
    N, T, m = 2500, 4, 8
    anc = synth.anchors_for(m)
    truth = synth.truth_lissajous(N, T, 0.1, seed=401, z=1.049)
    r = synth.ranges_mm(truth[1:], anc, seed=402, p_nlos=0.15)
    rng = np.random.default_rng(403)
    comp = rng.uniform(-3, 3, size=(T, N))
    events = []
    for t in range(T):
        events.append((synth.EV_COMPASS, 0.03, t, None))
        events.append((synth.EV_TOA, 0.07, t * m, None))
    x0 = np.zeros((8, N)); x0[:2] = truth[0][:2]; x0[6] = comp[0]
    cfg = oracle.k8_cfg(0.5, 0.5, variant=variant, n_ignore=n_ignore, **synth.K8_ORACLE_CFG)
    run = lambda rr: oracle.k8_replay(x0, None, events, rr, comp, anc, 0.01, cfg)
    ref = run(r)
    per = [run(p) for p in ulp_perturbations(to_metres(r), n_random=12)]
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5, variant=variant,
               num_ignored_rangings=n_ignore) as b:
        b.set_state(x0)
        b.replay_events(events, ranges=r, sensors=comp, err=0.01)
        x, P, st = b.get_state()
        cnt = b.counters()
    plain = oracle.k8_replay(x0, None, events, r, comp, anc, 0.01, oracle.k8_cfg(0.5, 0.5, **synth.K8_ORACLE_CFG))
    assert np.abs(plain["x"] - ref["x"]).max() > 1e-3  # the selection does change the estimate
    got = dict(x=x, P=P, status=st & ~32)
    for d in [ref] + per:
        d["status"] = d["status"] & ~32
    rep = assert_parity(got, ref, per, float_keys=("x",), cov_keys=("P",), int_keys=("status",),
                        min_stable=0.9 if variant == 1 else 0.6, max_tie_frac=1e-3,
                        what=f"K8 variant {variant}")
    print("parity report K8 variant", variant, rep, cnt)
    assert cnt["updates"] == N * len(events)


def test_k8_checkpoint_with_latches(kflib):
    """get_state + get_latches / set_state + set_latches is a checkpoint of a running multi-sensor batch: the
    restored batch continues bit-identically (latched PX4 / IMU / mag samples, their flags, the IMU covariances
    and the time carried over from PX4 frames of quality 0 are all part of the filter)."""
    from roskfpos_b200.batch import Batch
    N = 1500
    anc = synth.anchors_for(8)
    w = synth.k8_workload(N, 4, anc, seed=123, full=True)
    sens = w["sensors"].copy()
    rng = np.random.default_rng(3)
    for (kind, dt, off, aux) in w["events"]:
        if kind == synth.EV_PX4:
            sens[off + 4, rng.random(N) < 0.3] = 0.0  # quality 0: the frame is skipped, its dt carried
    ev = w["events"]
    # cut right after a PX4 frame: the filters that skipped it carry its dt into the next event
    cut = next(i for i in range(len(ev) // 2, len(ev)) if ev[i][0] == synth.EV_PX4) + 1
    kw = dict(anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5)
    with Batch(kflib.MODEL_K8, N, **kw) as b:
        b.set_state(w["x0"])
        b.replay_events(ev, ranges=w["ranges"], sensors=sens, err=0.01)
        x_ref, P_ref, _ = b.get_state()
    with Batch(kflib.MODEL_K8, N, **kw) as b:
        b.set_state(w["x0"])
        b.replay_events(ev[:cut], ranges=w["ranges"], sensors=sens, err=0.01)
        x1, P1, _ = b.get_state()
        latches = b.get_latches()
    assert latches[1].max() == 7 and np.abs(latches[0][8]).max() > 0  # every sensor latched, some dt carried
    with Batch(kflib.MODEL_K8, N, **kw) as b:
        b.set_state(x1, P1)
        b.set_latches(*latches)
        b.replay_events(ev[cut:], ranges=w["ranges"], sensors=sens, err=0.01)
        x2, P2, _ = b.get_state()
    assert np.array_equal(x2, x_ref)
    # the covariance crosses the ABI as a full matrix and is packed again: symmetric, so nothing is lost
    assert np.array_equal(P2, P_ref)


def test_k8_imu_event_break_patterns(kflib, oracle):
    """The IMU event runs as straight-line code that assumes the break tests come out as (continue, continue, stop)
    and repeats the event in the general form otherwise.  Filters whose IMU sample equals the predicted output
    exactly (zero cost: the relative-change quotient is 0 / 0, the loop runs to its 20-iteration cap) and filters
    with gross outliers take that other path: state, covariance, status and the iteration counters must be the
    oracle's either way."""
    N = 4096
    anc = synth.anchors_for(8)
    w = synth.k8_workload(N, 3, anc, seed=91, full=False)
    sens = w["sensors"].copy()
    first = True
    for (kind, dt, off, aux) in w["events"]:
        if kind != synth.EV_IMU:
            continue
        if first:  # first IMU event: omega is still x0's, the predicted accelerations are 0
            sens[off, 0::4] = w["x0"][7, 0::4]
            sens[off + 1, 0::4] = 0.0
            sens[off + 2, 0::4] = 0.0
            first = False
        else:  # later ones: gross outliers for another quarter of the filters
            sens[off, 1::4] += 50.0
            sens[off + 1, 1::4] -= 80.0
    w2 = dict(w, sensors=sens)
    ref = oracle_k8(oracle, w2, anc)
    got = gpu_k8(kflib, w2, anc)
    assert ref["counters"][1] != 3 * ref["counters"][4] or ref["counters"][2] != 2 * ref["counters"][4], \
        "every event still runs 3 cost evaluations and 2 gain steps: the workload does not leave the fast path"
    assert rel_err_state(got["x"], ref["x"]) < REL_TOL
    assert rel_err_cov(got["P"], ref["P"]) < REL_TOL
    c = got["counters"]
    assert [c["updates"], c["ml_iters"], c["cost_evals"], c["gain_evals"]] == [ref["counters"][4], *ref["counters"][:3]]
    assert np.array_equal(got["status"], ref["status"])
