"""GPU: the header-only C++ mirror of the reference classes (include/kfpos/*.hpp), driven like
PosGenerator drives the reference, against the oracle objects."""
import os
import subprocess

import numpy as np
import pytest

from roskfpos_b200 import lib as L, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "class_mirror")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "class_mirror.cpp"), L.SO_PATH,
                           "-Wl,-rpath," + os.path.dirname(L.SO_PATH), "-o", out])
    return out


def run(exe, script):
    p = subprocess.run([exe], input="\n".join(script) + "\n", capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    return [list(map(float, l.split()[1:])) for l in p.stdout.splitlines() if l.startswith("pose")]


def fmt(v):
    return " ".join(repr(float(x)) for x in v)


def test_cpp_kalman_filter_toa(exe, oracle):
    anc = synth.anchors_for(8)
    truth = synth.truth_lissajous(1, 20, 0.1, seed=5)
    r = synth.ranges_mm(truth[1:], anc, seed=6, p_missing=0.2)[:, :, 0] / 1000.0
    script = ["anchors 8 " + fmt(anc.ravel()), "t6 0.5 0 0.0 " + fmt(truth[0][:, 0]), "pose"]
    o = oracle.T6(0.5, False, 0.0, truth[0][:, 0])
    for t in range(20):
        script += [f"toa 0.1 0.01 {fmt(r[t])}"]
        o.new_toa(0.1, r[t], anc, 0.01)
    script += ["pose"]
    poses = run(exe, script)
    assert poses[0][0] == 0  # getPose is false before the first measurement (TOA.cpp:442-447)
    assert poses[1][0] == 1
    # the pose poll predicts to "now"; position is unchanged by the prediction (v = 0)
    assert np.abs(np.array(poses[1][1:4]) - o.pos).max() < 1e-9


def test_cpp_kalman_filter_multi_sensor(exe, oracle):
    anc = synth.anchors_for(8)
    w = synth.k8_workload(1, 3, anc, seed=9, full=True)
    o = oracle.K8(0.5, float(w["x0"][6, 0]), 0.5, w["x0"][:2, 0], **synth.K8_ORACLE_CFG)
    script = ["anchors 8 " + fmt(anc.ravel()), f"k8 0.5 {float(w['x0'][6, 0])!r} 0.5 {fmt(w['x0'][:2, 0])}"]
    s = w["sensors"][:, 0]
    rr = w["ranges"].reshape(-1, 1)[:, 0] / 1000.0
    cav = np.zeros(9); cac = np.zeros(9)
    for (kind, dt, off, aux) in w["events"]:
        if kind == synth.EV_TOA:
            script.append(f"toa {dt!r} 0.01 {fmt(rr[off:off + 8])}")
            o.new_toa(dt, rr[off:off + 8], anc, 0.01)
        elif kind == synth.EV_IMU:
            script.append(f"imu {dt!r} 0 0 {fmt(s[off:off + 3])} 0 {fmt(cav)} {fmt(cac)}")
            o.new_imu(dt, [0, 0, s[off]], cav, [s[off + 1], s[off + 2], 0], cac)
        elif kind == synth.EV_PX4:
            script.append(f"px4 {dt!r} {fmt(s[off:off + 4])} {int(s[off + 4])}")
            o.new_px4(dt, s[off], s[off + 1], s[off + 2], s[off + 3], int(s[off + 4]))
        else:
            script.append(f"compass {dt!r} {fmt(s[off:off + 1])}")
            o.new_compass(dt, s[off])
    script.append("pose")
    p = run(exe, script)[-1]
    assert p[0] == 1 and p[3] == 1.049
    assert abs(p[1] - o.x[0]) < 1e-3 and abs(p[2] - o.x[1]) < 1e-3  # predicted to "now" (a few ms)
    assert abs(p[6] - o.x[2]) < 1e-9 and abs(p[7] - o.x[3]) < 1e-9 and abs(p[8] - o.x[7]) < 1e-9
    assert p[10] == 0


def test_cpp_ml_and_t9(exe, oracle):
    anc = synth.anchors_for(8)
    tp = np.array([4.2, 6.1, 1.0])
    r = np.sqrt(((anc - tp) ** 2).sum(1)) + np.random.default_rng(2).normal(0, 0.05, 8)
    script = ["anchors 8 " + fmt(anc.ravel()), "ml 0 1 2 1 1 4", f"toa 0.1 0.01 {fmt(r)}", "pose",
              "t9 0.5 0.5 " + fmt(tp), f"toa 0.1 0.01 {fmt(r)}", f"toa 0.1 0.01 {fmt(r)}", "pose"]
    poses = run(exe, script)
    ref = oracle.ml_epoch(r, anc, 0.01, [1, 1, 4.0], variant=1, n_ignore=2)
    assert np.abs(np.array(poses[0][1:4]) - ref["pos"]).max() < 1e-9
    assert abs(poses[0][9] - ref["cov"][0, 0]) <= 1e-9 * abs(ref["cov"][0, 0])
    t9 = oracle.T9(0.5, 0.5, tp)
    t9.new_toa(0.1, r, anc, 0.01)
    t9.new_toa(0.1, r, anc, 0.01)
    assert np.abs(np.array(poses[1][6:8]) - t9.x[3:5]).max() < 1e-9  # velocities persisted


def test_plain_c_caller_of_the_stats_collective(kflib, tmp_path):
    """tests/cpp/stats_c_abi.c is compiled as C against include/kfpos_b200.h and calls the boundary the way a
    maintainer's code would, ending in kfpos_stats_allreduce; same numbers as the Python host side."""
    from roskfpos_b200.batch import Batch
    out = str(tmp_path / "stats_c_abi")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "stats_c_abi.c"), L.SO_PATH,
                           "-Wl,-rpath," + os.path.dirname(L.SO_PATH), "-o", out])
    N, T, M = 300, 5, 8
    anc = synth.anchors_for(M)
    truth = synth.truth_lissajous(N, T, 0.1, seed=31)
    r = synth.ranges_mm(truth[1:], anc, seed=32).astype(np.float64) / 1000
    x0 = np.zeros((6, N)); x0[:3] = truth[0]
    text = f"{N} {T} {M}\n" + "\n".join(" ".join(repr(float(v)) for v in a.ravel()) for a in (anc, x0, truth[-1], r)) + "\n"
    p = subprocess.run([out], input=text, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    got = np.array([float(v) for v in p.stdout.split()[1:]])
    with Batch(kflib.MODEL_T6, N, anchors=anc, accel_noise=0.5) as b:
        b.set_state(x0)
        b.replay_toa(0.1, r, err=0.01)
        ref = b.error_stats(truth[-1])
    assert np.array_equal(got[:4], ref[:4])
    assert got[4] == np.sqrt(ref[0] / ref[2]) and got[2] == N
