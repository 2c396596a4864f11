"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(roskfpos_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class KoInfo(C.Structure):
    _fields_ = [("status", C.c_int), ("ml_iters", C.c_int), ("cost_evals", C.c_int),
                ("gain_evals", C.c_int), ("ignored", C.c_int), ("cost", C.c_double)]


class KoT6(C.Structure):
    _fields_ = [("accel_noise", C.c_double), ("ignore_worst", C.c_int),
                ("ignore_cost_threshold", C.c_double), ("pos", C.c_double * 3),
                ("vel", C.c_double * 3), ("P", C.c_double * 36)]


class KoK8(C.Structure):
    _fields_ = [("accel_noise", C.c_double), ("jolt", C.c_double), ("tag_z", C.c_double),
                ("use_fixed_height", C.c_int),
                ("px4_height", C.c_double), ("px4_arm1", C.c_double), ("px4_arm2", C.c_double),
                ("px4_cov_vel", C.c_double), ("px4_cov_gyro", C.c_double),
                ("imu_fixed_cov_acc", C.c_int), ("imu_cov_acc", C.c_double),
                ("imu_fixed_cov_gyro", C.c_int), ("imu_cov_gyro", C.c_double),
                ("mag_offset", C.c_double), ("mag_cov", C.c_double),
                ("pos", C.c_double * 2), ("vel", C.c_double * 2), ("acc", C.c_double * 2),
                ("angle", C.c_double), ("omega", C.c_double), ("P", C.c_double * 64),
                ("has_mag", C.c_int), ("has_px4", C.c_int), ("has_imu", C.c_int),
                ("px4_itime", C.c_double), ("px4_vx", C.c_double), ("px4_vy", C.c_double),
                ("px4_gz", C.c_double), ("px4_cv", C.c_double), ("px4_cg", C.c_double),
                ("imu_wz", C.c_double), ("imu_cwz", C.c_double), ("imu_ax", C.c_double),
                ("imu_ay", C.c_double), ("imu_cxy", C.c_double * 4),
                ("mag_angle", C.c_double), ("mag_c", C.c_double),
                ("variant", C.c_int), ("n_ignore", C.c_int), ("best_mode", C.c_int), ("ml_init", C.c_int)]


class KoT9(C.Structure):
    _fields_ = [("accel_noise", C.c_double), ("jolt", C.c_double), ("pos", C.c_double * 3),
                ("vel", C.c_double * 3), ("acc", C.c_double * 3), ("P", C.c_double * 81),
                ("has_imu", C.c_int), ("imu_a", C.c_double * 3), ("imu_cov", C.c_double * 9),
                ("variant", C.c_int), ("n_ignore", C.c_int), ("best_mode", C.c_int), ("ml_init", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ko_sse.restype = C.c_double
        _LIB.ko_max_threads.restype = C.c_int
    return _LIB


def _p(a, ct=C.c_double):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ct))


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


FMT = {np.dtype(np.float64): 0, np.dtype(np.int32): 1, np.dtype(np.uint16): 2}


def max_threads() -> int:
    return lib().ko_max_threads()


# --------------------------------------------------------------------- dense
def inv(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    out = np.empty_like(A)
    rc = lib().ko_inv(n, _p(A), _p(out))
    return rc, out


def pinv(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    out = np.empty_like(A)
    lib().ko_pinv(A.shape[0], _p(A), _p(out))
    return out


def solve(A, b, equilibrate=False):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty_like(b)
    rc = lib().ko_solve(A.shape[0], _p(A), _p(b), _p(x), int(equilibrate))
    return rc, x


# ------------------------------------------------------------------------ ML
def ml_epoch(ranges, anchors, errs, start, use2d=False, variant=0, n_ignore=0, best_mode=0,
             b1_zero_z=False):
    ranges = np.ascontiguousarray(ranges, dtype=np.float64)
    anchors = np.ascontiguousarray(anchors, dtype=np.float64)
    errs = np.ascontiguousarray(np.broadcast_to(errs, ranges.shape), dtype=np.float64)
    start = np.ascontiguousarray(start, dtype=np.float64)
    pos = np.zeros(3)
    cov = np.zeros(9)
    it = C.c_int(0)
    sel = np.zeros(2, dtype=np.int32)
    rc = lib().ko_ml_epoch(len(ranges), _p(ranges), _p(anchors), _p(errs), _p(start), int(use2d),
                           int(variant), int(n_ignore), int(best_mode), int(b1_zero_z), _p(pos),
                           _p(cov), C.byref(it), _p(sel, C.c_int32))
    d = 2 if use2d else 3
    return dict(rc=rc, pos=pos, cov=cov[:d * d].reshape(d, d).copy(), iters=it.value,
                used_mask=int(sel[0]), index=int(sel[1]))


def ml_batch(ranges, anchors, err, start, use2d=False, variant=0, n_ignore=0, best_mode=0,
             threads=0):
    """ranges [M][N] (f64 m / i32 mm / u16 mm); err scalar or [M][N]."""
    ranges = np.ascontiguousarray(ranges)
    M, N = ranges.shape
    anchors = np.ascontiguousarray(anchors, dtype=np.float64)
    err_arr = None if np.isscalar(err) else np.ascontiguousarray(err, dtype=np.float64)
    start = np.ascontiguousarray(start, dtype=np.float64)
    pos = np.zeros((3, N)); cov = np.zeros((9, N))
    iters = np.zeros(N, dtype=np.int32); sel = np.zeros((2, N), dtype=np.int32)
    status = np.zeros(N, dtype=np.int32)
    lib().ko_ml_batch(C.c_int64(N), M, _p(anchors), _vp(ranges), FMT[ranges.dtype],
                      C.c_double(err if err_arr is None else 0.0), _p(err_arr), _p(start),
                      int(use2d), int(variant), int(n_ignore), int(best_mode), _p(pos), _p(cov),
                      _p(iters, C.c_int32), _p(sel, C.c_int32), _p(status, C.c_int32), int(threads))
    return dict(pos=pos, cov=cov, iters=iters, sel=sel, status=status)


# ------------------------------------------------------------------- replays
def t6_replay(x0, P0, ranges, anchors, dt, err, accel_noise=0.5, ignore_worst=False, thr=0.0, variant=0, n_ignore=0,
              best_mode=0, want_traj=False, threads=0):
    """x0 [3][N], P0 [36][N] or None (zeros); ranges [T][M][N]; dt scalar or [T]."""
    ranges = np.ascontiguousarray(ranges)
    T, M, N = ranges.shape
    anchors = np.ascontiguousarray(anchors, dtype=np.float64)
    dt = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, dtype=np.float64), (T,)))
    err_arr = None if np.isscalar(err) else np.ascontiguousarray(err, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, order="C", copy=True).reshape(3, N)
    P = np.zeros((36, N)) if P0 is None else np.array(P0, dtype=np.float64, order="C", copy=True).reshape(36, N)
    traj = np.zeros((T, 3, N)) if want_traj else None
    sel = np.zeros((T, N), dtype=np.int32)
    counters = np.zeros(4)
    status = np.zeros(N, dtype=np.int32)
    lib().ko_t6_replay_sel(C.c_int64(N), T, M, _p(anchors), _p(dt), _vp(ranges), FMT[ranges.dtype],
                       C.c_double(err if err_arr is None else 0.0), _p(err_arr),
                           C.c_double(accel_noise), int(ignore_worst), C.c_double(thr), int(variant), int(n_ignore),
                           int(best_mode), _p(x), _p(P), _p(traj), _p(sel, C.c_int32), _p(counters), _p(status, C.c_int32), int(threads))
    return dict(x=x, P=P, traj=traj, sel=sel, counters=counters, status=status)


def t9_replay(x0, P0, ranges, anchors, dt, err, accel_noise=0.5, jolt=0.5, want_traj=False,
              threads=0):
    ranges = np.ascontiguousarray(ranges)
    T, M, N = ranges.shape
    anchors = np.ascontiguousarray(anchors, dtype=np.float64)
    dt = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, dtype=np.float64), (T,)))
    err_arr = None if np.isscalar(err) else np.ascontiguousarray(err, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, order="C", copy=True).reshape(9, N)
    P = np.zeros((81, N)) if P0 is None else np.array(P0, dtype=np.float64, order="C", copy=True).reshape(81, N)
    traj = np.zeros((T, 3, N)) if want_traj else None
    counters = np.zeros(4)
    status = np.zeros(N, dtype=np.int32)
    lib().ko_t9_replay(C.c_int64(N), T, M, _p(anchors), _p(dt), _vp(ranges), FMT[ranges.dtype],
                       C.c_double(err if err_arr is None else 0.0), _p(err_arr),
                       C.c_double(accel_noise), C.c_double(jolt), _p(x), _p(P), _p(traj),
                       _p(counters), _p(status, C.c_int32), int(threads))
    return dict(x=x, P=P, traj=traj, counters=counters, status=status)


class KoEvent(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("dt", C.c_double), ("offset", C.c_int64),
                ("aux", C.c_double * 9)]


def _events(events):
    arr = (KoEvent * len(events))()
    for i, ev in enumerate(events):
        arr[i].kind, arr[i].dt, arr[i].offset = int(ev[0]), float(ev[1]), int(ev[2])
        if len(ev) > 3 and ev[3] is not None:
            for k, v in enumerate(ev[3]):
                arr[i].aux[k] = float(v)
    return arr


def k8_cfg(accel_noise=0.5, jolt=0.5, **cfg):
    c = KoK8()
    c.accel_noise, c.jolt = accel_noise, jolt
    for k, v in cfg.items():
        setattr(c, k, v)
    return c


def k8_replay(x0, P0, events, ranges, sensors, anchors, err, cfg, b1_zero_z=False, want_traj=False, threads=0,
              tagz=None):
    """x0 [8][N]; ranges [T][M][N] or None; sensors [R][N] or None; events: (kind, dt, offset_row[, aux]).
    cfg.ml_init = 1: filters whose x0 position is NaN initialise themselves from their first epoch with
    rangings (KF.cpp:244-285); `tagz` [N] then carries the per-filter tag height in and out."""
    anchors = np.ascontiguousarray(anchors, dtype=np.float64)
    M = len(anchors)
    x = np.array(x0, dtype=np.float64, order="C", copy=True)
    N = x.shape[-1]
    x = x.reshape(8, N)
    P = np.zeros((64, N)) if P0 is None else np.array(P0, dtype=np.float64, order="C", copy=True).reshape(64, N)
    ranges = None if ranges is None else np.ascontiguousarray(ranges)
    sensors = None if sensors is None else np.ascontiguousarray(sensors, dtype=np.float64)
    err_arr = None if np.isscalar(err) else np.ascontiguousarray(err, dtype=np.float64)
    n_toa = sum(1 for e in events if e[0] == 0)
    traj = np.zeros((n_toa, 3, N)) if want_traj else None
    counters = np.zeros(5)
    status = np.zeros(N, dtype=np.int32)
    arr = _events(events)
    tz = None
    if cfg.ml_init:
        tz = np.full(N, cfg.tag_z) if tagz is None else np.array(tagz, dtype=np.float64, copy=True)
    lib().ko_k8_replay(C.c_int64(N), len(events), arr, M, _p(anchors), _vp(ranges),
                       FMT[ranges.dtype] if ranges is not None else 0,
                       C.c_double(err if err_arr is None else 0.0), _p(err_arr), _p(sensors), C.byref(cfg),
                       int(b1_zero_z), _p(x), _p(P), _p(traj), _p(counters), _p(status, C.c_int32), int(threads),
                       _p(tz))
    return dict(x=x, P=P, traj=traj, counters=counters, status=status, tagz=tz)


def t9_events(x0, P0, events, ranges, sensors, anchors, err, accel_noise=0.5, jolt=0.5, want_traj=False,
              threads=0, variant=0, n_ignore=0, best_mode=0, ml_init=0):
    anchors = np.ascontiguousarray(anchors, dtype=np.float64)
    M = len(anchors)
    x = np.array(x0, dtype=np.float64, order="C", copy=True)
    N = x.shape[-1]
    x = x.reshape(9, N)
    P = np.zeros((81, N)) if P0 is None else np.array(P0, dtype=np.float64, order="C", copy=True).reshape(81, N)
    ranges = None if ranges is None else np.ascontiguousarray(ranges)
    sensors = None if sensors is None else np.ascontiguousarray(sensors, dtype=np.float64)
    err_arr = None if np.isscalar(err) else np.ascontiguousarray(err, dtype=np.float64)
    n_toa = sum(1 for e in events if e[0] == 0)
    traj = np.zeros((n_toa, 3, N)) if want_traj else None
    counters = np.zeros(5)
    status = np.zeros(N, dtype=np.int32)
    arr = _events(events)
    lib().ko_t9_events_sel(C.c_int64(N), len(events), arr, M, _p(anchors), _vp(ranges),
                           FMT[ranges.dtype] if ranges is not None else 0,
                           C.c_double(err if err_arr is None else 0.0), _p(err_arr), _p(sensors),
                           C.c_double(accel_noise), C.c_double(jolt), int(variant), int(n_ignore), int(best_mode),
                           _p(x), _p(P), _p(traj), _p(counters), _p(status, C.c_int32), int(threads), int(ml_init))
    return dict(x=x, P=P, traj=traj, counters=counters, status=status)


# ------------------------------------------------------- single-filter objects
class T6:
    def __init__(self, accel_noise, ignore_worst, thr, p0):
        self.f = KoT6()
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        lib().ko_t6_init(C.byref(self.f), C.c_double(accel_noise), int(ignore_worst),
                         C.c_double(thr), _p(p0))
        self.info = KoInfo()

    def new_toa(self, dt, ranges, anchors, errs):
        ranges = np.ascontiguousarray(ranges, dtype=np.float64)
        anchors = np.ascontiguousarray(anchors, dtype=np.float64)
        errs = np.ascontiguousarray(np.broadcast_to(errs, ranges.shape), dtype=np.float64)
        lib().ko_t6_new_toa(C.byref(self.f), C.c_double(dt), len(ranges), _p(ranges), _p(anchors),
                            _p(errs), C.byref(self.info))
        return self.info

    @property
    def pos(self):
        return np.array(self.f.pos)

    @property
    def P(self):
        return np.array(self.f.P).reshape(6, 6)

    def get_pose(self, dt):
        pos = np.zeros(3); P = np.zeros(36)
        lib().ko_t6_get_pose(C.byref(self.f), C.c_double(dt), _p(pos), _p(P))
        return pos, P.reshape(6, 6)


class K8:
    def __init__(self, accel_noise, init_angle, jolt, p0, **cfg):
        self.f = KoK8()
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        lib().ko_k8_init(C.byref(self.f), C.c_double(accel_noise), C.c_double(init_angle),
                         C.c_double(jolt), _p(p0))
        for k, v in cfg.items():
            setattr(self.f, k, v)
        self.info = KoInfo()

    def new_toa(self, dt, ranges, anchors, errs, b1_zero_z=False):
        ranges = np.ascontiguousarray(ranges, dtype=np.float64)
        anchors = np.ascontiguousarray(anchors, dtype=np.float64)
        errs = np.ascontiguousarray(np.broadcast_to(errs, ranges.shape), dtype=np.float64)
        lib().ko_k8_new_toa(C.byref(self.f), C.c_double(dt), len(ranges), _p(ranges), _p(anchors),
                            _p(errs), int(b1_zero_z), C.byref(self.info))
        return self.info

    def new_px4(self, dt, ix, iy, irz, itime_us, quality):
        lib().ko_k8_new_px4(C.byref(self.f), C.c_double(dt), C.c_double(ix), C.c_double(iy),
                            C.c_double(irz), C.c_double(itime_us), int(quality), C.byref(self.info))
        return self.info

    def new_imu(self, dt, angvel, cov_av, acc, cov_acc):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (angvel, cov_av, acc, cov_acc)]
        lib().ko_k8_new_imu(C.byref(self.f), C.c_double(dt), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]),
                            C.byref(self.info))
        return self.info

    def new_mag(self, dt, mag):
        mag = np.ascontiguousarray(mag, dtype=np.float64)
        lib().ko_k8_new_mag(C.byref(self.f), C.c_double(dt), _p(mag), C.byref(self.info))
        return self.info

    def new_compass(self, dt, compass):
        lib().ko_k8_new_compass(C.byref(self.f), C.c_double(dt), C.c_double(compass),
                                C.byref(self.info))
        return self.info

    @property
    def x(self):
        f = self.f
        return np.array([f.pos[0], f.pos[1], f.vel[0], f.vel[1], f.acc[0], f.acc[1], f.angle, f.omega])

    @property
    def P(self):
        return np.array(self.f.P).reshape(8, 8)

    def get_pose(self, dt):
        x = np.zeros(8); P = np.zeros(64)
        lib().ko_k8_get_pose(C.byref(self.f), C.c_double(dt), _p(x), _p(P))
        return x, P.reshape(8, 8)


class T9:
    def __init__(self, accel_noise, jolt, p0, **cfg):
        self.f = KoT9()
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        lib().ko_t9_init(C.byref(self.f), C.c_double(accel_noise), C.c_double(jolt), _p(p0))
        for k, v in cfg.items():
            setattr(self.f, k, v)
        self.info = KoInfo()

    def new_toa(self, dt, ranges, anchors, errs):
        ranges = np.ascontiguousarray(ranges, dtype=np.float64)
        anchors = np.ascontiguousarray(anchors, dtype=np.float64)
        errs = np.ascontiguousarray(np.broadcast_to(errs, ranges.shape), dtype=np.float64)
        lib().ko_t9_new_toa(C.byref(self.f), C.c_double(dt), len(ranges), _p(ranges), _p(anchors),
                            _p(errs), C.byref(self.info))
        return self.info

    def new_imu(self, dt, acc, cov_acc):
        acc = np.ascontiguousarray(acc, dtype=np.float64)
        cov_acc = np.ascontiguousarray(cov_acc, dtype=np.float64)
        lib().ko_t9_new_imu(C.byref(self.f), C.c_double(dt), _p(acc), _p(cov_acc), C.byref(self.info))
        return self.info

    @property
    def x(self):
        f = self.f
        return np.array(list(f.pos) + list(f.vel) + list(f.acc))

    @property
    def P(self):
        return np.array(self.f.P).reshape(9, 9)

    def get_pose(self, dt):
        x = np.zeros(9); P = np.zeros(81)
        lib().ko_t9_get_pose(C.byref(self.f), C.c_double(dt), _p(x), _p(P))
        return x, P.reshape(9, 9)


# ------------------------------------------------- ranging aggregation (Posgenerator.cpp:143-281)
def assemble(anchor, seq, range_mm, t, n_anchors, max_epochs, err=None, fix_b12=False, first_dt=0.1):
    """N logs of L messages, SoA [L][N] (a single log may be 1-D).  Returns dict(ranges int32
    [T][M][N], err [T][M][N], dt [T][N] (-1 = no epoch), n_epochs [N] (may exceed max_epochs))."""
    anchor = np.ascontiguousarray(np.atleast_2d(np.asarray(anchor, dtype=np.uint8).T).T)
    L, N = anchor.shape
    seq = np.ascontiguousarray(np.asarray(seq, dtype=np.uint8).reshape(L, N))
    range_mm = np.ascontiguousarray(np.asarray(range_mm, dtype=np.int32).reshape(L, N))
    t = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(L, N))
    e = None if err is None else np.ascontiguousarray(np.asarray(err, dtype=np.float64).reshape(L, N))
    M, T = int(n_anchors), int(max_epochs)
    ro = np.empty((T, M, N), dtype=np.int32)
    eo = np.empty((T, M, N))
    dt = np.empty((T, N))
    tt = np.empty((T, N))
    ne = np.empty(N, dtype=np.int32)
    lib().ko_assemble_batch(C.c_int64(N), C.c_int64(L), M, _p(anchor, C.c_uint8), _p(seq, C.c_uint8),
                            _p(range_mm, C.c_int32), _p(e), _p(t), C.c_int64(T), int(fix_b12),
                            C.c_double(first_dt), _p(ro, C.c_int32), _p(eo), _p(dt), _p(ne, C.c_int32), _p(tt))
    return dict(ranges=ro, err=eo, dt=dt, n_epochs=ne, t=tt)


EVENT_ROWS = {1: 5, 2: 3, 3: 2, 4: 1}


def slot_rows(slot_kind, n_anchors):
    """First output row of every slot of a merged schedule: TOA slots count rows of the range tensor
    (n_anchors each), the others rows of the sensor tensor.  Returns (rows int64 [S], range_rows, sensor_rows)."""
    rows, nr, ns = [], 0, 0
    for k in slot_kind:
        if k == 0:
            rows.append(nr); nr += n_anchors
        else:
            rows.append(ns); ns += EVENT_ROWS[int(k)]
    return np.array(rows, dtype=np.int64), nr, ns


def merge_streams(t_epoch, ranges, err, sensors, slot_kind, first_dt=0.1):
    """ko_merge_batch.  t_epoch [T][N], ranges int32 [T][M][N], err [T][M][N] or None; sensors = {kind: (t [L][N],
    payload [L][rows][N])}; slot_kind: the S kinds of the common schedule.  Returns dict(dt [S][N], ranges
    int32 [range_rows][N], err, sensors [sensor_rows][N], n_dropped [N], slot_row)"""
    t_epoch = np.ascontiguousarray(t_epoch, dtype=np.float64)
    ranges = np.ascontiguousarray(ranges, dtype=np.int32)
    T, M, N = ranges.shape
    err = None if err is None else np.ascontiguousarray(err, dtype=np.float64)
    sk = np.ascontiguousarray(slot_kind, dtype=np.int32)
    S = len(sk)
    row, nr, ns = slot_rows(sk, M)
    Ls = (C.c_int64 * 5)(T, 0, 0, 0, 0)
    tp = (C.POINTER(C.c_double) * 5)()
    sp = (C.POINTER(C.c_double) * 5)()
    keep = []
    tp[0] = _p(t_epoch)
    for k, (tk, pk) in sensors.items():
        tk = np.ascontiguousarray(tk, dtype=np.float64)
        pk = np.ascontiguousarray(pk, dtype=np.float64)
        assert pk.shape == (tk.shape[0], EVENT_ROWS[k], N)
        keep += [tk, pk]
        Ls[k] = tk.shape[0]
        tp[k] = _p(tk)
        sp[k] = _p(pk)
    dt = np.empty((S, N))
    ro = np.full((max(nr, 1), N), -1, dtype=np.int32)
    eo = None if err is None else np.zeros((max(nr, 1), N))
    so = np.zeros((max(ns, 1), N))
    nd = np.empty(N, dtype=np.int32)
    lib().ko_merge_batch(C.c_int64(N), M, Ls, tp, _p(ranges, C.c_int32), _p(err), sp, S, _p(sk, C.c_int32),
                         _p(row, C.c_int64), C.c_double(first_dt), _p(dt), _p(ro, C.c_int32), _p(eo), _p(so),
                         _p(nd, C.c_int32))
    return dict(dt=dt, ranges=ro, err=eo, sensors=so, n_dropped=nd, slot_row=row)


# ------------------------------------------------------------------ pose message
def pose_msg(model, x_pred, P_pred, tag_z=0.0):
    """stateToPose + the publisher's read-out (see ko_pose_msg).  x_pred [n] or [n][N], P_pred
    [n][n] or [n*n][N] (row-major).  Returns (pose [13][N], cov [36][N]) (1-D for a single filter)."""
    x = np.asarray(x_pred, dtype=np.float64)
    single = x.ndim == 1
    n = x.shape[0]
    X = x.reshape(n, -1)
    Pm = np.asarray(P_pred, dtype=np.float64).reshape(n * n, -1)
    N = X.shape[1]
    pose = np.empty((13, N)); cov = np.empty((36, N))
    for f in range(N):
        xf = np.ascontiguousarray(X[:, f]); Pf = np.ascontiguousarray(Pm[:, f])
        po = np.empty(13); co = np.empty(36)
        lib().ko_pose_msg(int(model), _p(xf), _p(Pf), C.c_double(tag_z), _p(po), _p(co))
        pose[:, f] = po; cov[:, f] = co
    return (pose[:, 0], cov[:, 0]) if single else (pose, cov)
