"""ctypes binding of oracle/_ref/libkfref.so: the reference's OWN .cpp files compiled
against the shim headers (oracle/shim) with LAPACK from scipy's bundled OpenBLAS.

TEST INFRASTRUCTURE ONLY (pins the oracle; optional "reference" CPU baseline).
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libkfref.so")
_LIB = None

XML_DEFAULT = {
    "kfpos_pos": '<config><algorithm type="0" variant="0"/></config>',
    "kfpos_px4": '<config><px4flow armP0="1" armP1="0" useFixedSensorHeight="1" sensorHeight="5" '
                 'sensorInitAngle="-1.570796326794897" covarianceVelocity="0.04" covarianceGyroZ="0.02"/></config>',
    "kfpos_tag": '<config><uwb useFixedHeight="0" fixedHeight="1.049" tagId="0"/></config>',
    "kfpos_imu": '<config><imu useFixedCovarianceAcceleration="1" covarianceAcceleration="0.003" '
                 'useFixedCovarianceAngularVelocityZ="1" covarianceAngularVelocityZ="0.089"/></config>',
    "kfpos_mag": '<config><mag angleOffset="0" covarianceMag="0.0001"/></config>',
}


def find_lapack():
    try:
        import scipy
        base = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    except ImportError:
        return None
    c = sorted(glob.glob(os.path.join(base, "libscipy_openblas-*.so")))
    return c[0] if c else None


def available() -> bool:
    return os.path.exists(SO) and find_lapack() is not None


def lib():
    global _LIB
    if _LIB is None:
        l = C.CDLL(SO)
        for n in ("ref_ml_create", "ref_t6_create", "ref_k8_create", "ref_t9_create", "ref_k8_create_nofix",
                  "ref_t9_create_nofix"):
            getattr(l, n).restype = C.c_void_p
        l.ref_k8_tag_z.restype = C.c_double
        if l.ref_init(find_lapack().encode()) != 0:
            raise RuntimeError("libkfref: cannot load LAPACK from scipy's OpenBLAS")
        _LIB = l
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _arr(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else np.ascontiguousarray(np.broadcast_to(a, shape))


def ns(dt: float) -> int:
    """dt in integer nanoseconds (the fake clock's tick); use dt values that are exact ns multiples."""
    return int(round(dt * 1e9))


def valid_only(ranges, anchors, errs):
    """PosGenerator only forwards slots with range > 0 (PG.cpp:481-489)."""
    ranges = np.asarray(ranges, dtype=np.float64)
    keep = ranges > 0
    return (np.ascontiguousarray(ranges[keep]), np.ascontiguousarray(np.asarray(anchors, dtype=np.float64)[keep]),
            np.ascontiguousarray(np.broadcast_to(np.asarray(errs, dtype=np.float64), ranges.shape)[keep]))


class RefML:
    def __init__(self, use2d, variant, n_ignore, start):
        self.h = C.c_void_p(lib().ref_ml_create(int(use2d), int(variant), int(n_ignore),
                                                C.c_double(start[0]), C.c_double(start[1]), C.c_double(start[2])))

    def solve(self, ranges, anchors, errs, mode=1):
        r, a, e = valid_only(ranges, anchors, errs)
        pos = np.zeros(3); cov = np.zeros(9); d = C.c_int(0)
        rc = lib().ref_ml_solve(self.h, int(mode), len(r), _p(r), _p(a), _p(e), _p(pos), _p(cov), C.byref(d))
        dd = d.value
        return dict(rc=rc, pos=pos, cov=cov[:dd * dd].reshape(dd, dd).copy() if dd else np.zeros((0, 0)))

    def __del__(self):
        if _LIB is not None and self.h:
            _LIB.ref_ml_destroy(self.h)


class RefT6:
    def __init__(self, accel_noise, ignore_worst, thr, p0):
        self.h = C.c_void_p(lib().ref_t6_create(C.c_double(accel_noise), int(ignore_worst), C.c_double(thr),
                                                C.c_double(p0[0]), C.c_double(p0[1]), C.c_double(p0[2])))

    def new_toa(self, dt, ranges, anchors, errs):
        r, a, e = valid_only(ranges, anchors, errs)
        return lib().ref_t6_toa(self.h, C.c_longlong(ns(dt)), len(r), _p(r), _p(a), _p(e))

    def state(self):
        pos = np.zeros(3); P = np.zeros(36)
        lib().ref_t6_get(self.h, _p(pos), _p(P))
        return pos, P.reshape(6, 6)

    def get_pose(self, dt):
        pos = np.zeros(3); P = np.zeros(36)
        rc = lib().ref_t6_get_pose(self.h, C.c_longlong(ns(dt)), _p(pos), _p(P))
        return rc, pos, P.reshape(6, 6)

    def __del__(self):
        if _LIB is not None and self.h:
            _LIB.ref_t6_destroy(self.h)


class RefK8:
    def __init__(self, accel_noise, init_angle, jolt, p0, xml=None):
        params = dict(XML_DEFAULT)
        params.update(xml or {})
        for k, v in params.items():
            lib().ref_set_param(k.encode(), v.encode())
        if p0 is None:  # the constructor without initialPosition: ML initialisation (KF.cpp:244-285)
            h = lib().ref_k8_create_nofix(C.c_double(accel_noise), C.c_double(init_angle), C.c_double(jolt))
        else:
            h = lib().ref_k8_create(C.c_double(accel_noise), C.c_double(init_angle), C.c_double(jolt),
                                    C.c_double(p0[0]), C.c_double(p0[1]), C.c_double(p0[2] if len(p0) > 2 else 0.0))
        if not h:
            raise RuntimeError("KalmanFilter::init() failed")
        self.h = C.c_void_p(h)

    def new_toa(self, dt, ranges, anchors, errs):
        r, a, e = valid_only(ranges, anchors, errs)
        return lib().ref_k8_toa(self.h, C.c_longlong(ns(dt)), len(r), _p(r), _p(a), _p(e))

    def new_px4(self, dt, ix, iy, irz, itime_us, quality):
        return lib().ref_k8_px4(self.h, C.c_longlong(ns(dt)), C.c_double(ix), C.c_double(iy), C.c_double(irz),
                                C.c_double(itime_us), int(quality))

    def new_imu(self, dt, angvel, cov_av, acc, cov_acc):
        a = [_arr(v) for v in (angvel, cov_av, acc, cov_acc)]
        return lib().ref_k8_imu(self.h, C.c_longlong(ns(dt)), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]))

    def new_mag(self, dt, mag):
        m = _arr(mag)
        return lib().ref_k8_mag(self.h, C.c_longlong(ns(dt)), _p(m))

    def new_compass(self, dt, compass):
        return lib().ref_k8_compass(self.h, C.c_longlong(ns(dt)), C.c_double(compass))

    def state(self):
        x = np.zeros(8); P = np.zeros(64)
        lib().ref_k8_get(self.h, _p(x), _p(P))
        return x, P.reshape(8, 8)

    def tag_z(self):
        return float(lib().ref_k8_tag_z(self.h))

    def __del__(self):
        if _LIB is not None and self.h:
            _LIB.ref_k8_destroy(self.h)


class RefT9:
    def __init__(self, accel_noise, jolt, p0):
        if p0 is None:  # the constructor without initialPosition (TOAIMU.cpp:6-24)
            self.h = C.c_void_p(lib().ref_t9_create_nofix(C.c_double(accel_noise), C.c_double(jolt)))
        else:
            self.h = C.c_void_p(lib().ref_t9_create(C.c_double(accel_noise), C.c_double(jolt), C.c_double(p0[0]),
                                                    C.c_double(p0[1]), C.c_double(p0[2])))

    def new_toa(self, dt, ranges, anchors, errs):
        r, a, e = valid_only(ranges, anchors, errs)
        return lib().ref_t9_toa(self.h, C.c_longlong(ns(dt)), len(r), _p(r), _p(a), _p(e))

    def new_imu(self, dt, acc, cov_acc):
        a, c = _arr(acc), _arr(cov_acc)
        return lib().ref_t9_imu(self.h, C.c_longlong(ns(dt)), _p(a), _p(c))

    def state(self):
        x = np.zeros(9); P = np.zeros(81)
        lib().ref_t9_get(self.h, _p(x), _p(P))
        return x, P.reshape(9, 9)

    def __del__(self):
        if _LIB is not None and self.h:
            _LIB.ref_t9_destroy(self.h)


def pose_msg(obj, dt):
    """(rc, pose13, cov36) of getPose `dt` after the last update, read as the publisher reads it
    (Posgenerator.cpp:385-470)."""
    kind = {RefT6: 1, RefK8: 2, RefT9: 3}[type(obj)]
    pose = np.zeros(13); cov = np.zeros(36)
    rc = lib().ref_get_pose_msg(obj.h, kind, C.c_longlong(ns(dt)), _p(pose), _p(cov))
    return rc, pose, cov


def t6_replay(x0, ranges_m, anchors, dt, err, accel_noise=0.5):
    """Batch driver (single thread): ranges f64 metres [T][M][N], x0 [3][N]."""
    ranges_m = np.ascontiguousarray(ranges_m, dtype=np.float64)
    T, M, N = ranges_m.shape
    anchors = _arr(anchors)
    x0 = _arr(x0).reshape(3, N)
    x = np.zeros((3, N)); P = np.zeros((36, N))
    rc = lib().ref_t6_replay(C.c_longlong(N), T, M, _p(anchors), C.c_longlong(ns(dt)), _p(ranges_m),
                             C.c_double(err), C.c_double(accel_noise), _p(x0), _p(x), _p(P))
    return dict(rc=rc, x=x, P=P)


class RefPosGenerator:
    """The reference's PosGenerator (publishers/Posgenerator.cpp, unmodified) behind a recording
    algorithm: feed a ranging log, read back the epochs it hands to newTOAMeasurement and the
    report the node would publish.  algorithm: -1 = record only, 2 = ML, 5 = KalmanFilterTOA,
    6 = KalmanFilterTOAIMU (Posgenerator.h:62-68)."""

    def __init__(self, anchors, algorithm=-1, tag_id=0, anchor_ids=None, accel_noise=0.5, jolt=0.5,
                 use_start=False, start=(0.0, 0.0, 0.0), start_angle=0.0, ignore_worst=False, cost_threshold=0.0,
                 use2d=False, variant=0, n_ignore=0, xml=None):
        anchors = _arr(anchors)
        self.M = anchors.shape[0]
        ids = np.arange(self.M, dtype=np.int32) if anchor_ids is None else np.ascontiguousarray(anchor_ids, np.int32)
        l = lib()
        if algorithm == 4:  # KalmanFilter reads its XML documents from the parameter server
            params = dict(XML_DEFAULT)
            params.update(xml or {})
            for k, v in params.items():
                l.ref_set_param(k.encode(), v.encode())
        l.ref_pg_create.restype = C.c_void_p
        l.ref_pg_epoch_times.restype = C.c_longlong
        l.ref_pg_feed.restype = C.c_longlong
        l.ref_pg_epochs.restype = C.c_longlong
        self.h = C.c_void_p(l.ref_pg_create(int(algorithm), int(tag_id), self.M, _p(anchors),
                                            ids.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(accel_noise),
                                            C.c_double(jolt), int(use_start), _p(_arr(start)),
                                            C.c_double(start_angle), int(ignore_worst),
                                            C.c_double(cost_threshold), int(use2d), int(variant), int(n_ignore)))

    def close(self):
        if self.h:
            lib().ref_pg_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def feed(self, anchor_id, range_mm, seq, t, err=None, tag_id=None, flush_tail=True):
        ip = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        a, r, s = ip(anchor_id), ip(range_mm), ip(seq)
        tg = np.zeros_like(a) if tag_id is None else ip(tag_id)
        t = _arr(t)
        e = None if err is None else _arr(err)
        i32 = lambda x: x.ctypes.data_as(C.POINTER(C.c_int))
        return int(lib().ref_pg_feed(self.h, C.c_longlong(len(a)), i32(a), i32(tg), i32(r), i32(s),
                                     _p(e) if e is not None else None, _p(t), int(flush_tail)))

    def sensor(self, kind, t, values):
        """kind 1 px4 / 2 imu / 3 mag / 4 compass (see ref_pg_sensor)."""
        return int(lib().ref_pg_sensor(self.h, int(kind), C.c_double(t), _p(_arr(values))))

    def epoch_times(self, max_epochs):
        t = np.zeros(max_epochs)
        n = int(lib().ref_pg_epoch_times(self.h, C.c_longlong(max_epochs), _p(t)))
        return t[:min(n, max_epochs)]

    def epochs(self, max_epochs):
        r = np.zeros((max_epochs, self.M)); e = np.zeros((max_epochs, self.M)); lag = np.zeros(max_epochs)
        n = int(lib().ref_pg_epochs(self.h, C.c_longlong(max_epochs), self.M, _p(r), _p(e), _p(lag)))
        k = min(n, max_epochs)
        return dict(n=n, ranges=r[:k], err=e[:k], time_lag=lag[:k])

    def report(self, t):
        pose = np.zeros(13); cov = np.zeros(36)
        rc = lib().ref_pg_report(self.h, C.c_double(t), _p(pose), _p(cov))
        return rc, pose, cov

    def errors(self):
        return int(lib().ref_pg_errors(self.h))
