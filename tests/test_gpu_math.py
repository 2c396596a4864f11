"""GPU: accuracy of the kernels' elementary functions (kfpos_math.cuh) against IEEE / libm."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ulps(a, b):
    """|a - b| in units in the last place of the double nearest to b (b: long double reference)."""
    b = np.asarray(b, dtype=np.longdouble)
    return np.asarray(np.abs(np.asarray(a, dtype=np.longdouble) - b) / np.spacing(np.abs(b.astype(np.float64))),
                      dtype=np.float64)


def ld(x):
    return np.asarray(x, dtype=np.longdouble)


def run(kflib, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = [np.empty_like(x) for _ in range(4)]
    kflib.check(kflib.lib().kfpos_selftest_math(0, x.size, C.c_void_p(x.ctypes.data),
                                                *[C.c_void_p(o.ctypes.data) for o in out]), "kfpos_selftest_math")
    return out


def test_rcp_rsqrt_within_one_ulp(kflib):
    rng = np.random.default_rng(0)
    # distances^2 in m^2, innovation variances, determinants: many decades, both signs for rcp
    mag = 10.0 ** rng.uniform(-12, 12, 400000)
    rcp, rsq, _, _ = run(kflib, mag)
    print('max ulps rcp', ulps(rcp, 1 / ld(mag)).max(), 'rsqrt', ulps(rsq, 1 / np.sqrt(ld(mag))).max())
    assert ulps(rcp, 1 / ld(mag)).max() <= 1.0
    assert ulps(rsq, 1 / np.sqrt(ld(mag))).max() <= 1.0
    rcp_neg, _, _, _ = run(kflib, -mag)
    assert ulps(rcp_neg, -1 / ld(mag)).max() <= 1.0
    edge = np.array([1.0, 2.0, 0.5, 4.0, 1e-300, 1e300, np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0)])
    rcp, rsq, _, _ = run(kflib, edge)
    assert ulps(rcp, 1 / ld(edge)).max() <= 1.0 and ulps(rsq, 1 / np.sqrt(ld(edge))).max() <= 1.0


def test_sincos_within_one_ulp(kflib):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-8, 8, 400000), rng.uniform(-1e4, 1e4, 100000), rng.normal(0, 1e-3, 50000),
                        np.array([0.0, np.pi, -np.pi, np.pi / 2, np.pi / 4, 1e5, -1e5, 3e7, 1e300])])
    _, _, sn, cs = run(kflib, x)
    ref_s, ref_c = np.sin(x), np.cos(x)
    # ulps of the result, except next to a zero of the function where 1 ulp of the ARGUMENT decides
    tol = np.maximum(1.0 * np.spacing(np.abs(ref_s)), 0.5 * np.spacing(np.abs(x)))
    assert np.all(np.abs(sn - ref_s) <= tol), float((np.abs(sn - ref_s) / tol).max())
    tol = np.maximum(1.0 * np.spacing(np.abs(ref_c)), 0.5 * np.spacing(np.abs(x)))
    assert np.all(np.abs(cs - ref_c) <= tol), float((np.abs(cs - ref_c) / tol).max())
    assert np.abs(sn ** 2 + cs ** 2 - 1).max() < 5e-16
