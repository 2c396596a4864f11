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


def run_ieee(kflib, a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = [np.empty_like(a) for _ in range(4)]
    flags = np.empty(a.size, dtype=np.int32)
    kflib.check(kflib.lib().kfpos_selftest_ieee(0, a.size, C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data),
                                                *[C.c_void_p(o.ctypes.data) for o in out],
                                                C.c_void_p(flags.ctypes.data)), "kfpos_selftest_ieee")
    return out, flags


def bits_equal(x, y):
    return (x.view(np.uint64) == y.view(np.uint64)) | (np.isnan(x) & np.isnan(y))


def test_branch_free_ieee_division_and_sqrt_are_bit_exact(kflib):
    """kfpos_exact.cu's xf_div / xf_sqrt (the compiler's fast-path sequences, range test accumulated instead of
    branched on) against the plain operators: identical bits on every operand pair -- ordinary magnitudes, values
    next to powers of two, subnormals, huge values, zeros, infinities and NaNs (where the flag must send the
    operation to the plain operator) -- and against numpy's correctly rounded results."""
    rng = np.random.default_rng(7)
    n = 4_000_000
    a = np.concatenate([
        rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-30, 30, n),
        np.ldexp(1.0 + rng.integers(0, 8, n // 4) * 2.0 ** -52, rng.integers(-60, 60, n // 4)),
        rng.uniform(0, 1, n // 4) * 10.0 ** rng.uniform(-320, -290, n // 4),
        rng.uniform(0, 1, n // 4) * 10.0 ** rng.uniform(290, 308, n // 4),
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, 4.0, 2.0 ** -1022, 5e-324, 1.7976931348623157e308] * 10),
    ])
    b = np.concatenate([
        rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-30, 30, n),
        np.ldexp(1.0 + rng.integers(0, 8, n // 4) * 2.0 ** -52, rng.integers(-60, 60, n // 4)),
        rng.uniform(0, 1, n // 4) * 10.0 ** rng.uniform(290, 308, n // 4),
        rng.uniform(0, 1, n // 4) * 10.0 ** rng.uniform(-320, -290, n // 4),
        np.repeat(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, 3.0, 2.0 ** -1022, 5e-324, 1.7976931348623157e308]), 10),
    ])
    (qf, q, sf, s), flags = run_ieee(kflib, a, b)
    assert bits_equal(qf, q).all(), int((~bits_equal(qf, q)).sum())
    assert bits_equal(sf, s).all(), int((~bits_equal(sf, s)).sum())
    with np.errstate(all="ignore"):
        assert bits_equal(q, a / b).all()
        assert bits_equal(s, np.sqrt(a)).all()
    # the fast path is what runs on ordinary operands (positive a for the square root) ...
    ordinary = slice(0, n)
    assert (flags[ordinary] & 1).mean() > 0.999
    assert ((flags[ordinary] & 2) != 0)[a[ordinary] > 0].mean() > 0.999
    # ... and never on operands it cannot round correctly
    special = ~np.isfinite(a) | ~np.isfinite(b) | (b == 0) | (np.abs(a) < 2.0 ** -1000)
    assert not (flags[special] & 1).any()
    assert not (flags[(a <= 0) | ~np.isfinite(a) | (a < 2.0 ** -1000)] & 2).any()
