// kfpos_math.cuh -- small FP64 building blocks shared by the kernels.
//
// Everything here is per-thread register arithmetic: the matrices of this path
// are 2x2 .. 9x9 with structural sparsity and data-dependent control flow, so
// they run on the FP64 CUDA-core pipe (DFMA), not on tensor cores.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace kfpos {

#define KF_DEV __device__ __forceinline__

// Anchor table, passed by value inside the kernel parameter block: parameters
// live in the constant bank, and every lane of a warp reads the same anchor at
// the same time, so each coordinate is a single broadcast constant-cache read
// that the compiler folds into the DADD/DFMA operand.
struct AnchorTable {
    double x[32], y[32], z[32];
    int n;
    int _pad;
};

// Packed symmetric matrix (lower triangle, row-major).  All indices are
// compile-time constants after unrolling, so `a` lives in registers.
template <int N>
struct Sym {
    static constexpr int SZ = N * (N + 1) / 2;
    double a[SZ];
    KF_DEV double &at(int i, int j) { return i >= j ? a[i * (i + 1) / 2 + j] : a[j * (j + 1) / 2 + i]; }
    KF_DEV double get(int i, int j) const { return i >= j ? a[i * (i + 1) / 2 + j] : a[j * (j + 1) / 2 + i]; }
};

// wire range -> metres, exactly `(double) ranges[i] / 1000` (PG.cpp:484) without a
// division: q = mm*0.001 corrected once with the exact FMA residual.  Verified
// bit-identical to the IEEE quotient for every positive int32.
KF_DEV double mm_to_m(double mm) {
    const double q = mm * 0.001;
    return fma(fma(-q, 1000.0, mm), 0.001, q);
}

// Reciprocal and reciprocal square root for well-scaled arguments (distances in
// metres, innovation variances, determinants): MUFU seed (relative error e <= 2^-20,
// the unit reads the upper 32 bits of the operand) + ONE cubic correction on the FP64 pipe
//   1/x       = y (1 + e + e^2)            + O(e^3),  e = 1 - x y
//   1/sqrt(x) = y (1 + e/2 + 3 e^2 / 8)    + O(e^3),  e = 1 - x y^2
// => <= 1 ulp (checked against IEEE division / sqrt by kfpos_selftest_math), 3 / 5 FP64
// instructions, no slow-path branch.  0 -> inf/NaN like the IEEE operations they replace;
// denormal arguments are flushed (never reached by this path's quantities).
KF_DEV double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

KF_DEV double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// 1/sqrt(x) for an unrolled anchor slot that may hold no ranging: `on` false seeds the refinement with rsqrt(+inf) = 0,
// and every term of the correction is a multiple of the seed, so the result is exactly 0 -- one 32-bit select on the
// operand's high word (all the MUFU unit reads) instead of two on the 64-bit result.
KF_DEV double fast_rsqrt_masked(double x, bool on) {
    const double xs = __hiloint2double(on ? __double2hiint(x) : 0x7ff00000, __double2loint(x));
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// sin and cos together for the angles of this path (heading, omega * dt: a few radians).
// |x| < 1e5: quadrant k = rint(x 2/pi), Cody-Waite reduction with a three-part pi/2 in FMAs
// (error ~ |k| 2^-110), then the fdlibm kernels on [-pi/4, pi/4] (< 1 ulp): ~27 FP64
// instructions and a handful of selects instead of the ~200 instructions of the library
// sincos with its Payne-Hanek branch; larger or non-finite arguments take the library path.
// (the library path out of line: inlined, its Payne-Hanek reduction costs ~150 instructions and a stack frame at
// every call site of a kernel whose code is already three times the instruction cache)
static __device__ __noinline__ double2 library_sincos(double x) {
    double s, c;
    sincos(x, &s, &c);
    return make_double2(s, c);
}
KF_DEV void fast_sincos(double x, double *sn, double *cs) {
    if (!(fabs(x) < 1.0e5)) {
        const double2 sc = library_sincos(x);
        *sn = sc.x;
        *cs = sc.y;
        return;
    }
    const double kd = rint(x * 0.63661977236758138243); // 2/pi
    const int q = (int)kd;
    double r = fma(-kd, 1.5707963267948966, x);
    r = fma(-kd, 6.123233995736766e-17, r);
    r = fma(-kd, -1.4973849048591698e-33, r);
    const double z = r * r;
    // __kernel_sin(r, 0)
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    const double s = fma(z * r, fma(z, ps, -1.66666666666666324348e-01), r);
    // __kernel_cos(r, 0)
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double hz = 0.5 * z, w = 1.0 - hz;
    const double c = w + (((1.0 - w) - hz) + z * (z * pc));
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

// one column of a per-thread array kept in shared memory: element i of thread t
// lives at base[i * stride + t] (conflict-free: consecutive lanes, consecutive words)
struct Col {
    double *p;
    int stride;
    KF_DEV double &operator[](int i) const { return p[i * stride]; }
};

KF_DEV double load_range(const void *base, int fmt, int64_t idx) {
    switch (fmt) {
    case 0: return __ldg(reinterpret_cast<const double *>(base) + idx);
    case 1: return mm_to_m((double)__ldg(reinterpret_cast<const int32_t *>(base) + idx));
    default: return mm_to_m((double)__ldg(reinterpret_cast<const uint16_t *>(base) + idx));
    }
}

// normalizeAngle, KF.cpp:699-706 (single wrap)
KF_DEV double wrap_angle(double a) {
    const double PI = 3.14159265358979323846;
    if (a > PI) return a - 2 * PI;
    else if (a <= -PI) return a + 2 * PI;
    return a;
}

// A determinant the dense solvers of the reference accept: non-zero and finite.  An error estimate
// of 0 (a ranging without errorEstimation in per-measurement mode) puts 1/0 into the normal matrix;
// arma::solve / inv reject that like an exactly singular matrix (update skipped, status SINGULAR).
KF_DEV bool usable_det(double det) {
    const double a = fabs(det);
    return a > 0.0 && a <= 1.7976931348623157e308;
}

// Symmetric 3x3 solve H s = g by cofactors.  H packed [xx, xy, yy, xz, yz, zz]
// (Sym<3> order).  Returns false when det is 0, infinite or NaN.
KF_DEV bool solve_sym3(const double (&H)[6], const double (&g)[3], double (&s)[3]) {
    const double a = H[0], b = H[1], c = H[3], d = H[2], e = H[4], f = H[5];
    // [a b c; b d e; c e f]
    const double c00 = d * f - e * e;
    const double c01 = c * e - b * f;
    const double c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    if (!usable_det(det)) return false;
    const double c11 = a * f - c * c;
    const double c12 = b * c - a * e;
    const double c22 = a * d - b * b;
    const double id = fast_rcp(det);
    s[0] = (c00 * g[0] + c01 * g[1] + c02 * g[2]) * id;
    s[1] = (c01 * g[0] + c11 * g[1] + c12 * g[2]) * id;
    s[2] = (c02 * g[0] + c12 * g[1] + c22 * g[2]) * id;
    return true;
}

// inverse of a symmetric 3x3 (packed as above) -> packed; false if singular
KF_DEV bool inv_sym3(const double (&H)[6], double (&I)[6]) {
    const double a = H[0], b = H[1], c = H[3], d = H[2], e = H[4], f = H[5];
    const double c00 = d * f - e * e;
    const double c01 = c * e - b * f;
    const double c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    if (!usable_det(det)) return false;
    const double id = fast_rcp(det);
    I[0] = c00 * id;
    I[1] = c01 * id;
    I[2] = (a * f - c * c) * id;
    I[3] = c02 * id;
    I[4] = (b * c - a * e) * id;
    I[5] = (a * d - b * b) * id;
    return true;
}

// Symmetric 2x2 solve [a b; b d] s = g
KF_DEV bool solve_sym2(double a, double b, double d, double g0, double g1, double &s0, double &s1) {
    const double det = a * d - b * b;
    if (!usable_det(det)) return false;
    const double id = fast_rcp(det);
    s0 = (d * g0 - b * g1) * id;
    s1 = (a * g1 - b * g0) * id;
    return true;
}

// Sequential scalar measurement update of the linear-KF form used inside one
// IEKF iteration: prior (0, P), row h with non-zeros at the
// columns set in MASK, innovation reference y, variance R:
//   s = h P h^T + R ; k = P h^T / s ; dx += k (y - h dx) ; P -= k (P h^T)^T
// After all rows of an iteration dx = K (eps - J delta) and P = (I - K J) P^-,
// identical (in exact arithmetic, R block-diagonal) to the reference's dense
// K = P J^T inv(J P J^T + R)  (TOA.cpp:316-319, KF.cpp:491-495, TOAIMU.cpp:330-334).
template <int N, unsigned MASK>
KF_DEV void scalar_update(Sym<N> &P, double (&dx)[N], const double (&h)[N], double y, double R) {
    double ph[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        bool first = true;
#pragma unroll
        for (int j = 0; j < N; ++j)
            if ((MASK >> j) & 1u) {
                s = first ? P.get(i, j) * h[j] : fma(P.get(i, j), h[j], s);
                first = false;
            }
        ph[i] = s;
    }
    double s = R, nu = y;
#pragma unroll
    for (int j = 0; j < N; ++j)
        if ((MASK >> j) & 1u) {
            s = fma(h[j], ph[j], s);
            nu = fma(-h[j], dx[j], nu);
        }
    const double inv_s = fast_rcp(s);
    const double g = nu * inv_s;
#pragma unroll
    for (int i = 0; i < N; ++i) dx[i] = fma(ph[i], g, dx[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double k = ph[i] * inv_s;
#pragma unroll
        for (int j = 0; j <= i; ++j) P.at(i, j) = fma(-k, ph[j], P.at(i, j));
    }
}

// 2-row block update for a correlated pair (K8 IMU accel block, KF.cpp:431-434):
// rows h0,h1 (non-zeros in MASK), innovations y0,y1, R = [[r00,r01],[r01,r11]].
template <int N, unsigned MASK>
KF_DEV void block2_update(Sym<N> &P, double (&dx)[N], const double (&h0)[N], const double (&h1)[N],
                          double y0, double y1, double r00, double r01, double r11) {
    double p0[N], p1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j)
            if ((MASK >> j) & 1u) {
                s0 = fma(P.get(i, j), h0[j], s0);
                s1 = fma(P.get(i, j), h1[j], s1);
            }
        p0[i] = s0;
        p1[i] = s1;
    }
    double s00 = r00, s01 = r01, s11 = r11, n0 = y0, n1 = y1;
#pragma unroll
    for (int j = 0; j < N; ++j)
        if ((MASK >> j) & 1u) {
            s00 = fma(h0[j], p0[j], s00);
            s01 = fma(h0[j], p1[j], s01);
            s11 = fma(h1[j], p1[j], s11);
            n0 = fma(-h0[j], dx[j], n0);
            n1 = fma(-h1[j], dx[j], n1);
        }
    const double id = fast_rcp(s00 * s11 - s01 * s01);
    const double i00 = s11 * id, i01 = -s01 * id, i11 = s00 * id;
    const double g0 = i00 * n0 + i01 * n1, g1 = i01 * n0 + i11 * n1;
#pragma unroll
    for (int i = 0; i < N; ++i) dx[i] = fma(p0[i], g0, fma(p1[i], g1, dx[i]));
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double k0 = p0[i] * i00 + p1[i] * i01;
        const double k1 = p0[i] * i01 + p1[i] * i11;
#pragma unroll
        for (int j = 0; j <= i; ++j) P.at(i, j) = fma(-k0, p0[j], fma(-k1, p1[j], P.at(i, j)));
    }
}

// per-thread work counters of one replay (summed into the device counters at the end)
// device counters (unsigned long long each)
enum { CNT_UPDATES = 0, CNT_ML_ITERS, CNT_COST_EVALS, CNT_GAIN_EVALS, CNT_BAD, CNT_IGNORED, CNT_ML_CAPPED, CNT_ML_CYCLES, CNT_N = 8 };

struct StepStats {
    unsigned ml_iters, cost_evals, gain_evals, status;
};

// warp-level sum of a per-thread counter, one atomic per warp
KF_DEV void warp_accumulate(unsigned long long *dst, unsigned v) {
    unsigned s = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(dst, (unsigned long long)s);
}

} // namespace kfpos
