// kfpos_ml.cuh -- per-thread ML (Newton) multilateration, the B200 formulation of
// MLLocation::estimatePosition / estimatePosition2D (ML.cpp:48-257).
//
// One thread owns one epoch.  Each Newton iteration is ONE pass over the anchor
// table that yields, at the current point, the stopping cost, the unweighted SSE
// (the EKFs' mlRangingError), the gradient and the Hessian together -- the
// reference evaluates the distances three times per iteration (ML.cpp:171,216,229).
// The anchor loop is NOT unrolled (the whole replay kernel has to stay inside the
// 32 KB instruction cache); the epoch's rangings live in a shared-memory column.
#pragma once
#include "kfpos_math.cuh"

namespace kfpos {

// Valid rangings of one epoch: z[i] metres (shared-memory column) for the slots
// set in `valid`.  PME = per-measurement errorEstimation column `e`; otherwise
// the scalar e0 applies to every ranging.
template <bool PME>
struct Epoch {
    Col z;
    Col e;
    double e0;
    unsigned valid;
    int m_slots;
    KF_DEV double err(int i) const { return PME ? e[i] : e0; }
};

struct MlPass3 {
    double wcost, sse;
    double g[3];
    double H[6]; // packed Sym<3>: xx, xy, yy, xz, yz, zz
};

// gradient / Hessian / costs at p over the slots in `mask` (ML.cpp:171-222).
// With a scalar errorEstimation the weight 1/e is common to g and H and cancels in
// the Newton step, so only the stop-test cost carries it.
// STORE: also park 1/d_i and r_i - d_i in the scratch columns -- the first pass runs at the
// predicted position, exactly where the IEKF's first cost evaluation needs them.
struct DistStore {
    Col invd, eps;
};

template <bool PME, bool STORE = false>
KF_DEV void ml_pass3(const AnchorTable &A, const Epoch<PME> &ep, unsigned mask, const double (&p)[3],
                     MlPass3 &o, const DistStore *ds = nullptr) {
    double wc = 0.0, sse = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
    double h0 = 0.0, h1 = 0.0, h2 = 0.0, h3 = 0.0, h4 = 0.0, h5 = 0.0, c1s = 0.0;
#pragma unroll 2
    for (int i = 0; i < ep.m_slots; ++i) {
        if (!((mask >> i) & 1u)) continue;
        const double dx = A.x[i] - p[0], dy = A.y[i] - p[1], dz = A.z[i] - p[2];
        const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
        const double invd = fast_rsqrt(d2);
        const double d = d2 * invd;
        const double r = ep.z[i];
        const double res = r - d;
        const double rid = r * invd;
        if (STORE) {
            ds->invd[i] = invd;
            ds->eps[i] = res;
        }
        if (PME) {
            const double w = 1.0 / ep.e[i];
            sse = fma(res, res, sse);
            wc = fma(res * res, w, wc);
            const double t = res * invd * w;
            g0 = fma(t, dx, g0); g1 = fma(t, dy, g1); g2 = fma(t, dz, g2);
            c1s = fma(1.0 - rid, w, c1s);
            const double c2 = rid * invd * invd * w;
            const double cx = c2 * dx, cy = c2 * dy, cz = c2 * dz;
            h0 = fma(cx, dx, h0); h1 = fma(cx, dy, h1); h2 = fma(cy, dy, h2);
            h3 = fma(cx, dz, h3); h4 = fma(cy, dz, h4); h5 = fma(cz, dz, h5);
        } else {
            sse = fma(res, res, sse);
            const double t = res * invd;
            g0 = fma(t, dx, g0); g1 = fma(t, dy, g1); g2 = fma(t, dz, g2);
            c1s += 1.0 - rid;
            const double c2 = rid * invd * invd;
            const double cx = c2 * dx, cy = c2 * dy, cz = c2 * dz;
            h0 = fma(cx, dx, h0); h1 = fma(cx, dy, h1); h2 = fma(cy, dy, h2);
            h3 = fma(cx, dz, h3); h4 = fma(cy, dz, h4); h5 = fma(cz, dz, h5);
        }
    }
    o.sse = sse;
    o.wcost = PME ? wc : sse * (1.0 / ep.e0);
    o.g[0] = g0; o.g[1] = g1; o.g[2] = g2;
    o.H[0] = h0 + c1s; o.H[1] = h1; o.H[2] = h2 + c1s; o.H[3] = h3; o.H[4] = h4; o.H[5] = h5 + c1s;
}

// return codes of the ML solvers
#define ML_OK 0
#define ML_FEW 1       // fewer than minRangings: position = start (ML.cpp:54-58,158-161)
#define ML_SINGULAR -1 // arma::solve / inv would throw

// estimatePosition (3-D), ML.cpp:153-257.  p: in = start, out = estimate.
// sse_out = estimationError at the returned point.  The covariance
// inv(J^T W^-1 J) (ML.cpp:229-254) is produced only when cov != nullptr.
template <bool PME, bool STORE = false>
KF_DEV int ml_solve3(const AnchorTable &A, const Epoch<PME> &ep, unsigned mask, double (&p)[3],
                     double &sse_out, unsigned &iters, double *cov /* packed Sym<3> or null */,
                     const DistStore *ds = nullptr, double *sse_start = nullptr) {
    MlPass3 ps;
    ml_pass3<PME, STORE>(A, ep, mask, p, ps, ds);
    sse_out = ps.sse;
    if (sse_start) *sse_start = ps.sse;
    if (__popc(mask) < 4) return ML_FEW;
    double cost = 1e20, newCost = 1.0;
    unsigned iter = 0;
    while ((fabs(cost - newCost) / cost > 1e-3) && (iter < 10000u)) {
        iter += 1;
        cost = newCost;
        double s[3];
        if (!solve_sym3(ps.H, ps.g, s)) { iters += iter; return ML_SINGULAR; }
        // newPos = solve(H, H pos - g)  ==  pos - H^-1 g
        p[0] -= s[0]; p[1] -= s[1]; p[2] -= s[2];
        ml_pass3<PME>(A, ep, mask, p, ps);
        newCost = ps.wcost;
    }
    iters += iter;
    sse_out = ps.sse;
    if (cov) {
        // J_i = (p - b_i)/d_i ; W = diag(max(e_i, SSE)) ; cov = inv(J^T W^-1 J)
        double M[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < ep.m_slots; ++i) {
            if (!((mask >> i) & 1u)) continue;
            const double dx = p[0] - A.x[i], dy = p[1] - A.y[i], dz = p[2] - A.z[i];
            const double id2 = 1.0 / (dx * dx + dy * dy + dz * dz);
            const double w = id2 / fmax(ep.err(i), ps.sse);
            M[0] = fma(w * dx, dx, M[0]);
            M[1] = fma(w * dx, dy, M[1]);
            M[2] = fma(w * dy, dy, M[2]);
            M[3] = fma(w * dx, dz, M[3]);
            M[4] = fma(w * dy, dz, M[4]);
            M[5] = fma(w * dz, dz, M[5]);
        }
        double I[6];
        if (!inv_sym3(M, I)) return ML_SINGULAR;
#pragma unroll
        for (int k = 0; k < 6; ++k) cov[k] = I[k];
    }
    return ML_OK;
}

struct MlPass2 {
    double sse;
    double g[2];
    double H[3]; // xx, xy, yy
};

// 2-D pass: distances are 3-D with z fixed, derivatives in x,y only (ML.cpp:74-95)
template <bool PME, bool STORE = false>
KF_DEV void ml_pass2(const AnchorTable &A, const Epoch<PME> &ep, unsigned mask, double px, double py,
                     double pz, MlPass2 &o, const DistStore *ds = nullptr) {
    double sse = 0.0, g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0, h2 = 0.0, c1s = 0.0;
#pragma unroll 2
    for (int i = 0; i < ep.m_slots; ++i) {
        if (!((mask >> i) & 1u)) continue;
        const double dx = A.x[i] - px, dy = A.y[i] - py, dz = A.z[i] - pz;
        const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
        const double invd = fast_rsqrt(d2);
        const double d = d2 * invd;
        const double r = ep.z[i];
        const double res = r - d;
        const double rid = r * invd;
        if (STORE) {
            ds->invd[i] = invd;
            ds->eps[i] = res;
        }
        const double w = PME ? 1.0 / ep.e[i] : 1.0;
        sse = fma(res, res, sse);
        const double t = res * invd * w;
        g0 = fma(t, dx, g0); g1 = fma(t, dy, g1);
        c1s = fma(1.0 - rid, w, c1s);
        const double c2 = rid * invd * invd * w;
        const double cx = c2 * dx, cy = c2 * dy;
        h0 = fma(cx, dx, h0); h1 = fma(cx, dy, h1); h2 = fma(cy, dy, h2);
    }
    o.sse = sse;
    o.g[0] = g0; o.g[1] = g1;
    o.H[0] = h0 + c1s; o.H[1] = h1; o.H[2] = h2 + c1s;
}

// estimatePosition2D, ML.cpp:48-143.  z stays at its start value.  The damping
// `step` of the reference never takes effect: a rejected step leaves
// newCost == cost, so the while-test fails on the next evaluation (ML.cpp:109-116).
// B-1: the tentative cost is evaluated at z = start z.
template <bool PME, bool STORE = false>
KF_DEV int ml_solve2(const AnchorTable &A, const Epoch<PME> &ep, unsigned mask, double (&p)[3],
                     double &sse_out, unsigned &iters, double *cov /* xx, xy, yy or null */,
                     const DistStore *ds = nullptr, double *sse_start = nullptr) {
    MlPass2 ps;
    ml_pass2<PME, STORE>(A, ep, mask, p[0], p[1], p[2], ps, ds);
    sse_out = ps.sse;
    if (sse_start) *sse_start = ps.sse;
    if (__popc(mask) < 3) return ML_FEW;
    double cost = 1e20, newCost = ps.sse;
    unsigned iter = 0;
    while ((fabs(cost - newCost) / cost > 1e-3) && (iter < 10000u)) {
        iter += 1;
        cost = newCost;
        double s0, s1;
        if (!solve_sym2(ps.H[0], ps.H[1], ps.H[2], ps.g[0], ps.g[1], s0, s1)) {
            iters += iter;
            return ML_SINGULAR;
        }
        const double nx = p[0] - s0, ny = p[1] - s1;
        MlPass2 pt;
        ml_pass2<PME>(A, ep, mask, nx, ny, p[2], pt);
        if (pt.sse > cost) break; // step /= 2; position kept; loop ends
        newCost = pt.sse;
        p[0] = nx; p[1] = ny;
        ps = pt;
    }
    iters += iter;
    sse_out = ps.sse;
    if (cov) {
        double m00 = 0, m01 = 0, m11 = 0;
        for (int i = 0; i < ep.m_slots; ++i) {
            if (!((mask >> i) & 1u)) continue;
            const double dx = p[0] - A.x[i], dy = p[1] - A.y[i], dz = p[2] - A.z[i];
            const double id2 = 1.0 / (dx * dx + dy * dy + dz * dz);
            const double w = id2 / fmax(ep.err(i), ps.sse);
            m00 = fma(w * dx, dx, m00);
            m01 = fma(w * dx, dy, m01);
            m11 = fma(w * dy, dy, m11);
        }
        const double det = m00 * m11 - m01 * m01;
        if (!(det != 0.0)) return ML_SINGULAR;
        const double id = 1.0 / det;
        cov[0] = m11 * id;
        cov[1] = -m01 * id;
        cov[2] = m00 * id;
    }
    return ML_OK;
}

// ---- asynchronous epoch loads (LDGSTS / cp.async): the rangings of the NEXT
// epoch stream from HBM straight into a shared-memory column while the current
// epoch is being processed, so no warp ever waits on DRAM and no registers are
// spent on staging.  ranges: SoA with the filter index fastest; `base` = element
// index of slot 0 for this filter, `N` = element stride between slots.
KF_DEV void cp_async_4(void *dst, const void *src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(src) : "memory");
}
KF_DEV void cp_async_8(void *dst, const void *src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(src) : "memory");
}
KF_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
KF_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

KF_DEV void prefetch_epoch(const Col &raw, int m, const void *ranges, int fmt, int64_t base, int64_t N) {
    for (int i = 0; i < m; ++i) {
        const int64_t idx = base + (int64_t)i * N;
        if (fmt == 0) cp_async_8(&raw[i], reinterpret_cast<const double *>(ranges) + idx);
        else if (fmt == 1) cp_async_4(&raw[i], reinterpret_cast<const int32_t *>(ranges) + idx);
        else // uint16: fetch the aligned 32-bit word that holds the element
            cp_async_4(&raw[i], reinterpret_cast<const void *>(
                                    reinterpret_cast<uintptr_t>(reinterpret_cast<const uint16_t *>(ranges) + idx) & ~(uintptr_t)3));
    }
    cp_async_commit();
}

// raw column -> metres + valid mask: keep rangings[i] > 0 (TOA.cpp:48-57, ML.cpp:478)
template <bool PME>
KF_DEV void convert_epoch(Epoch<PME> &ep, const Col &raw, const void *ranges, int fmt, const double *err,
                          int64_t base, int64_t N) {
    unsigned valid = 0u;
    for (int i = 0; i < ep.m_slots; ++i) {
        double r;
        if (fmt == 0) {
            r = raw[i];
        } else {
            const unsigned w = *reinterpret_cast<const unsigned *>(&raw[i]);
            if (fmt == 1) {
                r = mm_to_m((double)(int)w);
            } else {
                const uintptr_t a = reinterpret_cast<uintptr_t>(reinterpret_cast<const uint16_t *>(ranges) + base + (int64_t)i * N);
                r = mm_to_m((double)((a & 2) ? (w >> 16) : (w & 0xffffu)));
            }
        }
        ep.z[i] = r;
        if (r > 0) valid |= 1u << i;
        if (PME) ep.e[i] = __ldg(err + base + (int64_t)i * N);
    }
    ep.valid = valid;
}

// Packed landing zone: 32-bit words for the mm wire formats (half the shared memory of a
// double column), doubles for the f64 format.  Element i of thread t lives at word/double
// index i * stride + t of the region (conflict-free either way).
struct RawCol {
    void *p;
    int stride;
    KF_DEV unsigned *w(int i) const { return reinterpret_cast<unsigned *>(p) + i * stride; }
    KF_DEV double *d(int i) const { return reinterpret_cast<double *>(p) + i * stride; }
};
KF_DEV RawCol make_raw(double *region, int fmt, int tid, int block) {
    RawCol r;
    r.p = fmt == 0 ? static_cast<void *>(region + tid) : static_cast<void *>(reinterpret_cast<unsigned *>(region) + tid);
    r.stride = block;
    return r;
}

KF_DEV void prefetch_epoch(const RawCol &raw, int m, const void *ranges, int fmt, int64_t base, int64_t N) {
    for (int i = 0; i < m; ++i) {
        const int64_t idx = base + (int64_t)i * N;
        if (fmt == 0) cp_async_8(raw.d(i), reinterpret_cast<const double *>(ranges) + idx);
        else if (fmt == 1) cp_async_4(raw.w(i), reinterpret_cast<const int32_t *>(ranges) + idx);
        else // uint16: fetch the aligned 32-bit word that holds the element
            cp_async_4(raw.w(i), reinterpret_cast<const void *>(
                                     reinterpret_cast<uintptr_t>(reinterpret_cast<const uint16_t *>(ranges) + idx) & ~(uintptr_t)3));
    }
    cp_async_commit();
}

template <bool PME>
KF_DEV void convert_epoch(Epoch<PME> &ep, const RawCol &raw, const void *ranges, int fmt, const double *err,
                          int64_t base, int64_t N) {
    unsigned valid = 0u;
    for (int i = 0; i < ep.m_slots; ++i) {
        double r;
        if (fmt == 0) {
            r = *raw.d(i);
        } else {
            const unsigned w = *raw.w(i);
            if (fmt == 1) {
                r = mm_to_m((double)(int)w);
            } else {
                const uintptr_t a = reinterpret_cast<uintptr_t>(reinterpret_cast<const uint16_t *>(ranges) + base + (int64_t)i * N);
                r = mm_to_m((double)((a & 2) ? (w >> 16) : (w & 0xffffu)));
            }
        }
        ep.z[i] = r;
        if (r > 0) valid |= 1u << i;
        if (PME) ep.e[i] = __ldg(err + base + (int64_t)i * N);
    }
    ep.valid = valid;
}

} // namespace kfpos
