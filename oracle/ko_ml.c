/*
 * ko_ml.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * As-written restatement of MLLocation (reference: src/kfpos/algorithms/
 * MLLocation.cpp).  Restatement decisions for reference defects follow
 * SURVEY.md App. B and are repeated at each site.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "kfpos_oracle.h"

/* MLLocation::distanceToBeacons, ML.cpp:24-37 (expression order kept) */
void ko_dist(const ko_meas *m, int n, const double p[3], double *d) {
    for (int i = 0; i < n; ++i)
        d[i] = sqrt((m[i].bx - p[0]) * (m[i].bx - p[0]) + (m[i].by - p[1]) * (m[i].by - p[1]) +
                    (m[i].bz - p[2]) * (m[i].bz - p[2]));
}

/* MLLocation::estimationError, ML.cpp:263-278 */
double ko_sse(const ko_meas *m, int n, const double p[3]) {
    if (n == 0) return -1;
    double d[KO_MAX_ANCHORS], e = 0.0;
    ko_dist(m, n, p, d);
    for (int i = 0; i < n; ++i) e += (d[i] - m[i].r) * (d[i] - m[i].r);
    return e;
}

/* covariance tail shared by 2-D and 3-D: ML.cpp:118-140 / 229-254 */
static int ml_cov(const ko_meas *m, int n, const double p[3], int d, double *cov) {
    double dist[KO_MAX_ANCHORS], JtWJ[9] = {0};
    ko_dist(m, n, p, dist);
    double sse = ko_sse(m, n, p);
    for (int i = 0; i < n; ++i) {
        double J[3] = {(p[0] - m[i].bx) / dist[i], (p[1] - m[i].by) / dist[i],
                       (p[2] - m[i].bz) / dist[i]};
        double w = 1.0 / fmax(m[i].e, sse); /* inv(diagmat(realObservationError)) */
        for (int a = 0; a < d; ++a)
            for (int b = 0; b < d; ++b) JtWJ[a * d + b] += J[a] * w * J[b];
    }
    return ko_inv(d, JtWJ, cov);
}

/* MLLocation::estimatePosition2D, ML.cpp:48-143.
 * B-1: tentativePos.z is uninitialised in the reference (ML.cpp:64,102-106);
 * restated as z = previousEstimation.z.  A zero-initialised Vector3 (what the
 * shim-compiled reference sees) corresponds to b1_zero_z = 1. */
int ko_ml2d(const ko_meas *m, int n, const double start[3], int b1_zero_z, double pos[3],
            double cov[4], int *iters) {
    pos[0] = start[0]; pos[1] = start[1]; pos[2] = start[2];
    if (iters) *iters = 0;
    if (n < 3) return 1;
    double cost = 1e20, newCost, step = 1;
    double tent[3] = {0, 0, b1_zero_z ? 0.0 : start[2]};
    newCost = ko_sse(m, n, pos);
    int iter = 0, rc = 0;
    while ((fabs(cost - newCost) / cost > 1e-3) && (iter < 10000)) {
        iter += 1;
        cost = newCost;
        double d[KO_MAX_ANCHORS], g[2] = {0, 0}, H[4] = {0, 0, 0, 0};
        ko_dist(m, n, pos, d);
        for (int i = 0; i < n; ++i) {
            const ko_meas *r = &m[i];
            g[0] += (r->r - d[i]) * (r->bx - pos[0]) / (d[i] * r->e);
            g[1] += (r->r - d[i]) * (r->by - pos[1]) / (d[i] * r->e);
            double d3 = d[i] * d[i] * d[i];
            H[0] += (1 - r->r / d[i] + r->r * (r->bx - pos[0]) * (r->bx - pos[0]) / d3) / r->e;
            H[3] += (1 - r->r / d[i] + r->r * (r->by - pos[1]) * (r->by - pos[1]) / d3) / r->e;
            double dxy = r->r * (r->bx - pos[0]) * (r->by - pos[1]) / (d3 * r->e);
            H[1] += dxy;
            H[2] += dxy;
        }
        double rhs[2] = {H[0] * pos[0] + H[1] * pos[1] - g[0] * step,
                         H[2] * pos[0] + H[3] * pos[1] - g[1] * step};
        double np[2];
        if (ko_solve(2, H, rhs, np, 0) != 0) { rc = -1; break; }
        tent[0] = np[0];
        tent[1] = np[1];
        double tc = ko_sse(m, n, tent);
        if (tc > cost) {
            step /= 2;
        } else {
            newCost = tc;
            step = 1;
            pos[0] = np[0];
            pos[1] = np[1];
        }
    }
    if (iters) *iters = iter;
    if (rc != 0) return rc;
    return ml_cov(m, n, pos, 2, cov) != 0 ? -1 : 0;
}

/* MLLocation::estimatePosition, ML.cpp:153-257 */
int ko_ml3d(const ko_meas *m, int n, const double start[3], double pos[3], double cov[9],
            int *iters) {
    pos[0] = start[0]; pos[1] = start[1]; pos[2] = start[2];
    if (iters) *iters = 0;
    if (n < 4) return 1;
    double cost = 1e20, newCost = 1;
    int iter = 0, rc = 0;
    while ((fabs(cost - newCost) / cost > 1e-3) && (iter < 10000)) {
        iter += 1;
        cost = newCost;
        double d[KO_MAX_ANCHORS], g[3] = {0, 0, 0}, H[9] = {0};
        ko_dist(m, n, pos, d);
        for (int i = 0; i < n; ++i) {
            const ko_meas *r = &m[i];
            double dx = r->bx - pos[0], dy = r->by - pos[1], dz = r->bz - pos[2];
            g[0] += (r->r - d[i]) * dx / (d[i] * r->e);
            g[1] += (r->r - d[i]) * dy / (d[i] * r->e);
            g[2] += (r->r - d[i]) * dz / (d[i] * r->e);
            double d3 = d[i] * d[i] * d[i];
            H[0] += (1 - r->r / d[i] + r->r * dx * dx / d3) / r->e;
            H[4] += (1 - r->r / d[i] + r->r * dy * dy / d3) / r->e;
            H[8] += (1 - r->r / d[i] + r->r * dz * dz / d3) / r->e;
            double dxy = r->r * dx * dy / (d3 * r->e);
            double dxz = r->r * dx * dz / (d3 * r->e);
            double dyz = r->r * dy * dz / (d3 * r->e);
            H[1] += dxy; H[2] += dxz; H[5] += dyz;
            H[3] += dxy; H[6] += dxz; H[7] += dyz;
        }
        double rhs[3], np[3];
        for (int a = 0; a < 3; ++a)
            rhs[a] = H[a * 3 + 0] * pos[0] + H[a * 3 + 1] * pos[1] + H[a * 3 + 2] * pos[2] - g[a];
        if (ko_solve(3, H, rhs, np, 1) != 0) { rc = -1; break; }
        pos[0] = np[0]; pos[1] = np[1]; pos[2] = np[2];
        ko_dist(m, n, pos, d);
        newCost = 0.0;
        for (int i = 0; i < n; ++i) newCost += (m[i].r - d[i]) * (m[i].r - d[i]) / m[i].e;
    }
    if (iters) *iters = iter;
    if (rc != 0) return rc;
    return ml_cov(m, n, pos, 3, cov) != 0 ? -1 : 0;
}

/* MLLocation::bestRangingsByDistance, ML.cpp:284-300.
 * B-11: std::sort is unstable; ties broken by lower original index first. */
void ko_best_rangings(const ko_meas *m, int n, const double p[3], int *order) {
    double d[KO_MAX_ANCHORS], q[KO_MAX_ANCHORS];
    ko_dist(m, n, p, d);
    for (int i = 0; i < n; ++i) {
        q[i] = (d[i] - m[i].r) * (d[i] - m[i].r);
        order[i] = i;
    }
    for (int i = 1; i < n; ++i) { /* stable insertion sort on (q, index) */
        int oi = order[i], j = i - 1;
        while (j >= 0 && q[order[j]] > q[oi]) {
            order[j + 1] = order[j];
            --j;
        }
        order[j + 1] = oi;
    }
}

static int ml_any(const ko_meas *m, int n, const double start[3], int use2d, int b1_zero_z,
                  double pos[3], double *cov, int *iters) {
    return use2d ? ko_ml2d(m, n, start, b1_zero_z, pos, cov, iters)
                 : ko_ml3d(m, n, start, pos, cov, iters);
}

/* MLLocation::estimatePositionIgnoreN, ML.cpp:307-347 */
int ko_ml_ignore_n(const ko_meas *m, int n, const double start[3], int use2d, int n_ignore,
                   int b1_zero_z, double pos[3], double *cov, int *iters, int *order,
                   int *n_dropped) {
    double p0[3], cov0[9];
    int it0 = 0, it1 = 0;
    int min_r = use2d ? 3 : 4;
    if (ml_any(m, n, start, use2d, b1_zero_z, p0, cov0, &it0) < 0) {
        /* the first solve throws out of estimatePositionIgnoreN (ML.cpp:315-320): nothing is selected */
        for (int i = 0; i < n; ++i) order[i] = i;
        pos[0] = p0[0]; pos[1] = p0[1]; pos[2] = p0[2];
        if (n_dropped) *n_dropped = -1;
        if (iters) *iters = it0;
        return -1;
    }
    ko_best_rangings(m, n, p0, order);
    ko_meas sorted[KO_MAX_ANCHORS];
    for (int i = 0; i < n; ++i) sorted[i] = m[order[i]];
    int drop = n - min_r < n_ignore ? n - min_r : n_ignore; /* std::min, may be <= 0 */
    if (drop < 0) drop = 0;
    if (n_dropped) *n_dropped = drop;
    int rc = ml_any(sorted, n - drop, start, use2d, b1_zero_z, pos, cov, &it1);
    if (iters) *iters = it0 + it1;
    return rc;
}

/* std::prev_permutation on a bool vector */
static int prev_perm(unsigned char *v, int n) {
    int i = n - 1;
    while (i > 0 && v[i - 1] <= v[i]) --i;
    if (i <= 0) return 0;
    int j = n - 1;
    while (v[j] >= v[i - 1]) --j;
    unsigned char t = v[i - 1]; v[i - 1] = v[j]; v[j] = t;
    for (int a = i, b = n - 1; a < b; ++a, --b) { t = v[a]; v[a] = v[b]; v[b] = t; }
    return 1;
}

/* MLLocation::estimatePositionBestGroup, ML.cpp:351-414.
 * B-3: subset = measurements whose mask bit is true (the reference's
 *      erase-while-indexing loop is only correct for n = k+1).
 * B-4: 2-D criterion = cov(0,0)+cov(1,1); 3-D = trace, or cov(2,2) when
 *      best_mode = 1 (config_pos.xml:18-20 documentation). */
int ko_ml_best_group(const ko_meas *m, int n, const double start[3], int use2d, int best_mode,
                     int b1_zero_z, double pos[3], double *cov, int *iters, int *best_index,
                     uint32_t *best_mask, int *n_groups) {
    int k = use2d ? 3 : 4, d = use2d ? 2 : 3;
    int it_total = 0, it = 0;
    int rc = ml_any(m, n, start, use2d, b1_zero_z, pos, cov, &it);
    it_total += it;
    if (best_index) *best_index = -1;
    if (best_mask) *best_mask = 0;
    if (n_groups) *n_groups = 0;
    if (n < k || rc < 0) { /* a failed solve throws out of estimatePositionBestGroup (ML.cpp:353-358) */
        if (iters) *iters = it_total;
        return rc;
    }
    double pos_all[3] = {pos[0], pos[1], pos[2]}, cov_all[9];
    memcpy(cov_all, cov, sizeof(double) * d * d);
    unsigned char v[KO_MAX_ANCHORS];
    for (int i = 0; i < n; ++i) v[i] = i < k;
    double minErr = 0;
    int minIdx = -1, idx = 0;
    do {
        ko_meas sub[KO_MAX_ANCHORS];
        int ns = 0;
        uint32_t mask = 0;
        for (int i = 0; i < n; ++i)
            if (v[i]) {
                sub[ns++] = m[i];
                mask |= 1u << i;
            }
        double gp[3], gc[9] = {0};
        int grc = ml_any(sub, ns, start, use2d, b1_zero_z, gp, gc, &it);
        it_total += it;
        double cur;
        if (use2d) cur = gc[0] + gc[3];
        else if (best_mode == 1) cur = gc[8];
        else cur = gc[0] + gc[4] + gc[8];
        if (grc != 0) {
            /* arma::solve / inv throws inside the subset loop (ML.cpp:384-389): the reference never
             * reaches the selection.  Restated as: the epoch reports SINGULAR, keeps the all-ranging
             * estimate and selects no group -- whichever subset failed, first or not. */
            pos[0] = pos_all[0]; pos[1] = pos_all[1]; pos[2] = pos_all[2];
            memcpy(cov, cov_all, sizeof(double) * d * d);
            if (best_mask) *best_mask = 0;
            if (n_groups) *n_groups = idx + 1;
            if (iters) *iters = it_total;
            return -1;
        }
        if (minIdx == -1 || cur <= minErr) { /* ML.cpp:402-410: first seeds, then <= */
            minIdx = idx;
            minErr = cur;
            pos[0] = gp[0]; pos[1] = gp[1]; pos[2] = gp[2];
            memcpy(cov, gc, sizeof(double) * d * d);
            if (best_mask) *best_mask = mask;
            rc = grc;
        }
        ++idx;
    } while (prev_perm(v, n));
    if (best_index) *best_index = minIdx;
    if (n_groups) *n_groups = idx;
    if (iters) *iters = it_total;
    return rc;
}

/* MLLocation::newTOAMeasurement (ML.cpp:472-486) + getPose (ML.cpp:421-469) */
int ko_ml_epoch(int n_slots, const double *ranges, const double *anchors, const double *errs,
                const double start[3], int use2d, int variant, int n_ignore, int best_mode,
                int b1_zero_z, double pos[3], double cov[9], int *iters, int32_t *sel) {
    ko_meas m[KO_MAX_ANCHORS];
    int slot_of[KO_MAX_ANCHORS];
    int n = 0;
    for (int i = 0; i < n_slots; ++i)
        if (ranges[i] > 0) {
            m[n].r = ranges[i];
            m[n].e = errs[i];
            m[n].bx = anchors[3 * i];
            m[n].by = anchors[3 * i + 1];
            m[n].bz = anchors[3 * i + 2];
            m[n].slot = i;
            slot_of[n] = i;
            ++n;
        }
    for (int i = 0; i < 9; ++i) cov[i] = 0;
    uint32_t used = 0;
    for (int i = 0; i < n; ++i) used |= 1u << slot_of[i];
    int rc, idx = -1;
    if (variant == 0) {
        rc = use2d ? ko_ml2d(m, n, start, b1_zero_z, pos, cov, iters)
                   : ko_ml3d(m, n, start, pos, cov, iters);
    } else if (variant == 1) {
        int order[KO_MAX_ANCHORS], drop = 0;
        rc = ko_ml_ignore_n(m, n, start, use2d, n_ignore, b1_zero_z, pos, cov, iters, order, &drop);
        for (int i = n - (drop > 0 ? drop : 0); i < n; ++i) used &= ~(1u << slot_of[order[i]]);
        idx = drop;
    } else {
        uint32_t mask = 0;
        int ng = 0;
        rc = ko_ml_best_group(m, n, start, use2d, best_mode, b1_zero_z, pos, cov, iters, &idx,
                              &mask, &ng);
        if (idx >= 0) {
            used = 0;
            for (int i = 0; i < n; ++i)
                if (mask & (1u << i)) used |= 1u << slot_of[i];
        }
    }
    if (sel) {
        sel[0] = (int32_t)used; /* bit per anchor SLOT used in the final solve */
        sel[1] = idx;           /* variant 1: #dropped; variant 2: subset index */
    }
    return rc;
}
