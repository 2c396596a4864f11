/*
 * ko_dense.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Dense primitives standing in for the Armadillo/LAPACK calls on the reference
 * hot path (Armadillo is un-vendored and un-pinned: CMakeLists.txt:29-30,66):
 *   arma::inv    KF.cpp:446,492  TOA.cpp:289,317  TOAIMU.cpp:298,331  ML.cpp:139,252
 *   arma::pinv   KF.cpp:447      TOA.cpp:290      TOAIMU.cpp:299
 *   arma::solve  ML.cpp:100 (default), ML.cpp:210 (solve_opts::equilibrate)
 * Published algorithms restated: inv = LU with partial pivoting (getrf/getri),
 * pinv = SVD with tolerance max(m,n)*sigma_max*eps (gesdd + threshold),
 * solve = LU with partial pivoting (gesv) / with row+column equilibration
 * (gesvx 'E': geequ scaling applied when rowcnd/colcnd < 0.1).
 */
#include <float.h>
#include <math.h>
#include <string.h>

#include "kfpos_oracle.h"

#define NMAX KO_MAX_ROWS

/* LU with partial pivoting in place; piv[] row swaps; returns -1 on an exactly
 * zero (or NaN) pivot, like getrf's info>0. */
static int lu_factor(int n, double *A, int *piv) {
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double v = fabs(A[i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (!(best > 0.0)) return -1; /* zero or NaN pivot */
        if (p != k)
            for (int j = 0; j < n; ++j) {
                double t = A[k * n + j];
                A[k * n + j] = A[p * n + j];
                A[p * n + j] = t;
            }
        double inv_p = 1.0 / A[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            double l = A[i * n + k] * inv_p;
            A[i * n + k] = l;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
        }
    }
    return 0;
}

static void lu_solve(int n, const double *LU, const int *piv, double *b) {
    for (int k = 0; k < n; ++k) { /* all interchanges first (dlaswp), then L */
        int p = piv[k];
        if (p != k) { double t = b[k]; b[k] = b[p]; b[p] = t; }
    }
    for (int k = 0; k < n; ++k)
        for (int i = k + 1; i < n; ++i) b[i] -= LU[i * n + k] * b[k];
    for (int k = n - 1; k >= 0; --k) {
        double s = b[k];
        for (int j = k + 1; j < n; ++j) s -= LU[k * n + j] * b[j];
        b[k] = s / LU[k * n + k];
    }
}

int ko_inv(int n, const double *A, double *Ainv) {
    double LU[NMAX * NMAX];
    int piv[NMAX];
    double col[NMAX];
    if (n == 0) return 0;
    memcpy(LU, A, sizeof(double) * n * n);
    if (lu_factor(n, LU, piv) != 0) return -1;
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) col[i] = (i == j) ? 1.0 : 0.0;
        lu_solve(n, LU, piv, col);
        for (int i = 0; i < n; ++i) Ainv[i * n + j] = col[i];
    }
    return 0;
}

int ko_solve(int n, const double *A, const double *b, double *x, int equilibrate) {
    double LU[NMAX * NMAX];
    int piv[NMAX];
    double r[NMAX], c[NMAX];
    int row_scaled = 0, col_scaled = 0;
    memcpy(LU, A, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) x[i] = b[i];
    if (equilibrate) {
        /* dgeequ: r_i = 1/max_j|a_ij| ; c_j = 1/max_i(r_i|a_ij|) ; dlaqge applies
         * them when the condition ratios are below 0.1. */
        double rmin = DBL_MAX, rmax = 0, cmin = DBL_MAX, cmax = 0;
        for (int i = 0; i < n; ++i) {
            double m = 0;
            for (int j = 0; j < n; ++j) m = fmax(m, fabs(LU[i * n + j]));
            if (!(m > 0)) return -1;
            r[i] = 1.0 / m;
            rmin = fmin(rmin, m);
            rmax = fmax(rmax, m);
        }
        for (int j = 0; j < n; ++j) {
            double m = 0;
            for (int i = 0; i < n; ++i) m = fmax(m, r[i] * fabs(LU[i * n + j]));
            if (!(m > 0)) return -1;
            c[j] = 1.0 / m;
            cmin = fmin(cmin, m);
            cmax = fmax(cmax, m);
        }
        row_scaled = (rmin / rmax) < 0.1;
        col_scaled = (cmin / cmax) < 0.1;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                if (row_scaled) LU[i * n + j] *= r[i];
                if (col_scaled) LU[i * n + j] *= c[j];
            }
        if (row_scaled)
            for (int i = 0; i < n; ++i) x[i] *= r[i];
    }
    if (lu_factor(n, LU, piv) != 0) return -1;
    lu_solve(n, LU, piv, x);
    if (col_scaled)
        for (int i = 0; i < n; ++i) x[i] *= c[i];
    return 0;
}

/* One-sided (Hestenes) Jacobi SVD of a square matrix: A = U diag(S) V^T.
 * Row-major n x n. Columns of U for zero singular values are left zero. */
void ko_svd_jacobi(int n, const double *A, double *U, double *S, double *V) {
    double W[NMAX * NMAX]; /* working copy, columns get orthogonalised */
    memcpy(W, A, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < n; ++i) {
                    double wp = W[i * n + p], wq = W[i * n + q];
                    alpha += wp * wp;
                    beta += wq * wq;
                    gamma += wp * wq;
                }
                if (gamma == 0.0 || fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) continue;
                rotated = 1;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < n; ++i) {
                    double wp = W[i * n + p], wq = W[i * n + q];
                    W[i * n + p] = cs * wp - sn * wq;
                    W[i * n + q] = sn * wp + cs * wq;
                    double vp = V[i * n + p], vq = V[i * n + q];
                    V[i * n + p] = cs * vp - sn * vq;
                    V[i * n + q] = sn * vp + cs * vq;
                }
            }
        if (!rotated) break;
    }
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += W[i * n + j] * W[i * n + j];
        s = sqrt(s);
        S[j] = s;
        for (int i = 0; i < n; ++i) U[i * n + j] = (s > 0) ? W[i * n + j] / s : 0.0;
    }
}

/* arma::pinv default tolerance: max(rows,cols) * max_sv * eps */
void ko_pinv(int n, const double *A, double *Apinv) {
    double U[NMAX * NMAX], V[NMAX * NMAX], S[NMAX];
    ko_svd_jacobi(n, A, U, S, V);
    double smax = 0;
    for (int j = 0; j < n; ++j) smax = fmax(smax, S[j]);
    double tol = (double)n * smax * DBL_EPSILON;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0;
            for (int k = 0; k < n; ++k)
                if (S[k] > tol) acc += V[i * n + k] * U[j * n + k] / S[k];
            Apinv[i * n + j] = acc;
        }
}
