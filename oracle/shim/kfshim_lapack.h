// LAPACK back end of the mini-Armadillo shim (TEST INFRASTRUCTURE): the routines
// Armadillo itself calls (getrf/getri, gesv, gesvx, gesdd), taken from the
// OpenBLAS that ships inside the scipy wheel (symbols scipy_dgetrf_, ...), loaded
// with dlopen from the path in $KFSHIM_LAPACK or given to kfshim::lapack_open().
#pragma once
#include <dlfcn.h>

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace kfshim {

typedef void (*dgetrf_t)(const int *, const int *, double *, const int *, int *, int *);
typedef void (*dgetri_t)(const int *, double *, const int *, const int *, double *, const int *, int *);
typedef void (*dgesv_t)(const int *, const int *, double *, const int *, int *, double *, const int *, int *);
typedef void (*dgesvx_t)(const char *, const char *, const int *, const int *, double *, const int *, double *,
                         const int *, int *, char *, double *, double *, double *, const int *, double *,
                         const int *, double *, double *, double *, double *, int *, int *);
typedef void (*dgesdd_t)(const char *, const int *, const int *, double *, const int *, double *, double *,
                         const int *, double *, const int *, double *, const int *, int *, int *);

struct Lapack {
    void *h;
    dgetrf_t getrf;
    dgetri_t getri;
    dgesv_t gesv;
    dgesvx_t gesvx;
    dgesdd_t gesdd;
};

inline Lapack &lapack() {
    static Lapack L = {0, 0, 0, 0, 0, 0};
    return L;
}

inline bool lapack_open(const char *path) {
    Lapack &L = lapack();
    if (L.h) return true;
    if (!path) path = getenv("KFSHIM_LAPACK");
    if (!path) return false;
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return false;
    L.getrf = (dgetrf_t)dlsym(h, "scipy_dgetrf_");
    L.getri = (dgetri_t)dlsym(h, "scipy_dgetri_");
    L.gesv = (dgesv_t)dlsym(h, "scipy_dgesv_");
    L.gesvx = (dgesvx_t)dlsym(h, "scipy_dgesvx_");
    L.gesdd = (dgesdd_t)dlsym(h, "scipy_dgesdd_");
    if (!L.getrf || !L.getri || !L.gesv || !L.gesvx || !L.gesdd) return false;
    L.h = h;
    return true;
}

inline void need() {
    if (!lapack().h && !lapack_open(0)) throw std::runtime_error("kfshim: LAPACK not loaded (set KFSHIM_LAPACK)");
}

// arma::inv -> getrf + getri (in place, column-major)
inline bool lapack_inv(int n, double *a) {
    need();
    std::vector<int> ipiv(n);
    int info = 0;
    lapack().getrf(&n, &n, a, &n, ipiv.data(), &info);
    if (info != 0) return false;
    int lwork = 64 * n;
    std::vector<double> work(lwork);
    lapack().getri(&n, a, &n, ipiv.data(), work.data(), &lwork, &info);
    return info == 0;
}

// arma::solve: gesv, or gesvx with FACT='E' when solve_opts::equilibrate
inline bool lapack_solve(int n, int nrhs, const double *a, double *bx, bool equilibrate) {
    need();
    std::vector<double> A(a, a + (size_t)n * n);
    std::vector<int> ipiv(n);
    int info = 0;
    if (!equilibrate) {
        lapack().gesv(&n, &nrhs, A.data(), &n, ipiv.data(), bx, &n, &info);
        return info == 0;
    }
    std::vector<double> AF((size_t)n * n), R(n), C(n), B(bx, bx + (size_t)n * nrhs), X((size_t)n * nrhs);
    std::vector<double> ferr(nrhs), berr(nrhs), work(4 * n);
    std::vector<int> iwork(n);
    char equed = 'N';
    double rcond = 0;
    lapack().gesvx("E", "N", &n, &nrhs, A.data(), &n, AF.data(), &n, ipiv.data(), &equed, R.data(), C.data(),
                   B.data(), &n, X.data(), &n, &rcond, ferr.data(), berr.data(), work.data(), iwork.data(), &info);
    if (info != 0 && info != n + 1) return false; // n+1: ill-conditioned but solved
    std::copy(X.begin(), X.end(), bx);
    return true;
}

// arma::pinv -> gesdd, tolerance max(m,n) * max_sv * eps
inline bool lapack_pinv(int m, int n, const double *a, double *out) {
    need();
    std::vector<double> A(a, a + (size_t)m * n);
    const int k = std::min(m, n);
    std::vector<double> S(k), U((size_t)m * k), VT((size_t)k * n);
    int lwork = -1, info = 0;
    double wq = 0;
    std::vector<int> iwork(8 * k);
    lapack().gesdd("S", &m, &n, A.data(), &m, S.data(), U.data(), &m, VT.data(), &k, &wq, &lwork, iwork.data(), &info);
    lwork = (int)wq + 1;
    std::vector<double> work(lwork);
    lapack().gesdd("S", &m, &n, A.data(), &m, S.data(), U.data(), &m, VT.data(), &k, work.data(), &lwork,
                   iwork.data(), &info);
    if (info != 0) return false;
    const double tol = (double)std::max(m, n) * (k ? S[0] : 0.0) * DBL_EPSILON;
    // out (n x m) = V diag(1/s) U^T
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) {
            double acc = 0;
            for (int q = 0; q < k; ++q)
                if (S[q] > tol) acc += VT[(size_t)i * k + q] * U[(size_t)q * m + j] / S[q];
            out[(size_t)j * n + i] = acc;
        }
    return true;
}

} // namespace kfshim
