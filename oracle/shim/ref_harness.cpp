// ref_harness.cpp -- C entry points around the UNMODIFIED reference classes
// (TEST INFRASTRUCTURE).  Built by `make -C oracle ref` into oracle/_ref/libkfref.so
// from the reference's own .cpp files where they lie under /root/reference plus the
// shim headers of this directory.  Used only to pin the CPU oracle (tests/) and as
// the "reference" CPU baseline of bench.py; the product never loads it.
#define private public
#define protected public
#include "KalmanFilter.h"
#include "KalmanFilterTOA.h"
#include "KalmanFilterTOAIMU.h"
#include "MLLocation.h"
#undef private
#undef protected

#include <cstring>

namespace {

Vector3 make_v3(double x, double y, double z) {
    Vector3 v = {};
    v.x = x; v.y = y; v.z = z;
    return v;
}

void fill_inputs(int n, const double *ranges, const double *anchors, const double *errs,
                 std::vector<double> &r, std::vector<Beacon> &b, std::vector<double> &e) {
    // what PosGenerator::calculateTagLocationWithRangings builds (PG.cpp:476-496)
    for (int i = 0; i < n; ++i) {
        r.push_back(ranges[i]);
        e.push_back(errs[i]);
        Beacon be;
        be.id = i;
        be.index = i;
        be.position = make_v3(anchors[3 * i], anchors[3 * i + 1], anchors[3 * i + 2]);
        b.push_back(be);
    }
}

void advance(long long dt_ns) { kfshim::fake_clock::ticks() += dt_ns; }

template <typename F>
int guarded(F f) {
    try {
        f();
        return 0;
    } catch (const std::runtime_error &) {
        return 1;
    } catch (const std::logic_error &) {
        return 2;
    } catch (...) {
        return 3;
    }
}

void copy_mat(const arma::mat &m, int n, double *out) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            out[i * n + j] = ((std::size_t)i < m.n_rows && (std::size_t)j < m.n_cols) ? m.at(i, j) : 0.0;
}

} // namespace

extern "C" {

int ref_init(const char *lapack_path) { return kfshim::lapack_open(lapack_path) ? 0 : -1; }

void ref_set_param(const char *name, const char *value) { kfshim::params()[name] = value; }

// ------------------------------------------------------------------------- ML
void *ref_ml_create(int use2d, int variant, int n_ignore, double sx, double sy, double sz) {
    return new MLLocation(use2d != 0, variant, n_ignore, make_v3(sx, sy, sz));
}
void ref_ml_destroy(void *h) { delete (MLLocation *)h; }

// mode 0: newTOAMeasurement + getPose (ML.cpp:421-486); mode 1: the estimator selected by
// (use2d, variant) called directly, returning its own covarianceMatrix (d x d).
int ref_ml_solve(void *h, int mode, int n, const double *ranges, const double *anchors, const double *errs,
                 double *pos, double *cov9, int *cov_dim) {
    MLLocation *ml = (MLLocation *)h;
    return guarded([&] {
        std::vector<double> r, e;
        std::vector<Beacon> b;
        fill_inputs(n, ranges, anchors, errs, r, b, e);
        ml->newTOAMeasurement(r, b, e, 0.0);
        Vector3 out = {};
        if (mode == 0) {
            ml->getPose(out);
        } else {
            const std::vector<RangingMeasurement> &m = ml->_lastRangingMeasurements;
            if (ml->_variant == ML_VARIANT_NORMAL)
                out = ml->_use2d ? ml->estimatePosition2D(m, ml->_previousEstimation)
                                 : ml->estimatePosition(m, ml->_previousEstimation);
            else if (ml->_variant == ML_VARIANT_IGNORE_N)
                out = ml->estimatePositionIgnoreN(m, ml->_previousEstimation, ml->_numRangingsToIgnore);
            else
                out = ml->estimatePositionBestGroup(m, ml->_previousEstimation);
        }
        pos[0] = out.x; pos[1] = out.y; pos[2] = out.z;
        int d = (int)out.covarianceMatrix.n_rows;
        if (d > 3) d = 3;
        *cov_dim = d;
        for (int i = 0; i < 9; ++i) cov9[i] = 0.0;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) cov9[i * d + j] = out.covarianceMatrix.at(i, j);
    });
}

// ------------------------------------------------------------------------- T6
void *ref_t6_create(double accel_noise, int ignore_worst, double thr, double x, double y, double z) {
    return new KalmanFilterTOA(accel_noise, ignore_worst != 0, thr, make_v3(x, y, z));
}
void ref_t6_destroy(void *h) { delete (KalmanFilterTOA *)h; }

int ref_t6_toa(void *h, long long dt_ns, int n, const double *ranges, const double *anchors, const double *errs) {
    KalmanFilterTOA *f = (KalmanFilterTOA *)h;
    advance(dt_ns);
    return guarded([&] {
        std::vector<double> r, e;
        std::vector<Beacon> b;
        fill_inputs(n, ranges, anchors, errs, r, b, e);
        f->newTOAMeasurement(r, b, e, 0.0);
    });
}
void ref_t6_get(void *h, double *pos, double *P36) {
    KalmanFilterTOA *f = (KalmanFilterTOA *)h;
    pos[0] = f->mPosition.x; pos[1] = f->mPosition.y; pos[2] = f->mPosition.z;
    copy_mat(f->estimationCovariance, 6, P36);
}
int ref_t6_get_pose(void *h, long long dt_ns, double *pos, double *cov36) {
    KalmanFilterTOA *f = (KalmanFilterTOA *)h;
    advance(dt_ns);
    Vector3 p = {};
    bool ok = false;
    int rc = guarded([&] { ok = f->getPose(p); });
    kfshim::fake_clock::ticks() -= dt_ns; // the poll must not move the filter's clock
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    copy_mat(p.covarianceMatrix, 6, cov36);
    return rc ? rc : (ok ? 0 : 4);
}

// ------------------------------------------------------------------------- K8
// The five XML strings must have been registered with ref_set_param under the
// names "kfpos_pos", "kfpos_px4", "kfpos_tag", "kfpos_imu", "kfpos_mag".
void *ref_k8_create(double accel_noise, double init_angle, double jolt, double x, double y, double z) {
    KalmanFilter *f = new KalmanFilter(accel_noise, init_angle, jolt, "kfpos_pos", "kfpos_px4", "kfpos_tag",
                                       "kfpos_imu", "kfpos_mag", make_v3(x, y, z));
    bool ok = false;
    guarded([&] { ok = f->init(); });
    if (!ok) {
        delete f;
        return 0;
    }
    return f;
}
// the constructor WITHOUT initialPosition (KF.cpp:6-32): what PosGenerator builds when the launch file
// leaves useStartPosition at its default 0 (PG.cpp:519-528); the first epoch with rangings initialises
// the filter through MLLocation (KF.cpp:244-285)
void *ref_k8_create_nofix(double accel_noise, double init_angle, double jolt) {
    KalmanFilter *f = new KalmanFilter(accel_noise, init_angle, jolt, "kfpos_pos", "kfpos_px4", "kfpos_tag",
                                       "kfpos_imu", "kfpos_mag");
    bool ok = false;
    guarded([&] { ok = f->init(); });
    if (!ok) {
        delete f;
        return 0;
    }
    return f;
}
double ref_k8_tag_z(void *h) { return ((KalmanFilter *)h)->mUWBtagZ; }
void ref_k8_destroy(void *h) { delete (KalmanFilter *)h; }

int ref_k8_toa(void *h, long long dt_ns, int n, const double *ranges, const double *anchors, const double *errs) {
    KalmanFilter *f = (KalmanFilter *)h;
    advance(dt_ns);
    return guarded([&] {
        std::vector<double> r, e;
        std::vector<Beacon> b;
        fill_inputs(n, ranges, anchors, errs, r, b, e);
        f->newTOAMeasurement(r, b, e, 0.0);
    });
}
int ref_k8_px4(void *h, long long dt_ns, double ix, double iy, double irz, double itime_us, int quality) {
    KalmanFilter *f = (KalmanFilter *)h;
    advance(dt_ns);
    return guarded([&] { f->newPX4FlowMeasurement(ix, iy, irz, itime_us, quality); });
}
int ref_k8_imu(void *h, long long dt_ns, const double *angvel, const double *cov_av, const double *acc,
               const double *cov_acc) {
    KalmanFilter *f = (KalmanFilter *)h;
    advance(dt_ns);
    return guarded([&] {
        VectorDim3 w = {angvel[0], angvel[1], angvel[2]}, a = {acc[0], acc[1], acc[2]};
        double c1[9], c2[9];
        memcpy(c1, cov_av, sizeof c1);
        memcpy(c2, cov_acc, sizeof c2);
        f->newIMUMeasurement(w, c1, a, c2);
    });
}
int ref_k8_mag(void *h, long long dt_ns, const double *mag) {
    KalmanFilter *f = (KalmanFilter *)h;
    advance(dt_ns);
    return guarded([&] {
        VectorDim3 m = {mag[0], mag[1], mag[2]};
        double c[9] = {0};
        f->newMAGMeasurement(m, c);
    });
}
int ref_k8_compass(void *h, long long dt_ns, double compass) {
    KalmanFilter *f = (KalmanFilter *)h;
    advance(dt_ns);
    return guarded([&] { f->newCompassMeasurement(compass); });
}
void ref_k8_get(void *h, double *x8, double *P64) {
    KalmanFilter *f = (KalmanFilter *)h;
    x8[0] = f->mPosition.x; x8[1] = f->mPosition.y;
    x8[2] = f->mVelocity.x; x8[3] = f->mVelocity.y;
    x8[4] = f->mAcceleration.x; x8[5] = f->mAcceleration.y;
    x8[6] = f->mAngle; x8[7] = f->mAngularSpeed;
    copy_mat(f->mEstimationCovariance, 8, P64);
}

// ------------------------------------------------------------------------- T9
void *ref_t9_create(double accel_noise, double jolt, double x, double y, double z) {
    return new KalmanFilterTOAIMU(accel_noise, jolt, make_v3(x, y, z));
}
// the constructor without initialPosition (TOAIMU.cpp:6-24), see ref_k8_create_nofix
void *ref_t9_create_nofix(double accel_noise, double jolt) { return new KalmanFilterTOAIMU(accel_noise, jolt); }
void ref_t9_destroy(void *h) { delete (KalmanFilterTOAIMU *)h; }
int ref_t9_toa(void *h, long long dt_ns, int n, const double *ranges, const double *anchors, const double *errs) {
    KalmanFilterTOAIMU *f = (KalmanFilterTOAIMU *)h;
    advance(dt_ns);
    return guarded([&] {
        std::vector<double> r, e;
        std::vector<Beacon> b;
        fill_inputs(n, ranges, anchors, errs, r, b, e);
        f->newTOAMeasurement(r, b, e, 0.0);
    });
}
int ref_t9_imu(void *h, long long dt_ns, const double *acc, const double *cov_acc) {
    KalmanFilterTOAIMU *f = (KalmanFilterTOAIMU *)h;
    advance(dt_ns);
    return guarded([&] {
        VectorDim3 w = {0, 0, 0}, a = {acc[0], acc[1], acc[2]};
        double c1[9] = {0}, c2[9];
        memcpy(c2, cov_acc, sizeof c2);
        f->newIMUMeasurement(w, c1, a, c2);
    });
}
void ref_t9_get(void *h, double *x9, double *P81) {
    KalmanFilterTOAIMU *f = (KalmanFilterTOAIMU *)h;
    x9[0] = f->mPosition.x; x9[1] = f->mPosition.y; x9[2] = f->mPosition.z;
    x9[3] = f->mVelocity.x; x9[4] = f->mVelocity.y; x9[5] = f->mVelocity.z;
    x9[6] = f->mAcceleration.x; x9[7] = f->mAcceleration.y; x9[8] = f->mAcceleration.z;
    copy_mat(f->mEstimationCovariance, 9, P81);
}

// ------------------------------------------------- pose message (getPose + publisher)
// getPose of the given filter `dt` after its last update, then the fields exactly as
// PosGenerator::publishPositionReport reads them (Posgenerator.cpp:385-470): position,
// orientation quaternion, linear and angular speed, and covariance[i] = covarianceMatrix(i)
// for i < 36 (Armadillo's column-major linear index, also for T9's 9x9 matrix).
// kind: 1 = KalmanFilterTOA, 2 = KalmanFilter, 3 = KalmanFilterTOAIMU.
int ref_get_pose_msg(void *h, int kind, long long dt_ns, double *pose13, double *cov36) {
    advance(dt_ns);
    Vector3 p = {};
    bool ok = false;
    int rc = guarded([&] {
        if (kind == 1) ok = ((KalmanFilterTOA *)h)->getPose(p);
        else if (kind == 2) ok = ((KalmanFilter *)h)->getPose(p);
        else ok = ((KalmanFilterTOAIMU *)h)->getPose(p);
    });
    kfshim::fake_clock::ticks() -= dt_ns; // the poll must not move the filter's clock
    const double v[13] = {p.x, p.y, p.z, p.rotX, p.rotY, p.rotZ, p.rotW, p.linearSpeedX, p.linearSpeedY,
                          p.linearSpeedZ, p.angularSpeedX, p.angularSpeedY, p.angularSpeedZ};
    memcpy(pose13, v, sizeof v);
    if (rc == 0 && ok) rc = guarded([&] { for (int i = 0; i < 36; ++i) cov36[i] = p.covarianceMatrix(i); });
    return rc ? rc : (ok ? 0 : 4);
}

// ------------------------------------------------------- batch timing driver
// T6 replay of N filters x T steps over the SoA tensors of the C ABI (ranges f64
// metres [T][M][N]); used by bench.py as the "reference" CPU baseline.  One
// thread: the reference is single-threaded (node_pos.cpp:176-181); callers that
// want more cores run several processes.
int ref_t6_replay(long long N, int T, int M, const double *anchors, long long dt_ns, const double *ranges,
                  double err, double accel_noise, const double *x0, double *x_out, double *P_out) {
    std::vector<double> r(M), e(M, err);
    int rc_all = 0;
    for (long long f = 0; f < N; ++f) {
        void *h = ref_t6_create(accel_noise, 0, 0.0, x0[f], x0[N + f], x0[2 * N + f]);
        for (int t = 0; t < T; ++t) {
            for (int a = 0; a < M; ++a) r[a] = ranges[((long long)t * M + a) * N + f];
            rc_all |= ref_t6_toa(h, dt_ns, M, r.data(), anchors, e.data());
        }
        double pos[3], P[36];
        ref_t6_get(h, pos, P);
        for (int k = 0; k < 3; ++k) x_out[k * N + f] = pos[k];
        if (P_out)
            for (int k = 0; k < 36; ++k) P_out[k * N + f] = P[k];
        ref_t6_destroy(h);
    }
    return rc_all;
}

} // extern "C"

extern "C" {

// What KalmanFilter::init() parsed out of its five XML documents (KF.cpp:752-880), in the order
// of the kfpos_config fields they correspond to.
void ref_k8_config(void *h, double *out /* [15] */) {
    KalmanFilter *f = (KalmanFilter *)h;
    const double v[15] = {(double)f->mUseFixedHeight, f->mUWBtagZ, (double)f->mTagIdUWB,
                          (double)f->mUseFixedHeightPX4Flow, f->mPX4flowHeight, f->mPX4FlowArmP1, f->mPX4FlowArmP2,
                          f->mInitAnglePX4Flow, f->mCovarianceVelocityPX4Flow, f->mCovarianceGyroZPX4Flow,
                          (double)f->mUseImuFixedCovarianceAcceleration, f->mImuCovarianceAcceleration,
                          (double)f->mUseImuFixedCovarianceAngularVelocityZ, f->mUmuCovarianceAngularVelocityZ,
                          f->mMagAngleOffset};
    memcpy(out, v, sizeof v);
}
double ref_k8_config_mag_cov(void *h) { return ((KalmanFilter *)h)->mCovarianceMag; }
} // extern "C"
