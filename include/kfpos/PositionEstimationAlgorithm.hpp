// PositionEstimationAlgorithm.hpp -- header-only C++11 mirror of the reference's algorithm
// classes on top of the C ABI (include/kfpos_b200.h).  Same class names, constructor
// arguments and virtual methods as
//   src/kfpos/algorithms/PositionEstimationAlgorithm.h:8-37, MLLocation.h:25-76,
//   KalmanFilterTOA.h:18-55, KalmanFilter.h:28-134, KalmanFilterTOAIMU.h:15-79
// so that PosGenerator (publishers/Posgenerator.cpp:99-141,476-496,510-548) compiles against it
// with the changes listed in INTEGRATION.md.  Each object wraps an N = 1 kfpos_batch; the
// batched use of the library goes through the C ABI directly.  No ROS / Armadillo / Boost.
#ifndef KFPOS_POSITION_ESTIMATION_ALGORITHM_HPP
#define KFPOS_POSITION_ESTIMATION_ALGORITHM_HPP

#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <limits>
#include <vector>

#include "../kfpos_b200.h"

namespace kfpos {

// stands in for arma::mat in Vector3 (sensor_types.h:7-13): element access m(i, j)
struct SmallMat {
    int n_rows, n_cols;
    double v[81];
    SmallMat() : n_rows(0), n_cols(0) { std::memset(v, 0, sizeof v); }
    void eye(int n, double s) {
        n_rows = n_cols = n;
        std::memset(v, 0, sizeof v);
        for (int i = 0; i < n; ++i) v[i * n + i] = s;
    }
    double &operator()(int i, int j) { return v[i * n_cols + j]; }
    double operator()(int i, int j) const { return v[i * n_cols + j]; }
};

struct Vector3 { // sensor_types.h:7-13
    double x, y, z;
    double rotX, rotY, rotZ, rotW;
    double linearSpeedX, linearSpeedY, linearSpeedZ;
    double angularSpeedX, angularSpeedY, angularSpeedZ;
    SmallMat covarianceMatrix;
    Vector3() : x(0), y(0), z(0), rotX(0), rotY(0), rotZ(0), rotW(0), linearSpeedX(0), linearSpeedY(0),
                linearSpeedZ(0), angularSpeedX(0), angularSpeedY(0), angularSpeedZ(0) {}
    Vector3(double x_, double y_, double z_) : x(x_), y(y_), z(z_), rotX(0), rotY(0), rotZ(0), rotW(0),
                linearSpeedX(0), linearSpeedY(0), linearSpeedZ(0), angularSpeedX(0), angularSpeedY(0),
                angularSpeedZ(0) {}
};
struct VectorDim3 { double x, y, z; };                 // sensor_types.h:15-17
struct Beacon { int id; int index; Vector3 position; }; // sensor_types.h:20-24

class KfposError : public std::runtime_error {
public:
    int code;
    KfposError(int c, const char *what) : std::runtime_error(std::string(what) + ": " + kfpos_strerror(c)), code(c) {}
};

// PositionEstimationAlgorithm.h:8-37
class PositionEstimationAlgorithm {
public:
    virtual ~PositionEstimationAlgorithm() { kfpos_batch_destroy(b_); }
    virtual bool init() { return true; }
    virtual bool getPose(Vector3 &) { return false; }
    virtual void newPX4FlowMeasurement(double, double, double, double, int) {}
    virtual void newTOAMeasurement(const std::vector<double> &, const std::vector<Beacon> &,
                                   const std::vector<double> &, double) {}
    virtual void newIMUMeasurement(VectorDim3, double[9], VectorDim3, double[9]) {}
    virtual void newMAGMeasurement(VectorDim3, double[9]) {}
    virtual void newCompassMeasurement(double) {}

    // additions: per-filter status word (KFPOS_ST_*) instead of exceptions, explicit clock
    int status() {
        int32_t st = 0;
        if (b_) kfpos_batch_get_state(b_, nullptr, nullptr, &st, nullptr);
        return st;
    }
    // the next update uses this dt instead of the steady_clock difference (offline replay)
    void setNextDt(double dt) { forced_dt_ = dt; }

protected:
    PositionEstimationAlgorithm() : b_(nullptr), forced_dt_(-1.0), started_(false) {}
    void create(int model, const kfpos_config &cfg, int device = 0) {
        int rc = kfpos_batch_create(&b_, device, model, 1, &cfg);
        if (rc) throw KfposError(rc, "kfpos_batch_create");
    }
    void check(int rc, const char *what) {
        if (rc) throw KfposError(rc, what);
    }
    // dt exactly as the reference takes it (TOA.cpp:74-88, KF.cpp:232-243, TOAIMU.cpp:106-117):
    // 0.1 s on the first update, the steady_clock difference afterwards
    double nextDt() {
        std::chrono::steady_clock::time_point now = std::chrono::steady_clock::now();
        double dt = 0.1;
        if (started_) dt = std::chrono::duration<double>(now - last_).count();
        if (forced_dt_ >= 0) dt = forced_dt_;
        forced_dt_ = -1.0;
        last_ = now;
        started_ = true;
        return dt;
    }
    double sinceLast() const {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - last_).count();
    }
    // beacons of this call become the batch's anchor table (PG.cpp:476-496 forwards only the
    // slots with a valid range, so the table may change from call to call)
    void toa(const std::vector<double> &rangings, const std::vector<Beacon> &beacons,
             const std::vector<double> &errorEstimations, double dt) {
        const size_t n = rangings.size();
        if (n > (size_t)KFPOS_MAX_ANCHORS) throw KfposError(KFPOS_ERR_INVALID, "too many beacons");
        std::vector<double> xyz(3 * (n ? n : 1));
        for (size_t i = 0; i < n; ++i) {
            xyz[3 * i] = beacons[i].position.x;
            xyz[3 * i + 1] = beacons[i].position.y;
            xyz[3 * i + 2] = beacons[i].position.z;
        }
        check(kfpos_batch_set_anchors(b_, (int)n, xyz.data()), "kfpos_batch_set_anchors");
        static const double none = 0.0;
        check(kfpos_batch_step_toa(b_, dt, n ? rangings.data() : &none, KFPOS_FMT_F64_M, 0.0,
                                   n ? errorEstimations.data() : nullptr, nullptr),
              "kfpos_batch_step_toa");
    }
    kfpos_batch *b_;
    double forced_dt_;
    bool started_;
    std::chrono::steady_clock::time_point last_;
};

// MLLocation.h:25-76.  newTOAMeasurement stores, getPose solves (ML.cpp:421-486).
class MLLocation : public PositionEstimationAlgorithm {
public:
    MLLocation() { setup(false, 0, 0, Vector3(1, 1, 4)); }
    MLLocation(bool use2d, int variant, int numRangingsToIgnore, const Vector3 &previousEstimation) {
        setup(use2d, variant, numRangingsToIgnore, previousEstimation);
    }
    bool init() override { return true; }
    void newTOAMeasurement(const std::vector<double> &rangings, const std::vector<Beacon> &beacons,
                           const std::vector<double> &errorEstimations, double) override {
        r_ = rangings; b_list_ = beacons; e_ = errorEstimations;
    }
    bool getPose(Vector3 &pose) override {
        const size_t n = r_.size();
        std::vector<double> xyz(3 * (n ? n : 1));
        for (size_t i = 0; i < n; ++i) {
            xyz[3 * i] = b_list_[i].position.x; xyz[3 * i + 1] = b_list_[i].position.y; xyz[3 * i + 2] = b_list_[i].position.z;
        }
        check(kfpos_batch_set_anchors(b_, (int)n, xyz.data()), "kfpos_batch_set_anchors");
        double pos[3], cov[9];
        static const double none = 0.0;
        check(kfpos_batch_ml_solve(b_, n ? r_.data() : &none, KFPOS_FMT_F64_M, 0.0, n ? e_.data() : nullptr, pos, cov,
                                   nullptr, nullptr, &status_, nullptr),
              "kfpos_batch_ml_solve");
        pose.x = pos[0]; pose.y = pos[1]; pose.z = pos[2];
        pose.rotX = pose.rotY = pose.rotZ = pose.rotW = 0.0;
        pose.covarianceMatrix.eye(6, 0.0); // ML.cpp:455-464 (the 2-D case reads out of bounds there)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) pose.covarianceMatrix(i, j) = cov[i * 3 + j];
        return true;
    }
    int mlStatus() const { return status_; }

private:
    void setup(bool use2d, int variant, int nIgnore, const Vector3 &prev) {
        kfpos_config c;
        kfpos_config_default(&c);
        c.use2d = use2d; c.variant = variant; c.num_ignored_rangings = nIgnore;
        c.ml_start[0] = prev.x; c.ml_start[1] = prev.y; c.ml_start[2] = prev.z;
        create(KFPOS_MODEL_ML, c);
        status_ = 0;
    }
    std::vector<double> r_, e_;
    std::vector<Beacon> b_list_;
    int32_t status_;
};

// KalmanFilterTOA.h:18-55 (fixed-initial-position constructor)
class KalmanFilterTOA : public PositionEstimationAlgorithm {
public:
    KalmanFilterTOA(double accelerationNoise, bool ignoreWorstAnchorMode, double ignoreCostThreshold,
                    Vector3 initialPosition) {
        kfpos_config c;
        kfpos_config_default(&c);
        c.accel_noise = accelerationNoise;
        c.ignore_worst_anchor = ignoreWorstAnchorMode;
        c.ignore_cost_threshold = ignoreCostThreshold;
        create(KFPOS_MODEL_T6, c);
        const double x[6] = {initialPosition.x, initialPosition.y, initialPosition.z, 0, 0, 0};
        check(kfpos_batch_set_state(b_, x, nullptr, nullptr), "kfpos_batch_set_state");
    }
    void newTOAMeasurement(const std::vector<double> &rangings, const std::vector<Beacon> &beacons,
                           const std::vector<double> &errorEstimations, double) override {
        toa(rangings, beacons, errorEstimations, nextDt());
    }
    bool getPose(Vector3 &pose) override { // TOA.cpp:438-473 + stateToPose :159-183
        if (!started_) return false;
        double x[6], P[36];
        check(kfpos_batch_get_pose(b_, sinceLast(), x, P, nullptr), "kfpos_batch_get_pose");
        pose.x = x[0]; pose.y = x[1]; pose.z = x[2];
        pose.rotX = pose.rotY = pose.rotZ = pose.rotW = 0.0;
        pose.covarianceMatrix.eye(6, 0.0);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) pose.covarianceMatrix(i, j) = P[i * 6 + j];
        return true;
    }
};

// KalmanFilter.h:28-134.  The five `filename*` arguments are the XML CONTENTS (the reference
// passes parameter names and reads the content with getParam, KF.cpp:759-764).
class KalmanFilter : public PositionEstimationAlgorithm {
public:
    KalmanFilter(double accelerationNoise, double initialAngle, double jolt, std::string xmlPos,
                 std::string xmlPX4Flow, std::string xmlTag, std::string xmlImu, std::string xmlMag,
                 Vector3 initialPosition)
        : ok_(true) {
        kfpos_config_default(&cfg_);
        cfg_.accel_noise = accelerationNoise;
        cfg_.jolt = jolt;
        cfg_.initial_angle = initialAngle;
        const std::string *xml[5] = {&xmlPX4Flow, &xmlTag, &xmlImu, &xmlMag, &xmlPos};
        for (int i = 0; i < 5; ++i)
            if (kfpos_config_load_xml(&cfg_, xml[i]->c_str()) != KFPOS_OK) ok_ = false; // init() -> false
        create(KFPOS_MODEL_K8, cfg_);
        const double x[8] = {initialPosition.x, initialPosition.y, 0, 0, 0, 0, initialAngle, 0};
        check(kfpos_batch_set_state(b_, x, nullptr, nullptr), "kfpos_batch_set_state");
    }
    // KF.cpp:6-32: no initial position -- the first epoch with rangings initialises the filter through
    // MLLocation (KF.cpp:244-285); what PosGenerator constructs with useStartPosition = 0 (PG.cpp:519-528)
    KalmanFilter(double accelerationNoise, double initialAngle, double jolt, std::string xmlPos,
                 std::string xmlPX4Flow, std::string xmlTag, std::string xmlImu, std::string xmlMag)
        : ok_(true) {
        kfpos_config_default(&cfg_);
        cfg_.accel_noise = accelerationNoise;
        cfg_.jolt = jolt;
        cfg_.initial_angle = initialAngle;
        const std::string *xml[5] = {&xmlPX4Flow, &xmlTag, &xmlImu, &xmlMag, &xmlPos};
        for (int i = 0; i < 5; ++i)
            if (kfpos_config_load_xml(&cfg_, xml[i]->c_str()) != KFPOS_OK) ok_ = false;
        cfg_.ml_initial_position = 1;
        create(KFPOS_MODEL_K8, cfg_);
        const double nan = std::numeric_limits<double>::quiet_NaN();
        const double x[8] = {nan, nan, 0, 0, 0, 0, initialAngle, 0};
        check(kfpos_batch_set_state(b_, x, nullptr, nullptr), "kfpos_batch_set_state");
    }
    bool init() override { return ok_; } // loadConfigurationFiles (KF.cpp:749-893)
    void newTOAMeasurement(const std::vector<double> &rangings, const std::vector<Beacon> &beacons,
                           const std::vector<double> &errorEstimations, double) override {
        toa(rangings, beacons, errorEstimations, nextDt());
    }
    void newPX4FlowMeasurement(double integrationX, double integrationY, double integrationRotationZ,
                               double integrationTime, int quality) override {
        if (quality == 0) return; // returns before the clock is read (KF.cpp:111-113)
        const int32_t q = quality;
        check(kfpos_batch_step_px4(b_, nextDt(), &integrationX, &integrationY, &integrationRotationZ,
                                   &integrationTime, &q, nullptr), "kfpos_batch_step_px4");
    }
    void newIMUMeasurement(VectorDim3 angularVelocity, double covarianceAngularVelocity[9],
                           VectorDim3 linearAcceleration, double covarianceAcceleration[9]) override {
        const double w[3] = {angularVelocity.x, angularVelocity.y, angularVelocity.z};
        const double a[3] = {linearAcceleration.x, linearAcceleration.y, linearAcceleration.z};
        check(kfpos_batch_step_imu(b_, nextDt(), w, covarianceAngularVelocity, a, covarianceAcceleration, nullptr),
              "kfpos_batch_step_imu");
    }
    void newMAGMeasurement(VectorDim3 mag, double[9]) override {
        const double m[3] = {mag.x, mag.y, mag.z};
        check(kfpos_batch_step_mag(b_, nextDt(), m, nullptr), "kfpos_batch_step_mag");
    }
    void newCompassMeasurement(double compass) override {
        check(kfpos_batch_step_compass(b_, nextDt(), &compass, nullptr), "kfpos_batch_step_compass");
    }
    bool getPose(Vector3 &pose) override { // KF.cpp:709-747 + stateToPose :324-363
        if (!started_) return false;
        double x[8], P[64];
        check(kfpos_batch_get_pose(b_, sinceLast(), x, P, nullptr), "kfpos_batch_get_pose");
        pose.x = x[0]; pose.y = x[1]; pose.z = cfg_.fixed_height;
        if (cfg_.ml_initial_position && !cfg_.use_fixed_height) { // mUWBtagZ = the ML estimate's z (KF.cpp:257)
            double pose13[13];
            check(kfpos_batch_get_pose_msg(b_, 0.0, pose13, nullptr, nullptr), "kfpos_batch_get_pose_msg");
            pose.z = pose13[2];
        }
        const double half = x[6] * 0.5;
        pose.rotX = 0.0; pose.rotY = 0.0; pose.rotZ = std::sin(half); pose.rotW = std::cos(half);
        pose.linearSpeedX = x[2]; pose.linearSpeedY = x[3]; pose.linearSpeedZ = 0.0;
        pose.angularSpeedX = 0.0; pose.angularSpeedY = 0.0; pose.angularSpeedZ = x[7];
        pose.covarianceMatrix.eye(6, 0.01);
        pose.covarianceMatrix(0, 0) = P[0 * 8 + 0]; pose.covarianceMatrix(0, 1) = P[0 * 8 + 1];
        pose.covarianceMatrix(1, 0) = P[1 * 8 + 0]; pose.covarianceMatrix(1, 1) = P[1 * 8 + 1];
        pose.covarianceMatrix(0, 5) = P[0 * 8 + 6]; pose.covarianceMatrix(1, 5) = P[1 * 8 + 6];
        pose.covarianceMatrix(5, 0) = P[6 * 8 + 0]; pose.covarianceMatrix(5, 1) = P[6 * 8 + 1];
        pose.covarianceMatrix(5, 5) = P[6 * 8 + 6];
        return true;
    }

private:
    kfpos_config cfg_;
    bool ok_;
};

// KalmanFilterTOAIMU.h:15-79 (IMU rows: the restatement of SURVEY App. B-5)
class KalmanFilterTOAIMU : public PositionEstimationAlgorithm {
public:
    KalmanFilterTOAIMU(double accelerationNoise, double jolt, Vector3 initialPosition) {
        kfpos_config c;
        kfpos_config_default(&c);
        c.accel_noise = accelerationNoise;
        c.jolt = jolt;
        create(KFPOS_MODEL_T9, c);
        const double x[9] = {initialPosition.x, initialPosition.y, initialPosition.z, 0, 0, 0, 0, 0, 0};
        check(kfpos_batch_set_state(b_, x, nullptr, nullptr), "kfpos_batch_set_state");
    }
    // TOAIMU.cpp:6-24: no initial position -- ML initialisation from the first epoch with rangings (:118-162)
    KalmanFilterTOAIMU(double accelerationNoise, double jolt) {
        kfpos_config c;
        kfpos_config_default(&c);
        c.accel_noise = accelerationNoise;
        c.jolt = jolt;
        c.ml_initial_position = 1;
        create(KFPOS_MODEL_T9, c);
        const double nan = std::numeric_limits<double>::quiet_NaN();
        const double x[9] = {nan, nan, nan, 0, 0, 0, 0, 0, 0};
        check(kfpos_batch_set_state(b_, x, nullptr, nullptr), "kfpos_batch_set_state");
    }
    void newTOAMeasurement(const std::vector<double> &rangings, const std::vector<Beacon> &beacons,
                           const std::vector<double> &errorEstimations, double) override {
        toa(rangings, beacons, errorEstimations, nextDt());
    }
    void newIMUMeasurement(VectorDim3, double[9], VectorDim3 linearAcceleration,
                           double covarianceAcceleration[9]) override {
        const double a[3] = {linearAcceleration.x, linearAcceleration.y, linearAcceleration.z};
        check(kfpos_batch_step_imu(b_, nextDt(), nullptr, nullptr, a, covarianceAcceleration, nullptr),
              "kfpos_batch_step_imu");
    }
    bool getPose(Vector3 &pose) override { // TOAIMU.cpp:476-510 + stateToPose :198-240
        if (!started_) return false;
        double x[9], P[81];
        check(kfpos_batch_get_pose(b_, sinceLast(), x, P, nullptr), "kfpos_batch_get_pose");
        pose.x = x[0]; pose.y = x[1]; pose.z = x[2];
        pose.rotX = pose.rotY = pose.rotZ = pose.rotW = 0.0;
        pose.linearSpeedX = x[3]; pose.linearSpeedY = x[4]; pose.linearSpeedZ = x[5];
        pose.angularSpeedX = x[6]; pose.angularSpeedY = x[7]; pose.angularSpeedZ = x[8];
        pose.covarianceMatrix.eye(9, 0.01);
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) pose.covarianceMatrix(i, j) = P[i * 9 + j];
            pose.covarianceMatrix(i, 7) = P[i * 9 + 8];
            pose.covarianceMatrix(7, i) = P[8 * 9 + i];
        }
        pose.covarianceMatrix(7, 7) = P[8 * 9 + 8];
        return true;
    }
};

} // namespace kfpos
#endif
