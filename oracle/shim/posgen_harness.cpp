// posgen_harness.cpp -- C entry points around the UNMODIFIED reference PosGenerator
// (TEST INFRASTRUCTURE; publishers/Posgenerator.cpp compiled where it lies against the ROS /
// tf / message stand-ins of this directory).  The harness plays ROS's event loop on the fake
// clock: before a ranging that arrives at time t is delivered, the one-shot ranging timer fires
// if it was armed and due before t (Posgenerator.cpp:11, :143-152).  A recording
// PositionEstimationAlgorithm sits in front of the real one (or alone) and keeps every epoch
// PosGenerator hands to newTOAMeasurement (:476-496).  Used only to pin oracle/ko_assemble.c
// and the log -> report pipeline in tests/; the product never loads it.
#define private public
#define protected public
#include "Posgenerator.h"
#undef private
#undef protected

#include <cstring>

namespace {

struct Epoch {
    std::vector<double> r, e;
    std::vector<int> index, id;
    double time_lag;
    long long t_ns;
};

class Recorder : public PositionEstimationAlgorithm {
public:
    std::unique_ptr<PositionEstimationAlgorithm> inner;
    std::vector<Epoch> epochs;
    int errors = 0;
    std::string last_error;
    bool init() override { return inner ? inner->init() : true; }
    bool getPose(Vector3 &pose) override { return inner ? inner->getPose(pose) : false; }
    void newPX4FlowMeasurement(double ix, double iy, double irz, double itime, int quality) override {
        if (inner) inner->newPX4FlowMeasurement(ix, iy, irz, itime, quality);
    }
    void newIMUMeasurement(VectorDim3 w, double cw[9], VectorDim3 a, double ca[9]) override {
        if (inner) inner->newIMUMeasurement(w, cw, a, ca);
    }
    void newMAGMeasurement(VectorDim3 mag, double cm[9]) override {
        if (inner) inner->newMAGMeasurement(mag, cm);
    }
    void newCompassMeasurement(double compass) override {
        if (inner) inner->newCompassMeasurement(compass);
    }
    void newTOAMeasurement(const std::vector<double> &rangings, const std::vector<Beacon> &beacons,
                           const std::vector<double> &errorEstimations, double timeLag) override {
        Epoch ep;
        ep.r = rangings;
        ep.e = errorEstimations;
        for (std::size_t i = 0; i < beacons.size(); ++i) {
            ep.index.push_back(beacons[i].index);
            ep.id.push_back(beacons[i].id);
        }
        ep.time_lag = timeLag;
        ep.t_ns = kfshim::fake_clock::ticks();
        epochs.push_back(ep);
        if (inner) {
            try {
                inner->newTOAMeasurement(rangings, beacons, errorEstimations, timeLag);
            } catch (const std::exception &ex) {
                ++errors;
                last_error = ex.what();
            } catch (...) {
                ++errors;
                last_error = "unknown";
            }
        }
    }
};

struct Handle {
    PosGenerator gen;
    Recorder *rec; // owned by gen.mPositionAlgorithm
    ros::Publisher pub, pub_path, pub_odom;
    long long t0_ns;
};

void fire_due_timer(Handle *h, long long t_ns) {
    ros::TimerState *st = h->gen.timerRanging.st.get();
    if (st && st->armed && st->due_ns < t_ns) {
        kfshim::fake_clock::ticks() = st->due_ns;
        st->armed = false; // one-shot
        st->cb(ros::TimerEvent());
    }
}

} // namespace

extern "C" {

// algorithm: -1 = recorder only, else ALGORITHM_* of Posgenerator.h:62-68 (2 = ML, 5 = KF_TOA, 6 = KF_TOA_IMU).
// anchors: n x (x, y, z); anchor_ids: the marker ids (the anchorId of the rangings).
// ALGORITHM_KF (4) reads its five XML documents from the parameters "kfpos_pos", "kfpos_px4",
// "kfpos_tag", "kfpos_imu", "kfpos_mag" (ref_set_param).
void *ref_pg_create(int algorithm, int tag_id, int n_anchors, const double *anchors, const int *anchor_ids,
                    double accel_noise, double jolt, int use_start, const double *start_xyz, double start_angle,
                    int ignore_worst, double cost_threshold, int use2d, int variant, int n_ignore) {
    Handle *h = new Handle();
    h->t0_ns = kfshim::fake_clock::ticks();
    PosGenerator &g = h->gen;
    g.setPublishers(h->pub, h->pub_path, h->pub_odom);
    g.setDynamicParameters(accel_noise, jolt);
    g.setStartParameters(use_start != 0, start_xyz[0], start_xyz[1], start_xyz[2], start_angle);
    g.setExternalFilesParameters("kfpos_pos", "kfpos_px4", "kfpos_tag", "kfpos_imu", "kfpos_mag");
    g.setHeuristicIgnore(ignore_worst != 0, cost_threshold);
    g.setDeviceIdentifiers(tag_id);
    g.setHeuristicML(use2d != 0, variant, n_ignore);
    h->rec = new Recorder();
    if (algorithm == ALGORITHM_ML) {
        // PosGenerator::setAlgorithm (Posgenerator.cpp:530-537) minus its init() call: MLLocation::init()
        // flows off the end of a bool function (MLLocation.cpp:415-417), which g++ -O2 compiles to a fall
        // through into the next function
        Vector3 s0 = {};
        if (use_start) { s0.x = start_xyz[0]; s0.y = start_xyz[1]; s0.z = start_xyz[2]; }
        else { s0.x = 1; s0.y = 1; s0.z = 4; }
        h->rec->inner.reset(new MLLocation(use2d != 0, variant, n_ignore, s0));
    } else if (algorithm >= 0) {
        g.setAlgorithm(algorithm);
        h->rec->inner.reset(g.mPositionAlgorithm.release());
    }
    g.mPositionAlgorithm.reset(h->rec);
    visualization_msgs::MarkerArray *arr = new visualization_msgs::MarkerArray();
    for (int i = 0; i < n_anchors; ++i) {
        visualization_msgs::Marker m;
        m.id = anchor_ids[i];
        m.pose.position.x = anchors[3 * i];
        m.pose.position.y = anchors[3 * i + 1];
        m.pose.position.z = anchors[3 * i + 2];
        arr->markers.push_back(m);
    }
    g.newAnchorsMarkerArray(visualization_msgs::MarkerArray::ConstPtr(arr));
    return h;
}

void ref_pg_destroy(void *hv) { delete (Handle *)hv; }

// L rangings in arrival order; t = arrival time in seconds since ref_pg_create (non-decreasing).
// flush_tail: let the timer fire after the last ranging.  Returns the number of epochs so far.
long long ref_pg_feed(void *hv, long long L, const int *anchor_id, const int *tag_id, const int *range_mm,
                      const int *seq, const double *err, const double *t, int flush_tail) {
    Handle *h = (Handle *)hv;
    for (long long i = 0; i < L; ++i) {
        const long long t_ns = h->t0_ns + (long long)(t[i] * 1e9 + 0.5);
        fire_due_timer(h, t_ns);
        kfshim::fake_clock::ticks() = t_ns;
        gtec_msgs::Ranging *m = new gtec_msgs::Ranging();
        m->anchorId = (uint16_t)anchor_id[i];
        m->tagId = (uint16_t)tag_id[i];
        m->range = range_mm[i];
        m->seq = seq[i];
        m->errorEstimation = err ? (float)err[i] : 0.f;
        h->gen.newTOAMeasurement(gtec_msgs::Ranging::ConstPtr(m));
    }
    if (flush_tail) fire_due_timer(h, (long long)1 << 62);
    return (long long)h->rec->epochs.size();
}

// One sensor message at time t (seconds since create), through PosGenerator's own callbacks
// (Posgenerator.cpp:100-140).  kind 1: PX4Flow v = {integrated_x, integrated_y, integrated_zgyro,
// integration_time_us, quality}; 2: IMU v = {angular_velocity[3], its covariance[9],
// linear_acceleration[3], its covariance[9]}; 3: magnetometer v = {field[3], covariance[9]};
// 4: compass v = {heading}.  Returns 0, or 1/2/3 when the algorithm threw.
int ref_pg_sensor(void *hv, int kind, double t, const double *v) {
    Handle *h = (Handle *)hv;
    const long long t_ns = h->t0_ns + (long long)(t * 1e9 + 0.5);
    fire_due_timer(h, t_ns);
    kfshim::fake_clock::ticks() = t_ns;
    try {
        if (kind == 1) {
            mavros_msgs::OpticalFlowRad *m = new mavros_msgs::OpticalFlowRad();
            m->integrated_x = (float)v[0]; m->integrated_y = (float)v[1]; m->integrated_zgyro = (float)v[2];
            m->integration_time_us = (uint32_t)v[3]; m->quality = (uint8_t)v[4];
            h->gen.newPX4FlowMeasurement(mavros_msgs::OpticalFlowRad::ConstPtr(m));
        } else if (kind == 2) {
            sensor_msgs::Imu *m = new sensor_msgs::Imu();
            m->angular_velocity.x = v[0]; m->angular_velocity.y = v[1]; m->angular_velocity.z = v[2];
            for (int i = 0; i < 9; ++i) m->angular_velocity_covariance[i] = v[3 + i];
            m->linear_acceleration.x = v[12]; m->linear_acceleration.y = v[13]; m->linear_acceleration.z = v[14];
            for (int i = 0; i < 9; ++i) m->linear_acceleration_covariance[i] = v[15 + i];
            h->gen.newIMUMeasurement(sensor_msgs::Imu::ConstPtr(m));
        } else if (kind == 3) {
            sensor_msgs::MagneticField *m = new sensor_msgs::MagneticField();
            m->magnetic_field.x = v[0]; m->magnetic_field.y = v[1]; m->magnetic_field.z = v[2];
            for (int i = 0; i < 9; ++i) m->magnetic_field_covariance[i] = v[3 + i];
            h->gen.newMAGMeasurement(sensor_msgs::MagneticField::ConstPtr(m));
        } else if (kind == 4) {
            std_msgs::Float64 m;
            m.data = v[0];
            h->gen.newCompassMeasurement(m);
        } else {
            return -1;
        }
    } catch (const std::runtime_error &) {
        return 1;
    } catch (const std::logic_error &) {
        return 2;
    } catch (...) {
        return 3;
    }
    return 0;
}

// Times (seconds since create) at which the recorded epochs reached the algorithm.
long long ref_pg_epoch_times(void *hv, long long max_epochs, double *t) {
    Handle *h = (Handle *)hv;
    const std::vector<Epoch> &ep = h->rec->epochs;
    for (long long k = 0; k < (long long)ep.size() && k < max_epochs; ++k) t[k] = (ep[k].t_ns - h->t0_ns) * 1e-9;
    return (long long)ep.size();
}

// Dense copy of the recorded epochs: ranges [max][M] in metres by beacon INDEX (0 = slot not in
// the epoch), err [max][M], time_lag [max].  Returns the number of recorded epochs.
long long ref_pg_epochs(void *hv, long long max_epochs, int M, double *ranges, double *err, double *time_lag) {
    Handle *h = (Handle *)hv;
    const std::vector<Epoch> &ep = h->rec->epochs;
    for (long long k = 0; k < (long long)ep.size() && k < max_epochs; ++k) {
        for (int a = 0; a < M; ++a) ranges[k * M + a] = err[k * M + a] = 0.0;
        for (std::size_t i = 0; i < ep[k].r.size(); ++i) {
            const int a = ep[k].index[i];
            if (a < 0 || a >= M) continue;
            ranges[k * M + a] = ep[k].r[i];
            err[k * M + a] = ep[k].e[i];
        }
        time_lag[k] = ep[k].time_lag;
    }
    return (long long)ep.size();
}

// publishFixedRateReport (Posgenerator.cpp:540-547) at time t (seconds since create): what the
// node would put on the wire.  pose13 = position, orientation (x, y, z, w), linear and angular
// twist of the Odometry message; cov36 = the pose covariance.  Returns 0 when a report was sent.
int ref_pg_report(void *hv, double t, double *pose13, double *cov36) {
    Handle *h = (Handle *)hv;
    const long long keep = kfshim::fake_clock::ticks();
    kfshim::fake_clock::ticks() = h->t0_ns + (long long)(t * 1e9 + 0.5);
    h->pub.last->clear();
    h->pub_odom.last->clear();
    int rc = 0;
    try {
        h->gen.publishFixedRateReport();
    } catch (...) {
        rc = 3;
    }
    kfshim::fake_clock::ticks() = keep; // the poll must not move the clock
    if (rc) return rc;
    const geometry_msgs::PoseWithCovarianceStamped *p = h->pub.get<geometry_msgs::PoseWithCovarianceStamped>();
    const nav_msgs::Odometry *o = h->pub_odom.get<nav_msgs::Odometry>();
    if (!p || !o) return 4;
    const double v[13] = {p->pose.pose.position.x,    p->pose.pose.position.y,    p->pose.pose.position.z,
                          p->pose.pose.orientation.x, p->pose.pose.orientation.y, p->pose.pose.orientation.z,
                          p->pose.pose.orientation.w, o->twist.twist.linear.x,    o->twist.twist.linear.y,
                          o->twist.twist.linear.z,    o->twist.twist.angular.x,   o->twist.twist.angular.y,
                          o->twist.twist.angular.z};
    memcpy(pose13, v, sizeof v);
    for (int i = 0; i < 36; ++i) cov36[i] = p->pose.covariance[i];
    return 0;
}

int ref_pg_errors(void *hv) { return ((Handle *)hv)->rec->errors; }
const char *ref_pg_last_error(void *hv) { return ((Handle *)hv)->rec->last_error.c_str(); }

} // extern "C"
