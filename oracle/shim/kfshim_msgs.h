// Minimal ROS message / tf stand-ins (TEST INFRASTRUCTURE) so that the reference's
// publishers/Posgenerator.cpp compiles UNMODIFIED where it lies: only the members that
// file touches exist.  Field names follow the public ROS1 message definitions
// (std_msgs, geometry_msgs, sensor_msgs, nav_msgs, visualization_msgs, mavros_msgs) and
// gtec_msgs/Ranging of GTEC-UDC/rosmsgs (an un-vendored, unpinned dependency of the
// reference: uint16 anchorId, uint16 tagId, int32 range, int32 seq, int32 rss,
// float32 errorEstimation).
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "ros/ros.h"

#define KFSHIM_MSG_PTRS(T)                  \
    typedef std::shared_ptr<T> Ptr;         \
    typedef std::shared_ptr<const T> ConstPtr

namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
struct String { std::string data; KFSHIM_MSG_PTRS(String); };
struct Float64 { double data = 0.0; KFSHIM_MSG_PTRS(Float64); };
struct Int32MultiArray { std::vector<int32_t> data; KFSHIM_MSG_PTRS(Int32MultiArray); };
} // namespace std_msgs

namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; KFSHIM_MSG_PTRS(PoseStamped); };
struct PoseWithCovariance { Pose pose; std::array<double, 36> covariance{}; };
struct PoseWithCovarianceStamped {
    std_msgs::Header header;
    PoseWithCovariance pose;
    KFSHIM_MSG_PTRS(PoseWithCovarianceStamped);
};
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; std::array<double, 36> covariance{}; };
} // namespace geometry_msgs

namespace gtec_msgs {
struct Ranging {
    uint16_t anchorId = 0, tagId = 0;
    int32_t range = 0, seq = 0, rss = 0;
    float errorEstimation = 0.f;
    KFSHIM_MSG_PTRS(Ranging);
};
} // namespace gtec_msgs

namespace mavros_msgs {
struct OpticalFlowRad {
    std_msgs::Header header;
    uint32_t integration_time_us = 0;
    float integrated_x = 0, integrated_y = 0, integrated_xgyro = 0, integrated_ygyro = 0, integrated_zgyro = 0;
    int16_t temperature = 0;
    uint8_t quality = 0;
    uint32_t time_delta_distance_us = 0;
    float distance = 0;
    KFSHIM_MSG_PTRS(OpticalFlowRad);
};
} // namespace mavros_msgs

namespace sensor_msgs {
struct Imu {
    std_msgs::Header header;
    geometry_msgs::Quaternion orientation;
    std::array<double, 9> orientation_covariance{};
    geometry_msgs::Vector3 angular_velocity;
    std::array<double, 9> angular_velocity_covariance{};
    geometry_msgs::Vector3 linear_acceleration;
    std::array<double, 9> linear_acceleration_covariance{};
    KFSHIM_MSG_PTRS(Imu);
};
struct MagneticField {
    std_msgs::Header header;
    geometry_msgs::Vector3 magnetic_field;
    std::array<double, 9> magnetic_field_covariance{};
    KFSHIM_MSG_PTRS(MagneticField);
};
} // namespace sensor_msgs

namespace nav_msgs {
struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; KFSHIM_MSG_PTRS(Path); };
struct Odometry {
    std_msgs::Header header;
    std::string child_frame_id;
    geometry_msgs::PoseWithCovariance pose;
    geometry_msgs::TwistWithCovariance twist;
    KFSHIM_MSG_PTRS(Odometry);
};
} // namespace nav_msgs

namespace visualization_msgs {
struct Marker { std_msgs::Header header; int32_t id = 0; geometry_msgs::Pose pose; KFSHIM_MSG_PTRS(Marker); };
struct MarkerArray { std::vector<Marker> markers; KFSHIM_MSG_PTRS(MarkerArray); };
} // namespace visualization_msgs

namespace tf {
struct Quaternion { double x, y, z, w; Quaternion(double x_, double y_, double z_, double w_) : x(x_), y(y_), z(z_), w(w_) {} };
struct Vector3 { double x, y, z; Vector3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {} };
struct Transform { Quaternion q; Vector3 v; Transform(const Quaternion &q_, const Vector3 &v_) : q(q_), v(v_) {} };
struct StampedTransform {
    Transform t; ros::Time stamp; std::string frame, child;
    StampedTransform(const Transform &t_, const ros::Time &s, const std::string &f, const std::string &c)
        : t(t_), stamp(s), frame(f), child(c) {}
};
struct TransformBroadcaster { void sendTransform(const StampedTransform &) {} };
} // namespace tf
