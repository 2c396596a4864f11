// ros/ros.h stub (TEST INFRASTRUCTURE): logging macros are no-ops, NodeHandle
// serves the string parameters the harness registered (the reference reads the
// CONTENT of its XML files through getParam, KF.cpp:759-764).
#pragma once
#include <map>
#include <string>

#define ROS_INFO(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)

namespace kfshim {
inline std::map<std::string, std::string> &params() {
    static std::map<std::string, std::string> p;
    return p;
}
} // namespace kfshim

namespace ros {
class NodeHandle {
public:
    NodeHandle() {}
    explicit NodeHandle(const std::string &) {}
    bool getParam(const std::string &name, std::string &out) const {
        std::map<std::string, std::string>::const_iterator it = kfshim::params().find(name);
        if (it == kfshim::params().end()) return false;
        out = it->second;
        return true;
    }
};
} // namespace ros
