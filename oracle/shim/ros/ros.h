// ros/ros.h stub (TEST INFRASTRUCTURE): logging macros are no-ops, NodeHandle
// serves the string parameters the harness registered (the reference reads the
// CONTENT of its XML files through getParam, KF.cpp:759-764).  For Posgenerator.cpp:
// a one-shot Timer that only records whether it is armed and when it is due on the
// fake clock (the harness plays the event loop), and a Publisher that keeps the last
// message of each type it was given.
#pragma once
#include <cstring> // the real ros.h brings it in (Posgenerator.cpp:501 relies on that)
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <typeinfo>

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
struct Time {
    double sec = 0.0;
    static Time now() { return Time(); }
};
struct Duration {
    double sec;
    explicit Duration(double s = 0.0) : sec(s) {}
};
struct TimerEvent {};

// shared state so that copies of a Timer (createTimer returns by value) stay one timer
struct TimerState {
    bool armed = false;
    long long period_ns = 0, due_ns = 0;
    std::function<void(const TimerEvent &)> cb;
};
class Timer {
public:
    std::shared_ptr<TimerState> st;
    void start() {
        if (!st) return;
        if (!st->armed) { // ros::Timer::start() on a running timer is a no-op
            st->armed = true;
            st->due_ns = kfshim::fake_clock::ticks() + st->period_ns;
        }
    }
    void stop() {
        if (st) st->armed = false;
    }
};

// keeps the last message of each type (type-erased) for the harness to read back
class Publisher {
public:
    std::shared_ptr<std::map<std::string, std::shared_ptr<void> > > last;
    Publisher() : last(new std::map<std::string, std::shared_ptr<void> >()) {}
    template <typename M>
    void publish(const M &m) const {
        (*last)[typeid(M).name()] = std::shared_ptr<void>(new M(m), [](void *p) { delete static_cast<M *>(p); });
    }
    template <typename M>
    const M *get() const {
        std::map<std::string, std::shared_ptr<void> >::const_iterator it = last->find(typeid(M).name());
        return it == last->end() ? nullptr : static_cast<const M *>(it->second.get());
    }
};

class NodeHandle {
public:
    NodeHandle() {}
    explicit NodeHandle(const std::string &) {}
    template <typename T>
    Timer createTimer(Duration period, void (T::*cb)(const TimerEvent &), T *obj, bool oneshot = false,
                      bool autostart = true) {
        Timer t;
        t.st.reset(new TimerState());
        t.st->period_ns = (long long)(period.sec * 1e9 + 0.5);
        t.st->cb = [obj, cb](const TimerEvent &e) { (obj->*cb)(e); };
        (void)oneshot; // the reference only creates one-shot timers
        if (autostart) t.start();
        return t;
    }
    bool getParam(const std::string &name, std::string &out) const {
        std::map<std::string, std::string>::const_iterator it = kfshim::params().find(name);
        if (it == kfshim::params().end()) return false;
        out = it->second;
        return true;
    }
};
} // namespace ros
