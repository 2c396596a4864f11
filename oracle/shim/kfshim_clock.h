// Deterministic clock for the shim build (TEST INFRASTRUCTURE), force-included
// (-include) before the reference sources: the reference takes dt from
// std::chrono::steady_clock::now() (TOA.cpp:76-88, KF.cpp:234-243,
// TOAIMU.cpp:108-117); here `steady_clock` names a fake clock the
// harness advances by an exact number of nanoseconds before each call.
#pragma once
#include <chrono>
#include <condition_variable>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace kfshim {
struct fake_clock {
    typedef std::chrono::nanoseconds duration;
    typedef duration::rep rep;
    typedef duration::period period;
    typedef std::chrono::time_point<fake_clock> time_point;
    static const bool is_steady = true;
    static long long &ticks() {
        static long long t = 1000000000LL;
        return t;
    }
    static time_point now() { return time_point(duration(ticks())); }
};
} // namespace kfshim
namespace std {
namespace chrono {
typedef kfshim::fake_clock kfshim_fake_clock;
}
} // namespace std
#define steady_clock kfshim_fake_clock
