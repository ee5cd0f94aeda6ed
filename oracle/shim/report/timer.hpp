/* report/timer.hpp — SHIM: the reference's Timer (Boost.Accumulators statistics per key and thread) reduced to wall-clock sums
 * per key in one process-wide table, so that the harness can read e.g. "sd:runner" (the device's compute + write loop). */
#ifndef ORACLE_SHIM_TIMER_HPP
#define ORACLE_SHIM_TIMER_HPP
#include <chrono>
#include <map>
#include <mutex>
#include <string>
struct ShimTimerTable {
    std::mutex m;
    std::map<std::string, double> sum;
    static ShimTimerTable &Inst() {
        static ShimTimerTable t;
        return t;
    }
};
class Timer {
    std::map<std::string, std::chrono::steady_clock::time_point> started_;
   public:
    void start(const std::string &k) { started_[k] = std::chrono::steady_clock::now(); }
    void stop(const std::string &k) {
        std::map<std::string, std::chrono::steady_clock::time_point>::iterator it = started_.find(k);
        if (it == started_.end()) return;
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - it->second).count();
        std::lock_guard<std::mutex> l(ShimTimerTable::Inst().m);
        ShimTimerTable::Inst().sum[k] += dt;
    }
};
#endif
