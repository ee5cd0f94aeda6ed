/* report/timer.hpp — SHIM: the reference's Timer (Boost.Accumulators statistics per key) reduced to start/stop no-ops */
#ifndef ORACLE_SHIM_TIMER_HPP
#define ORACLE_SHIM_TIMER_HPP
#include <string>
class Timer {
   public:
    void start(const std::string &) {}
    void stop(const std::string &) {}
};
#endif
