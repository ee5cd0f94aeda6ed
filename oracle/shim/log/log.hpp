/* log/log.hpp — SHIM: the reference's Err singleton reduced to a message sink (the real one needs Boost + MPI).
 * src/decomposition/assignment.cpp writes to it before its bare `throw;`. */
#ifndef ORACLE_SHIM_LOG_HPP
#define ORACLE_SHIM_LOG_HPP
#include <string>
class Err {
    std::string last_;
   public:
    static Err *Inst() {
        static Err e;
        return &e;
    }
    void write(const std::string &s) { last_ = s; }
    const std::string &last() const { return last_; }
};
#endif
