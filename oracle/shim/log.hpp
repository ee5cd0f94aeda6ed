/* log.hpp — SHIM: the reference's Info / Warn / Err singletons reduced to message sinks (the real ones need Boost + MPI).
 * The code built here writes to them before a bare `throw;`. */
#ifndef ORACLE_SHIM_LOG_HPP
#define ORACLE_SHIM_LOG_HPP
#include <string>
#define ORACLE_SHIM_SINK(NAME)                          \
    class NAME {                                        \
        std::string last_;                              \
       public:                                          \
        static NAME *Inst() {                           \
            static NAME e;                              \
            return &e;                                  \
        }                                               \
        void write(const std::string &s) { last_ = s; } \
        const std::string &last() const { return last_; } \
    };
ORACLE_SHIM_SINK(Info)
ORACLE_SHIM_SINK(Warn)
ORACLE_SHIM_SINK(Err)
#endif
