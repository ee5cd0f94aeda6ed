/* log/log.hpp — SHIM redirect: the message sinks of oracle/shim/log.hpp (the reference's log.cpp needs Boost.Format) */
#include "../../shim/log.hpp"
