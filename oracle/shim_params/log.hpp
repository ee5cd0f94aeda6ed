/* log.hpp — SHIM redirect to oracle/shim/log.hpp */
#include "../shim/log.hpp"
