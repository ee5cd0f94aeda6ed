/* log/log.hpp — SHIM: see ../log.hpp */
#include "../log.hpp"
