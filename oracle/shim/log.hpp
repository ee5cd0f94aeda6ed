/* log.hpp — empty SHIM: src/math/smath.cpp includes it but uses nothing of it (the real one needs Boost) */
