/* control.hpp — SHIM.  src/math/smath.cpp includes it without using it; src/decomposition/decomposition_plan.cpp reads
 * three values from the Params singleton (limits.decomposition.partitions.automatic / .size, limits.decomposition
 * .utilization; the real class parses them from scatter.xml, parameters.cpp:655-676, and needs Boost + libxml2).
 * Here the singleton is a plain struct the test harness fills in. */
#ifndef ORACLE_SHIM_CONTROL_HPP
#define ORACLE_SHIM_CONTROL_HPP
#include <cstddef>
struct ShimPartitions {
    bool automatic = true;
    size_t size = 1;
};
struct ShimDecomposition {
    ShimPartitions partitions;
    double utilization = 0.95;
};
struct ShimLimits {
    ShimDecomposition decomposition;
};
class Params {
   public:
    ShimLimits limits;
    static Params *Inst() {
        static Params p;
        return &p;
    }
};
#endif
