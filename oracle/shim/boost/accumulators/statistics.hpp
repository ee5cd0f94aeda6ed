/* boost/accumulators/statistics.hpp — empty SHIM: included by the reference's decomposition_plan.hpp, nothing of it is used by the code built here */
