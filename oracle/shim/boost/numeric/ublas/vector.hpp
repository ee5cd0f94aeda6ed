/* boost/numeric/ublas/vector.hpp — SHIM: the dense vector lives in the matrix shim */
#include <boost/numeric/ublas/matrix.hpp>
