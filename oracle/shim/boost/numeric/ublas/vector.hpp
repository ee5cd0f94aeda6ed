/* boost/numeric/ublas/vector.hpp — empty SHIM (included, unused by the code built here) */
