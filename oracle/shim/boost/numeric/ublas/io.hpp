/* boost/numeric/ublas/io.hpp — empty SHIM (included, unused by the code built here) */
