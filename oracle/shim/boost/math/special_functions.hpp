/* boost/math/special_functions.hpp — empty SHIM (data_stager.cpp includes it without using it) */
