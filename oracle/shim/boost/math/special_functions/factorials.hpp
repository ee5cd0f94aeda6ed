/* boost/math/special_functions/factorials.hpp — SHIM: factorial<double>(n).  Boost returns a table of correctly rounded n!;
 * the running product below is exact up to 22! and the Slater form factors (database.cpp:504) ask for (2n)! with n <= 5. */
#ifndef ORACLE_SHIM_BOOST_MATH_FACTORIALS_HPP
#define ORACLE_SHIM_BOOST_MATH_FACTORIALS_HPP
namespace boost {
namespace math {
template <class T> inline T factorial(unsigned n) {
    T r = 1;
    for (unsigned i = 2; i <= n; i++) r *= (T)i;
    return r;
}
}  // namespace math
}  // namespace boost
#endif
