/* boost/math/special_functions.hpp — SHIM.  data_stager.cpp includes it without using it; multipole_scatter_device.cpp calls
 * sph_bessel, spherical_harmonic and cyl_bessel_j.  Boost.Math is not in this image, so the three functions are served by the
 * oracle's own restatements (sassena_oracle.c: orc_sph_bessel, orc_spherical_harmonic; the C library's jn for cyl_bessel_j),
 * which tests/test_oracle.py checks against scipy.  Building the reference's multipole devices over this shim therefore pins the
 * oracle's use of the functions (moment order, i^l prefactors, conjugation, summation order, normalisation, dsp, store) to the
 * reference's code -- not the functions' last bits. */
#ifndef ORACLE_SHIM_BOOST_MATH_SPECIAL_FUNCTIONS_HPP
#define ORACLE_SHIM_BOOST_MATH_SPECIAL_FUNCTIONS_HPP
#include <cmath>
#include <complex>
extern "C" {
double orc_sph_bessel(long l, double x);
void orc_spherical_harmonic(long n, long m, double theta, double phi, double *re, double *im);
}
namespace boost {
namespace math {
inline double sph_bessel(long l, double x) { return orc_sph_bessel(l, x); }
inline std::complex<double> spherical_harmonic(long n, long m, double theta, double phi) {
    double re, im;
    orc_spherical_harmonic(n, m, theta, phi, &re, &im);
    return std::complex<double>(re, im);
}
inline double cyl_bessel_j(long n, double x) { return ::jn((int)n, x); }
}  // namespace math
}  // namespace boost
#endif
