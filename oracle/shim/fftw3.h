/* fftw3.h — SHIM for building the reference's src/math/smath.cpp in an image without FFTW3 (test infrastructure only).
 * It declares the handful of FFTW3 symbols that file uses; fftw_execute_dft runs the oracle's own DFT (orc_fft,
 * oracle/sassena_oracle.c).  What oracle/_ref/libsmath_ref.so pins is therefore the reference's OWN code around the
 * transform — zero padding, power spectrum, 1/(2NF(NF-tau)) normalisation, the direct O(NF^2) form and its
 * conjugation, square/multiply/add — not FFTW's arithmetic (any correct DFT agrees with it to rounding). */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef double fftw_complex[2];
typedef struct shim_fftw_plan_s {
    int n, sign;
} *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_destroy_plan(fftw_plan p);
void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out); /* in place or out of place, unnormalised */
#ifdef __cplusplus
}
#endif
#endif
