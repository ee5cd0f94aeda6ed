// sincos_qt.cuh — FP64 sincos in quarter-turn units, fused with the b_j-weighted accumulation.
//
// The amplitude kernels evaluate exp(i q.r) for ~1e13 (atom, frame, q-vector) triples, so the cost of
// one evaluation in FP64-pipe instructions is the whole game (B200: 64 DFMA/clk/SM, no tensor-core
// shortcut for a K=3 contraction followed by a transcendental).  Design:
//   * q is pre-scaled on the host by 2/pi, so u = q'.r is the phase in quarter turns (3 FP64 ops);
//   * k = rint(u) with the 1.5*2^52 magic-number add, f = u-k in [-1/2,1/2] (3 DADD, exact);
//   * sin(pi/2 f) = f*S(f^2), cos(pi/2 f) = 1+f^2*C(f^2), degree-5 near-minimax polynomials in f^2
//     (coefficients from tools/gen_sincos_coeffs.py; max abs error 3.4e-15 / 1.9e-16);
//   * the quadrant swap/sign is integer-pipe work (selects + sign-bit xor on b), not FP64.
// Total 21 FP64-pipe instructions per evaluation including the two accumulate FMAs.
// Valid for |u| < 2^31 quarter turns (|q.r| < 3.3e9 rad).
#pragma once
#include <cuda_runtime.h>

namespace sass {

// polynomial coefficients, see tools/gen_sincos_coeffs.py
#define SASS_S0 0x1.921fb54442cfap+0
#define SASS_S1 -0x1.4abbce6257a2ap-1
#define SASS_S2 0x1.466bc67123fa1p-4
#define SASS_S3 -0x1.32d2c644adc0bp-8
#define SASS_S4 0x1.5071ce4b47930p-13
#define SASS_S5 -0x1.dd54805f3f706p-19
#define SASS_C0 -0x1.3bd3cc9be45dbp+0
#define SASS_C1 0x1.03c1f081b4b74p-2
#define SASS_C2 -0x1.55d3c7e10153fp-6
#define SASS_C3 0x1.e1f50093e4c8fp-11
#define SASS_C4 -0x1.a6cc32d07dc98p-16
#define SASS_C5 0x1.f4b536b457d9ep-22

#define SASS_TWO_OVER_PI 0.63661977236758134308

// coefficient tables in constant memory: DFMA takes them as c[bank][offset] operands.  As literals ptxas re-materialises
// them with two UMOV each inside unrolled loops (9 % of the instructions the split self kernel executed).
__constant__ double kSinC[6] = {SASS_S0, SASS_S1, SASS_S2, SASS_S3, SASS_S4, SASS_S5};
__constant__ double kCosC[6] = {SASS_C0, SASS_C1, SASS_C2, SASS_C3, SASS_C4, SASS_C5};

// sin/cos of (pi/2)*u: returns cos in c, sin in s (full quadrant handling).
__device__ __forceinline__ void sincos_qt(double u, double &s, double &c) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    double t = u + MAGIC;
    int k = __double2loint(t);
    double kd = t - MAGIC;
    double f = u - kd;
    double z = f * f;
    double S = fma(z, kSinC[5], kSinC[4]);
    double Cp = fma(z, kCosC[5], kCosC[4]);
    S = fma(z, S, kSinC[3]);
    Cp = fma(z, Cp, kCosC[3]);
    S = fma(z, S, kSinC[2]);
    Cp = fma(z, Cp, kCosC[2]);
    S = fma(z, S, kSinC[1]);
    Cp = fma(z, Cp, kCosC[1]);
    S = fma(z, S, kSinC[0]);
    Cp = fma(z, Cp, kCosC[0]);
    double sv = f * S;
    double cv = fma(z, Cp, 1.0);
    bool odd = (k & 1) != 0;
    double cm = odd ? sv : cv;
    double sm = odd ? cv : sv;
    int chi = __double2hiint(cm) ^ (((k + 1) & 2) << 30);
    int shi = __double2hiint(sm) ^ ((k & 2) << 30);
    c = __hiloint2double(chi, __double2loint(cm));
    s = __hiloint2double(shi, __double2loint(sm));
}

// re += b*cos((pi/2)u), im += b*sin((pi/2)u).  The quadrant signs are applied to f (odd polynomial) and to the even
// polynomial value rather than to b, which keeps register-pair moves out of the loop (ptxas turns predicated FMAs back
// into FMA+select pairs, so the odd-quadrant swap stays four 32-bit selects):
//   sign(cos-type term) = bit1(k), sign(sin-type term) = bit1(k+1)   (DESIGN.md "quadrant algebra").
__device__ __forceinline__ void sincos_qt_accumulate2(double u, double b, double &re, double &im) {
    const double MAGIC = 6755399441055744.0;
    double t = u + MAGIC;
    const int k = __double2loint(t);
    double kd = t - MAGIC;
    double f = u - kd;
    double z = f * f;
    double S = fma(z, kSinC[5], kSinC[4]);
    double Cp = fma(z, kCosC[5], kCosC[4]);
    S = fma(z, S, kSinC[3]);
    Cp = fma(z, Cp, kCosC[3]);
    S = fma(z, S, kSinC[2]);
    Cp = fma(z, Cp, kCosC[2]);
    S = fma(z, S, kSinC[1]);
    Cp = fma(z, Cp, kCosC[1]);
    S = fma(z, S, kSinC[0]);
    Cp = fma(z, Cp, kCosC[0]);
    const int ks = k << 30;
    // signed odd part: (+-f) * S
    const double fs = __hiloint2double(__double2hiint(f) ^ ((ks + 0x40000000) & 0x80000000), __double2loint(f));
    const double sv = fs * S;
    double cv = fma(z, Cp, 1.0);
    cv = __hiloint2double(__double2hiint(cv) ^ (ks & 0x80000000), __double2loint(cv));
    // odd quadrant: (re, im) += b*(sv, cv); even: += b*(cv, sv)
    const bool odd = (k & 1) != 0;
    const double cm = odd ? sv : cv;
    const double sm = odd ? cv : sv;
    re = fma(b, cm, re);
    im = fma(b, sm, im);
}

}  // namespace sass
