/*
 * sassena_oracle.c — CPU restatement of the Sassena scattering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (sassena_b200/) may link, load or call
 * this file.  It is the checker used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY PINNED TO THE REFERENCE'S OWN CODE for all four scatter devices: the reference
 * (benlabs/sassena v1.4.2) ships no golden vectors or asserting tests for this path and its
 * build system cannot be used here (Boost, FFTW3, MPI, HDF5, libxml2 are absent), but its
 * translation units compile where they lie over the shim headers in oracle/shim (make -C oracle
 * ref -> oracle/_ref/libsmath_ref.so, libparams_ref.so): AllVectorsScatterDevice,
 * SelfVectorsScatterDevice, MPSphereScatterDevice, MPCylinderScatterDevice with their abstract
 * bases, the stagers, smath.cpp, coor3d.cpp, assignment.cpp, decomposition_plan.cpp,
 * coordinate_writer.cpp, motion_walker.cpp, database.cpp, parameters.cpp (generators), atoms.cpp,
 * atomselection*.cpp, frames.cpp.  orc_compute_all_vectors / orc_compute_self_vectors /
 * orc_compute_mpsphere / orc_compute_mpcylinder reproduce the reference devices' fqt / fq / fq2
 * BIT FOR BIT (tests/test_reference_devices.py, fixtures tests/golden/ref_devices.npz,
 * ref_multipole_devices.npz); the helper functions and generators likewise (tests/test_oracle.py,
 * tests/test_reference_params.py).
 * NOT pinned by reference output: the VALUES of Boost.Math's sph_bessel / spherical_harmonic /
 * cyl_bessel_j (the reference's multipole code is built over this file's restatements of them;
 * scipy cross-checks and closed-form cluster averages pin the values), FFTW's arithmetic, and
 * the Boost.Random streams.  vendor/xdrfile-1.1.1 pins the product's XTC / TRR readers.
 *
 * Third-party arithmetic that is not in /root/reference and is restated here:
 *   FFTW3 (unpinned version)         -> own mixed-radix / Bluestein complex FFT
 *   Boost.Math sph_bessel, spherical_harmonic (Boost >= 1.42, unpinned)
 *                                    -> own recurrences, checked against scipy.special
 *   Boost.Random mt19937 + uniform_on_sphere (Boost 1.4x Box-Muller normal_distribution)
 *                                    -> best effort restatement (version dependent)
 *
 * Build: see oracle/Makefile (gcc -O3 -DNDEBUG, no -march, no fast-math: the reference's
 * Release flags, CMakeLists.txt:7,15-19).
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------ */
/* Decomposition arithmetic — src/decomposition/assignment.cpp:27-132                     */
/* ------------------------------------------------------------------------------------ */

/* DivAssignment(NN,rank,NAF): offset/size — assignment.cpp:27-35 */
void orc_div_assignment(size_t NN, size_t rank, size_t NAF, size_t *offset, size_t *size, size_t *max) {
    size_t first = (rank * NAF) / NN;
    size_t next = ((rank + 1) * NAF) / NN;
    *offset = first;
    *size = next - first;
    size_t m = NAF / NN; /* assignment.cpp:61-65 */
    if ((NAF % NN) != 0) m += 1;
    *max = m;
}

/* ModAssignment(NN,rank,NAF) — assignment.cpp:82-118; element i is rank + i*NN */
void orc_mod_assignment(size_t NN, size_t rank, size_t NAF, size_t *offset, size_t *size, size_t *max) {
    size_t s = NAF / NN;
    if ((NAF % NN) != 0) {
        if (rank < (NAF - NN * (NAF / NN))) s += 1;
    }
    *offset = rank;
    *size = s;
    size_t m = NAF / NN;
    if ((NAF % NN) != 0) m += 1;
    *max = m;
}

/* DecompositionParameters — src/decomposition/decomposition_plan.cpp:29-66.
 * Returns the penalty and fills NP, NAFcycles, NQcycles, nbytesize. */
size_t orc_decomposition_penalty(size_t NN, size_t NQ, size_t NAF, size_t NNpP, size_t elbytesize,
                                 size_t *NP_out, size_t *NAFcycles_out, size_t *NQcycles_out,
                                 size_t *nbytesize_out) {
    size_t NP = NN / NNpP;
    size_t NPused = NP;
    if (NQ < NP) NPused = NQ;
    size_t NNnotused = NN - NPused * NNpP;
    size_t NQcycles = ((NQ % NP) == 0) ? NQ / NP : NQ / NP + 1;
    size_t NAFcycles = ((NAF % NNpP) == 0) ? NAF / NNpP : NAF / NNpP + 1;
    size_t penalty = NNnotused * NQcycles * NAFcycles;
    penalty += (NPused * NQcycles - NQ) * (NNpP * NAFcycles);
    penalty += (NNpP * NAFcycles - NAF) * NQ;
    if (NP_out) *NP_out = NP;
    if (NAFcycles_out) *NAFcycles_out = NAFcycles;
    if (NQcycles_out) *NQcycles_out = NQcycles;
    if (nbytesize_out) *nbytesize_out = NAFcycles * elbytesize;
    return penalty;
}

/* DecompositionPlan automatic search — decomposition_plan.cpp:82-111 (+ utilisation :141-154).
 * Returns 0 on success, 1 when nothing fits the byte limit, 2 when utilisation is too low. */
int orc_decomposition_plan(size_t nn, size_t nq, size_t naf, size_t elbytesize, size_t nmaxbytesize,
                           double min_utilization, size_t *partitions, size_t *partitionsize,
                           size_t *penalty_out) {
    size_t npmax = naf;
    if (naf > nn) npmax = nn;
    int have = 0;
    size_t best_pen = 0, best_nnpp = 0, best_np = 0;
    for (size_t nnpp = npmax; nnpp >= 1; nnpp--) {
        size_t np, nafc, nqc, nbytes;
        size_t pen = orc_decomposition_penalty(nn, nq, naf, nnpp, elbytesize, &np, &nafc, &nqc, &nbytes);
        if (nbytes > nmaxbytesize) continue;
        if (!have || pen < best_pen) {
            have = 1;
            best_pen = pen;
            best_nnpp = nnpp;
            best_np = np;
        }
    }
    if (!have) return 1;
    *partitions = best_np;
    *partitionsize = best_nnpp;
    if (penalty_out) *penalty_out = best_pen;
    size_t used = nq * naf;
    double utilization = used * 1.0 / (used + best_pen);
    if (utilization < min_utilization) return 2;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Generators — src/control/parameters.cpp:930-1189                                       */
/* ------------------------------------------------------------------------------------ */

/* One scan unfolded — parameters.cpp:1136-1155.  Note powf(): float-rounded fractions. */
size_t orc_scan_unfold(const double base[3], double from, double to, size_t points, double exponent,
                       double *out /* [points][3] */) {
    size_t n = 0;
    if (points == 0) return 0;
    if (points == 1) {
        double scal = (from + to) / 2;
        for (int c = 0; c < 3; c++) out[c] = scal * base[c];
        return 1;
    }
    if (points == 2) {
        for (int c = 0; c < 3; c++) out[c] = from * base[c];
        for (int c = 0; c < 3; c++) out[3 + c] = to * base[c];
        return 2;
    }
    for (int c = 0; c < 3; c++) out[3 * n + c] = from * base[c];
    n++;
    for (size_t j = 1; j < (points - 1); j++) {
        double scal = from + powf((j * 1.0 / (points - 1)), exponent) * (to - from);
        for (int c = 0; c < 3; c++) out[3 * n + c] = scal * base[c];
        n++;
    }
    for (int c = 0; c < 3; c++) out[3 * n + c] = to * base[c];
    n++;
    return n;
}

/* mt19937 (Matsumoto-Nishimura), boost::mt19937 seeding: x[i]=1812433253*(x[i-1]^(x[i-1]>>30))+i */
typedef struct {
    uint32_t mt[624];
    int idx;
} orc_mt19937;

static void mt_seed(orc_mt19937 *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; i++) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static uint32_t mt_next(orc_mt19937 *g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

void orc_mt19937_stream(uint32_t seed, size_t n, uint32_t *out) {
    orc_mt19937 g;
    mt_seed(&g, seed);
    for (size_t i = 0; i < n; i++) out[i] = mt_next(&g);
}

/* boost::uniform_on_sphere<double>(dim) over boost::mt19937, Boost 1.4x semantics
 * (parameters.cpp:946-959 sphere, :1001-1013 cylinder): uniform_01 = x/2^32, Box-Muller
 * normal_distribution with one cached value, vector normalised by 1/sqrt(sum sq).
 * BEST EFFORT: the stream is Boost-version dependent. */
void orc_uniform_on_sphere(uint32_t seed, int dim, size_t count, double *out /* [count][3] */) {
    orc_mt19937 g;
    mt_seed(&g, seed);
    int valid = 0;
    double r1 = 0, cached_rho = 0;
    for (size_t i = 0; i < count; i++) {
        double v[3] = {0, 0, 0};
        double sqsum = 0;
        for (int d = 0; d < dim; d++) {
            double val;
            if (!valid) {
                r1 = mt_next(&g) / 4294967296.0;
                double r2 = mt_next(&g) / 4294967296.0;
                cached_rho = sqrt(-2.0 * log(1.0 - r2));
                valid = 1;
                val = cached_rho * cos(2 * M_PI * r1);
            } else {
                valid = 0;
                val = cached_rho * sin(2 * M_PI * r1);
            }
            v[d] = val;
            sqsum += val * val;
        }
        double inv = 1.0 / sqrt(sqsum);
        for (int d = 0; d < 3; d++) out[3 * i + d] = (d < dim) ? v[d] * inv : 0.0;
    }
}

/* vectors.type=file: normalise each row — parameters.cpp:932-943 */
void orc_normalize_rows(double *v, size_t n) {
    for (size_t i = 0; i < n; i++) {
        double x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
        double ql = sqrt(pow(x, 2) + pow(y, 2) + pow(z, 2)); /* coor3d.cpp:58-60 */
        if (ql != 0) {
            double s = 1.0 / ql;
            v[3 * i] = s * x;
            v[3 * i + 1] = s * y;
            v[3 * i + 2] = s * z;
        }
    }
}

/* cylinder raster_linear — parameters.cpp:1014-1023.  Returns count (out may be NULL to size). */
size_t orc_cylinder_raster_linear(size_t resolution, double *out) {
    const double M_2PI = 2 * M_PI;
    const double radincr = (M_2PI) / (360 * resolution);
    size_t n = 0;
    for (double phi = 0; phi < M_2PI; phi += radincr) {
        if (out) {
            out[3 * n] = cos(phi);
            out[3 * n + 1] = sin(phi);
            out[3 * n + 2] = 0;
        }
        n++;
    }
    return n;
}

/* multipole moments, type=resolution, sphere — parameters.cpp:1051-1060 */
size_t orc_moments_sphere(long resolution, long *out /* [n][2] */) {
    size_t n = 0;
    if (out) {
        out[0] = 0;
        out[1] = 0;
    }
    n++;
    for (long l = 1; l <= resolution; ++l)
        for (long m = -l; m <= l; ++m) {
            if (out) {
                out[2 * n] = l;
                out[2 * n + 1] = m;
            }
            n++;
        }
    return n;
}

/* CartesianVectorBase(axis) — coor3d.cpp:278-295.  base = {er, ephi, ez} row-major. */
static double len3(const double *v) { return sqrt(pow(v[0], 2) + pow(v[1], 2) + pow(v[2], 2)); }
static void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
void orc_vector_base(const double axis[3], double base[9]) {
    double ek[3] = {0, 0, 1}, ej[3] = {0, 1, 0};
    double ez[3], ekez[3], er[3], ezer[3], ephi[3];
    double al = len3(axis);
    for (int c = 0; c < 3; c++) ez[c] = axis[c] / al;
    cross3(ek, ez, ekez);
    if (len3(ekez) == 0) cross3(ej, ez, ekez);
    double l = len3(ekez);
    for (int c = 0; c < 3; c++) er[c] = ekez[c] / l;
    cross3(ez, er, ezer);
    l = len3(ezer);
    for (int c = 0; c < 3; c++) ephi[c] = ezer[c] / l;
    for (int c = 0; c < 3; c++) {
        base[c] = er[c];
        base[3 + c] = ephi[c];
        base[6 + c] = ez[c];
    }
}

/* AbstractVectorsScatterDevice::init_subvectors — abstract_vectors_scatter_device.cpp:112-175.
 * type: 0 none (no orientation vectors), 1 sphere/file, 2 cylinder.  Returns NM. */
size_t orc_init_subvectors(int type, const double q[3], const double *orient, size_t norient,
                           const double axis[3], double *out /* [max(norient,1)][3] */) {
    if (norient == 0 || type == 0) {
        for (int c = 0; c < 3; c++) out[c] = q[c];
        return 1;
    }
    if (type == 1) {
        double ql = len3(q);
        for (size_t i = 0; i < norient; i++)
            for (int c = 0; c < 3; c++) out[3 * i + c] = ql * orient[3 * i + c];
        return norient;
    }
    /* cylinder */
    double base[9];
    orc_vector_base(axis, base);
    double qp[3];
    for (int b = 0; b < 3; b++) qp[b] = q[0] * base[3 * b] + q[1] * base[3 * b + 1] + q[2] * base[3 * b + 2];
    double r = sqrt(pow(qp[0], 2) + pow(qp[1], 2)); /* CylinderCoor3D(cartesian) coor3d.cpp:116-118 */
    double z = qp[2];
    if (r == 0) { /* quirk kept: the *projected* vector is pushed (:134-135) */
        for (int c = 0; c < 3; c++) out[c] = qp[c];
        return 1;
    }
    for (size_t i = 0; i < norient; i++) {
        double vx = orient[3 * i], vy = orient[3 * i + 1];
        for (int c = 0; c < 3; c++) {
            /* qcylinder.z*base[2] + qcylinder.r*(vec.x*base[0] + vec.y*base[1]) (:141) */
            out[3 * i + c] = z * base[6 + c] + r * (vx * base[c] + vy * base[3 + c]);
        }
    }
    return norient;
}

/* cartesian -> spherical (r, phi, theta), SphericalCoor3D(CartesianCoor3D) coor3d.cpp:168-215,
 * narrowed to float as the stager does (data_stager.cpp:111-113).  in: float xyz [n][3] */
void orc_cart_to_spherical(const float *xyz, size_t n, float *out) {
    for (size_t i = 0; i < n; i++) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        double r = sqrt(pow(x, 2) + pow(y, 2) + pow(z, 2));
        double theta = 0, phi = 0;
        if (r != 0) {
            theta = acos(z / r);
            if (x != 0.0) {
                phi = atan(y / x);
                if (x < 0.0) phi += M_PI;
                else if (y < 0.0) phi += 2 * M_PI;
            } else if (y != 0.0) {
                if (y > 0) phi = M_PI_2;
                if (y < 0) phi = 3 * M_PI_2;
            } else {
                phi = 0.0;
            }
        }
        out[3 * i] = (float)r;
        out[3 * i + 1] = (float)phi;
        out[3 * i + 2] = (float)theta;
    }
}

/* ------------------------------------------------------------------------------------ */
/* FFT (stands in for FFTW3 c2c f64, smath.cpp:143,148; plans all_vectors...cpp:52-53)    */
/* ------------------------------------------------------------------------------------ */

static void fft_pow2(cplx *a, size_t n, int sign) {
    /* iterative radix-2, bit reversal; twiddles from long double sincos for accuracy */
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            cplx t = a[i];
            a[i] = a[j];
            a[j] = t;
        }
    }
    cplx *tw = (cplx *)malloc(sizeof(cplx) * (n / 2 + 1));
    for (size_t k = 0; k < n / 2; k++) {
        long double ang = sign * 2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)n;
        tw[k] = (double)cosl(ang) + I * (double)sinl(ang);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t step = n / len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; k++) {
                cplx u = a[i + k], v = a[i + k + len / 2] * tw[k * step];
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
    free(tw);
}

static size_t smallest_factor(size_t n) {
    for (size_t p = 2; p * p <= n; p++)
        if (n % p == 0) return p;
    return n;
}

static int is_smooth(size_t n, size_t maxp) {
    while (n > 1) {
        size_t p = smallest_factor(n);
        if (p > maxp) return 0;
        n /= p;
    }
    return 1;
}

/* recursive mixed radix DIT; tw = W_N^k table (k<N) for the top-level N */
static void fft_mixed_rec(const cplx *in, cplx *out, size_t n, size_t stride, const cplx *tw, size_t N) {
    if (n == 1) {
        out[0] = in[0];
        return;
    }
    size_t p = smallest_factor(n);
    size_t m = n / p;
    for (size_t r = 0; r < p; r++) fft_mixed_rec(in + r * stride, out + r * m, m, stride * p, tw, N);
    size_t tws = N / n;
    cplx tmp[16];
    for (size_t k = 0; k < m; k++) {
        for (size_t r = 0; r < p; r++) tmp[r] = out[r * m + k] * tw[(r * k * tws) % N];
        for (size_t j = 0; j < p; j++) {
            cplx s = 0;
            for (size_t r = 0; r < p; r++) s += tmp[r] * tw[((r * j * m) % n) * tws];
            out[j * m + k] = s;
        }
    }
}

static void fft_bluestein(cplx *a, size_t n, int sign) {
    size_t M = 1;
    while (M < 2 * n - 1) M <<= 1;
    cplx *w = (cplx *)malloc(sizeof(cplx) * n);
    cplx *A = (cplx *)calloc(M, sizeof(cplx));
    cplx *B = (cplx *)calloc(M, sizeof(cplx));
    for (size_t k = 0; k < n; k++) {
        /* k^2 mod 2n keeps the angle argument small */
        size_t k2 = (size_t)(((unsigned __int128)k * k) % (2 * n));
        long double ang = sign * 3.141592653589793238462643383279502884L * (long double)k2 / (long double)n;
        w[k] = (double)cosl(ang) + I * (double)sinl(ang);
    }
    for (size_t k = 0; k < n; k++) A[k] = a[k] * w[k];
    B[0] = conj(w[0]);
    for (size_t k = 1; k < n; k++) B[k] = B[M - k] = conj(w[k]);
    fft_pow2(A, M, -1);
    fft_pow2(B, M, -1);
    for (size_t k = 0; k < M; k++) A[k] *= B[k];
    fft_pow2(A, M, +1);
    for (size_t k = 0; k < n; k++) a[k] = (A[k] / (double)M) * w[k];
    free(w);
    free(A);
    free(B);
}

/* in-place unnormalised DFT of length n; sign=-1 forward (FFTW_FORWARD), +1 backward */
void orc_fft(double *data /* [n][2] */, size_t n, int sign) {
    cplx *a = (cplx *)data;
    if (n <= 1) return;
    if ((n & (n - 1)) == 0) {
        fft_pow2(a, n, sign);
        return;
    }
    if (is_smooth(n, 13)) {
        cplx *tw = (cplx *)malloc(sizeof(cplx) * n);
        cplx *out = (cplx *)malloc(sizeof(cplx) * n);
        for (size_t k = 0; k < n; k++) {
            long double ang = sign * 2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)n;
            tw[k] = (double)cosl(ang) + I * (double)sinl(ang);
        }
        fft_mixed_rec(a, out, n, 1, tw, n);
        memcpy(a, out, sizeof(cplx) * n);
        free(tw);
        free(out);
        return;
    }
    fft_bluestein(a, n, sign);
}

/* ------------------------------------------------------------------------------------ */
/* DSP — src/math/smath.cpp                                                               */
/* ------------------------------------------------------------------------------------ */

/* smath::auto_correlate_fftw(fftw_complex*,pF,pB,NF) — smath.cpp:141-156.
 * data has 2*NF entries, upper half zero on entry; first NF entries valid on exit. */
void orc_auto_correlate_fftw(double *data, size_t NF) {
    orc_fft(data, 2 * NF, -1);
    for (size_t i = 0; i < 2 * NF; ++i) {
        data[2 * i] = data[2 * i] * data[2 * i] + data[2 * i + 1] * data[2 * i + 1];
        data[2 * i + 1] = 0;
    }
    orc_fft(data, 2 * NF, +1);
    for (size_t i = 0; i < NF; ++i) {
        double factor = (1.0 / (2 * NF * (NF - i)));
        data[2 * i] *= factor;
        data[2 * i + 1] *= factor;
    }
}

/* smath::auto_correlate_direct(fftw_complex*,N) — smath.cpp:51-76 (a1*conj(a2): conjugate of fftw form) */
void orc_auto_correlate_direct(double *data, size_t N) {
    size_t NF = N;
    double *local = (double *)malloc(2 * N * sizeof(double));
    memcpy(local, data, 2 * N * sizeof(double));
    for (size_t tau = 0; tau < NF; ++tau) {
        data[2 * tau] = 0;
        data[2 * tau + 1] = 0;
        size_t last_starting_frame = NF - tau;
        for (size_t k = 0; k < last_starting_frame; ++k) {
            const double *a1 = &local[2 * k];
            const double *a2 = &local[2 * (k + tau)];
            data[2 * tau] += a1[0] * a2[0] + a1[1] * a2[1];
            data[2 * tau + 1] += -a1[0] * a2[1] + a1[1] * a2[0];
        }
        data[2 * tau] /= (last_starting_frame);
        data[2 * tau + 1] /= (last_starting_frame);
    }
    free(local);
}

/* smath::square_elements(fftw_complex*,N) — smath.cpp:168-177 */
void orc_square_elements(double *data, size_t N) {
    for (size_t n = 0; n < N; n++) {
        double r = data[2 * n] * data[2 * n] + data[2 * n + 1] * data[2 * n + 1];
        data[2 * n] = r;
        data[2 * n + 1] = 0;
    }
}

enum { ORC_DSP_AUTOCORRELATE = 0, ORC_DSP_SQUARE = 1, ORC_DSP_PLAIN = 2 };
enum { ORC_METHOD_FFTW = 0, ORC_METHOD_DIRECT = 1 };

/* dsp() — all_vectors_scatter_device.cpp:209-229.  at has 2*NF entries (alignpad'ed). */
static int dsp(double *at, size_t NF, int dsp_type, int dsp_method) {
    if (dsp_type == ORC_DSP_AUTOCORRELATE) {
        if (dsp_method == ORC_METHOD_DIRECT) orc_auto_correlate_direct(at, NF);
        else if (dsp_method == ORC_METHOD_FFTW) orc_auto_correlate_fftw(at, NF);
        else return 1;
    } else if (dsp_type == ORC_DSP_SQUARE) {
        orc_square_elements(at, NF);
    } else if (dsp_type != ORC_DSP_PLAIN) {
        return 1;
    }
    return 0;
}

/* store() — all_vectors_scatter_device.cpp:231-236: a = mean_t at[t]; afinal += a;
 * a2final += a*conj(a); atfinal += at  (smath::reduce :239-247, add_elements :209-215) */
static void store(const double *at, size_t NF, double *atfinal, cplx *afinal, cplx *a2final) {
    cplx s = 0;
    for (size_t i = 0; i < NF; ++i) s += at[2 * i] + I * at[2 * i + 1];
    cplx a = s * (1.0 / NF);
    *afinal += a;
    *a2final += a * conj(a);
    for (size_t n = 0; n < NF; n++) {
        atfinal[2 * n] += at[2 * n];
        atfinal[2 * n + 1] += at[2 * n + 1];
    }
}

/* ------------------------------------------------------------------------------------ */
/* Amplitudes                                                                             */
/* ------------------------------------------------------------------------------------ */

/* AllVectorsScatterDevice::scatter — all_vectors_scatter_device.cpp:418-438.
 * coords float [NF][NA][3]; writes at[f] = (Ar, Ai) for f in [f0, f1). */
void orc_scatter_all(const float *coords, size_t NA, size_t f0, size_t f1, const double *sfs,
                     const double q[3], double *at /* [NF][2] */) {
    double qx = q[0], qy = q[1], qz = q[2];
    for (size_t fi = f0; fi < f1; ++fi) {
        const float *p_data = &coords[fi * NA * 3];
        double Ar = 0, Ai = 0;
        for (size_t j = 0; j < NA; ++j) {
            double esf = sfs[j];
            float x = p_data[3 * j], y = p_data[3 * j + 1], z = p_data[3 * j + 2];
            double p = x * qx + y * qy + z * qz;
            Ar += esf * cos(p);
            Ai += esf * sin(p);
        }
        at[2 * fi] = Ar;
        at[2 * fi + 1] = Ai;
    }
}

/* SelfVectorsScatterDevice::scatter — self_vectors_scatter_device.cpp:288-322.
 * coords float [NF][3] of ONE atom; at has 2*NF entries, upper half zeroed. */
void orc_scatter_self(const float *p_data, size_t NF, double s, const double q[3], double *at) {
    double qx = q[0], qy = q[1], qz = q[2];
    for (size_t j = 0; j < NF; ++j) {
        float x1 = p_data[j * 3], y1 = p_data[j * 3 + 1], z1 = p_data[j * 3 + 2];
        double p1 = x1 * qx + y1 * qy + z1 * qz;
        double sp1 = sin(p1);
        double cp1 = cos(p1);
        at[2 * j] = s * cp1;
        at[2 * j + 1] = s * sp1;
    }
    memset(&at[2 * NF], 0, NF * 2 * sizeof(double));
}

/* boost::math::sph_bessel(l, x) stand-in.  Upward recurrence for x > l (stable), Miller
 * downward recurrence otherwise, power series for tiny x. */
double orc_sph_bessel(long l, double x) {
    if (x == 0) return (l == 0) ? 1.0 : 0.0;
    double ax = fabs(x);
    double sgn = (x < 0 && (l & 1)) ? -1.0 : 1.0;
    x = ax;
    if (x < 1e-3 || x * x < 0.01 * (2.0 * l + 3.0)) {
        /* series: x^l/(2l+1)!! * sum_k (-x^2/2)^k / (k! (2l+3)(2l+5)...(2l+2k+1)) */
        double pref = 1.0;
        for (long k = 1; k <= l; k++) pref *= x / (2.0 * k + 1.0);
        double term = 1.0, sum = 1.0;
        for (int k = 1; k < 60; k++) {
            term *= -0.5 * x * x / (k * (2.0 * l + 2.0 * k + 1.0));
            sum += term;
            if (fabs(term) < 1e-18 * fabs(sum)) break;
        }
        return sgn * pref * sum;
    }
    double j0 = sin(x) / x;
    if (l == 0) return j0;
    double j1 = sin(x) / (x * x) - cos(x) / x;
    if (l == 1) return sgn * j1;
    if (x > (double)l) {
        double jm = j0, jc = j1;
        for (long k = 1; k < l; k++) {
            double jn = (2.0 * k + 1.0) / x * jc - jm;
            jm = jc;
            jc = jn;
        }
        return sgn * jc;
    }
    /* Miller: start well above l */
    long start = l + 20 + (long)(sqrt(60.0 * (double)(l + 10)));
    double jp = 0.0, jc = 1e-280, jl = 0.0;
    for (long k = start; k >= 1; k--) {
        double jm = (2.0 * k + 1.0) / x * jc - jp; /* j_{k-1} */
        jp = jc;
        jc = jm;
        if (k - 1 == l) jl = jc;
        if (fabs(jc) > 1e250) {
            jc *= 1e-250;
            jp *= 1e-250;
            jl *= 1e-250;
        }
    }
    /* jc is now ~ j_0, jp ~ j_1 (unnormalised) */
    double scale = (fabs(j0) >= fabs(j1)) ? j0 / jc : j1 / jp;
    return sgn * jl * scale;
}

/* boost::math::spherical_harmonic(n, m, theta, phi) stand-in: theta polar, phi azimuth,
 * Condon-Shortley phase; Y_{n,-m} = (-1)^m conj(Y_{n,m}). Returns (re, im). */
void orc_spherical_harmonic(long n, long m, double theta, double phi, double *re, double *im) {
    long am = labs(m);
    if (am > n) {
        *re = 0;
        *im = 0;
        return;
    }
    double ct = cos(theta), st = sin(theta);
    double pmm = sqrt(1.0 / (4.0 * M_PI));
    for (long i = 1; i <= am; i++) pmm *= -sqrt((2.0 * i + 1.0) / (2.0 * i)) * st;
    double p;
    if (n == am) {
        p = pmm;
    } else {
        double pm1 = pmm;
        double pc = sqrt(2.0 * am + 3.0) * ct * pmm;
        for (long l = am + 2; l <= n; l++) {
            double a = sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)am * am));
            double b = sqrt((((double)l - 1.0) * (l - 1.0) - (double)am * am) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
            double pn = a * (ct * pc - b * pm1);
            pm1 = pc;
            pc = pn;
        }
        p = pc;
    }
    double c = cos(am * phi), s = sin(am * phi);
    if (m >= 0) {
        *re = p * c;
        *im = p * s;
    } else {
        double sg = (am & 1) ? -1.0 : 1.0;
        *re = sg * p * c;
        *im = -sg * p * s;
    }
}

/* MPSphereScatterDevice::scatter — multipole_scatter_device.cpp:467-497.
 * coords float [NF][NA][3] holding (r, phi, theta). */
void orc_scatter_mpsphere(const float *coords, size_t NA, size_t f0, size_t f1, const double *sfs,
                          double ql, long l, long m, double *at /* [NF][2] */) {
    double M_PI_four = 4 * M_PI;
    /* pow(complex<double>(0,1.0), l) -- multipole_scatter_device.cpp:483.  Under C++11 and later libstdc++ resolves this to
     * pow(complex<double>, double) = polar(exp(l * log|i|), l * arg(i)) = (cos(l pi/2), sin(l pi/2)) in double arithmetic, i.e.
     * i^l with a ~6e-17 residue in the component that is zero analytically (the C++98 overload pow(complex, int) multiplied
     * exactly).  The reference built here (oracle/_ref, g++ 13) does the former; restated so that the oracle reproduces it
     * bit for bit.  The product kernels use the exact i^l: the difference is ~1e-16 relative. */
    cplx il = cos((double)l * M_PI_2) + I * sin((double)l * M_PI_2);
    for (size_t fi = f0; fi < f1; ++fi) {
        const float *p_data = &coords[fi * NA * 3];
        cplx A = 0;
        for (size_t j = 0; j < NA; ++j) {
            double r = p_data[3 * j];
            double phi = p_data[3 * j + 1];
            double theta = p_data[3 * j + 2];
            double esf = sfs[j];
            double p = ql * r;
            cplx fmpiilesf = M_PI_four * il * esf;
            double aabess = orc_sph_bessel(l, p);
            double yr, yi;
            orc_spherical_harmonic(l, m, theta, phi, &yr, &yi);
            cplx aa = yr - I * yi; /* conj */
            A += fmpiilesf * aabess * aa;
        }
        at[2 * fi] = creal(A);
        at[2 * fi + 1] = cimag(A);
    }
}

/* ------------------------------------------------------------------------------------ */
/* compute() drivers                                                                      */
/* ------------------------------------------------------------------------------------ */

/* AllVectorsScatterDevice::compute (NNPP==1 branch) — all_vectors_scatter_device.cpp:238-361.
 * qvecs are the already-expanded subvectors (init_subvectors).  Threads play the role of the
 * reference's limits.computation.threads workers (one subvector each, :259-286); the
 * dsp/store loop stays sequential in m exactly like :270-279.
 * If at_out != NULL it receives the raw amplitudes [NM][NF][2] (pre-DSP). */
int orc_compute_all_vectors(const float *coords, size_t NF, size_t NA, const double *sfs,
                            const double *qvecs, size_t NM, int dsp_type, int dsp_method, int nthreads,
                            double *atfinal /* [NF][2] */, double afinal[2], double a2final[2],
                            double *at_out) {
    memset(atfinal, 0, NF * 2 * sizeof(double));
    cplx af = 0, a2f = 0;
    if (nthreads < 1) nthreads = 1;
    size_t NT = (size_t)nthreads;
    double *at_ = (double *)malloc(NF * NT * 2 * sizeof(double));
    double *nat = (double *)malloc(2 * NF * 2 * sizeof(double));
    int err = 0;
    for (size_t i = 0; i < NM; i += NT) {
        size_t cnt = (NM - i < NT) ? NM - i : NT;
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
        for (size_t j = 0; j < cnt; j++) {
            orc_scatter_all(coords, NA, 0, NF, sfs, &qvecs[3 * (i + j)], &at_[((j + i) % NT) * NF * 2]);
        }
        for (size_t j = 0; j < cnt; ++j) {
            const double *pat = &at_[((j + i) % NT) * NF * 2];
            if (at_out) memcpy(&at_out[(i + j) * NF * 2], pat, NF * 2 * sizeof(double));
            /* alignpad :186-207 */
            memcpy(nat, pat, NF * 2 * sizeof(double));
            memset(&nat[2 * NF], 0, NF * 2 * sizeof(double));
            err |= dsp(nat, NF, dsp_type, dsp_method);
            store(nat, NF, atfinal, &af, &a2f);
        }
    }
    free(at_);
    free(nat);
    double factor = 1.0 / NM; /* :355-360 */
    for (size_t n = 0; n < NF; n++) {
        atfinal[2 * n] *= factor;
        atfinal[2 * n + 1] *= factor;
    }
    af *= factor;
    a2f *= factor;
    afinal[0] = creal(af);
    afinal[1] = cimag(af);
    a2final[0] = creal(a2f);
    a2final[1] = cimag(a2f);
    return err;
}

/* Frame-decomposed variant used only as the CPU timing baseline: the reference's MPI ranks
 * split frames with DivAssignment (all_vectors...cpp:408,418); here P threads each take a
 * DivAssignment slice of the frames for every subvector. Same results as above. */
int orc_compute_all_vectors_framesplit(const float *coords, size_t NF, size_t NA, const double *sfs,
                                       const double *qvecs, size_t NM, int dsp_type, int dsp_method,
                                       int nranks, double *atfinal, double afinal[2], double a2final[2]) {
    memset(atfinal, 0, NF * 2 * sizeof(double));
    cplx af = 0, a2f = 0;
    if (nranks < 1) nranks = 1;
    double *at_ = (double *)malloc(NF * 2 * sizeof(double));
    double *nat = (double *)malloc(2 * NF * 2 * sizeof(double));
    int err = 0;
    for (size_t i = 0; i < NM; i++) {
#pragma omp parallel for num_threads(nranks) schedule(static, 1)
        for (int r = 0; r < nranks; r++) {
            size_t off, sz, mx;
            orc_div_assignment((size_t)nranks, (size_t)r, NF, &off, &sz, &mx);
            orc_scatter_all(coords, NA, off, off + sz, sfs, &qvecs[3 * i], at_);
        }
        memcpy(nat, at_, NF * 2 * sizeof(double));
        memset(&nat[2 * NF], 0, NF * 2 * sizeof(double));
        err |= dsp(nat, NF, dsp_type, dsp_method);
        store(nat, NF, atfinal, &af, &a2f);
    }
    free(at_);
    free(nat);
    double factor = 1.0 / NM;
    for (size_t n = 0; n < NF; n++) {
        atfinal[2 * n] *= factor;
        atfinal[2 * n + 1] *= factor;
    }
    af *= factor;
    a2f *= factor;
    afinal[0] = creal(af);
    afinal[1] = cimag(af);
    a2final[0] = creal(a2f);
    a2final[1] = cimag(a2f);
    return err;
}

/* SelfVectorsScatterDevice::compute — self_vectors_scatter_device.cpp:145-239.
 * coords float [NA_local][NF][3] (atom-major, DataStagerByAtom layout); sfs_local[n] is the
 * factor of local atom n (scatterfactors.get(assignment_[ai]) :291).  Sequential over atoms
 * then vectors like :167-195; threads emulate the reference's worker threads over vectors.
 * The final 1/NM scale is applied (rank-0 view of a single-rank partition). */
int orc_compute_self_vectors(const float *coords, size_t NA_local, size_t NF, const double *sfs_local,
                             const double *qvecs, size_t NM, int dsp_type, int dsp_method, int nthreads,
                             double *atfinal, double afinal[2], double a2final[2]) {
    memset(atfinal, 0, NF * 2 * sizeof(double));
    cplx af = 0, a2f = 0;
    if (nthreads < 1) nthreads = 1;
    size_t NT = (size_t)nthreads;
    double *at_ = (double *)malloc(2 * NF * NT * 2 * sizeof(double));
    int err = 0;
    for (size_t n = 0; n < NA_local; ++n) {
        for (size_t i = 0; i < NM; i += NT) {
            size_t cnt = (NM - i < NT) ? NM - i : NT;
            int errs = 0;
#pragma omp parallel for num_threads(nthreads) schedule(static, 1) reduction(| : errs)
            for (size_t j = 0; j < cnt; j++) {
                double *pat = &at_[((j + i) % NT) * 2 * NF * 2];
                orc_scatter_self(&coords[n * NF * 3], NF, sfs_local[n], &qvecs[3 * (i + j)], pat);
                errs |= dsp(pat, NF, dsp_type, dsp_method); /* dsp is per-timeline independent */
            }
            err |= errs;
            for (size_t j = 0; j < cnt; ++j) {
                store(&at_[((j + i) % NT) * 2 * NF * 2], NF, atfinal, &af, &a2f);
            }
        }
    }
    free(at_);
    double factor = 1.0 / NM; /* :233-238 */
    for (size_t n = 0; n < NF; n++) {
        atfinal[2 * n] *= factor;
        atfinal[2 * n + 1] *= factor;
    }
    af *= factor;
    a2f *= factor;
    afinal[0] = creal(af);
    afinal[1] = cimag(af);
    a2final[0] = creal(a2f);
    a2final[1] = cimag(a2f);
    return err;
}

/* MPSphereScatterDevice::compute (NNPP==1) — multipole_scatter_device.cpp:278-401; norm 1/(4 pi). */
int orc_compute_mpsphere(const float *coords_sph, size_t NF, size_t NA, const double *sfs, double ql,
                         const long *moments /* [NM][2] */, size_t NM, int dsp_type, int dsp_method,
                         int nthreads, double *atfinal, double afinal[2], double a2final[2],
                         double *at_out) {
    memset(atfinal, 0, NF * 2 * sizeof(double));
    cplx af = 0, a2f = 0;
    if (nthreads < 1) nthreads = 1;
    size_t NT = (size_t)nthreads;
    double *at_ = (double *)malloc(NF * NT * 2 * sizeof(double));
    double *nat = (double *)malloc(2 * NF * 2 * sizeof(double));
    int err = 0;
    for (size_t i = 0; i < NM; i++)
        if (labs(moments[2 * i + 1]) > moments[2 * i]) return 2; /* :459-465 */
    for (size_t i = 0; i < NM; i += NT) {
        size_t cnt = (NM - i < NT) ? NM - i : NT;
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
        for (size_t j = 0; j < cnt; j++) {
            orc_scatter_mpsphere(coords_sph, NA, 0, NF, sfs, ql, moments[2 * (i + j)], moments[2 * (i + j) + 1],
                                 &at_[((j + i) % NT) * NF * 2]);
        }
        for (size_t j = 0; j < cnt; ++j) {
            const double *pat = &at_[((j + i) % NT) * NF * 2];
            if (at_out) memcpy(&at_out[(i + j) * NF * 2], pat, NF * 2 * sizeof(double));
            memcpy(nat, pat, NF * 2 * sizeof(double));
            memset(&nat[2 * NF], 0, NF * 2 * sizeof(double));
            err |= dsp(nat, NF, dsp_type, dsp_method);
            store(nat, NF, atfinal, &af, &a2f);
        }
    }
    free(at_);
    free(nat);
    double factor = 1.0 / (4 * M_PI); /* :395-400 */
    for (size_t n = 0; n < NF; n++) {
        atfinal[2 * n] *= factor;
        atfinal[2 * n + 1] *= factor;
    }
    af *= factor;
    a2f *= factor;
    afinal[0] = creal(af);
    afinal[1] = cimag(af);
    a2final[0] = creal(a2f);
    a2final[1] = cimag(a2f);
    return err;
}

/* ------------------------------------------------------------------------------------ */
/* multipole cylinder                                                                    */
/* ------------------------------------------------------------------------------------ */

/* include/math/coor3d.hpp:30 -- note the FLOAT arguments and result: sign(M_PI, y) is float-rounded pi */
static float sign_f(float a, float b) { return (b < 0.0) ? -a : a; }

/* CylinderCoor3D(CartesianCoor3D) -- coor3d.cpp:113-138 */
static void cart_to_cyl(double x, double y, double z, double *r, double *phi, double *zo) {
    *r = sqrt(pow(x, 2) + pow(y, 2));
    double p;
    if (x != 0.0) {
        p = atan(y / x);
        if (x < 0.0) p = sign_f(M_PI, y) + p;
    } else if (y != 0.0) {
        p = sign_f(M_PI_2, y);
    } else {
        p = 0.0;
    }
    p = p < 0 ? 2 * M_PI + p : p;
    *phi = p;
    *zo = z;
}

/* CylindricalCoordinateSet(cs, axis) -- src/sample/coordinate_set.cpp:278-296: project on CartesianVectorBase(axis)
 * (coor3d.cpp:278-304), convert, then the stager's narrowing to float (data_stager.cpp:111-113).  in: float xyz [n][3] */
void orc_cart_to_cylindrical(const float *xyz, size_t n, const double axis[3], float *out) {
    double base[9];
    orc_vector_base(axis, base);
    for (size_t i = 0; i < n; i++) {
        double c[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        double p[3];
        for (int k = 0; k < 3; k++) p[k] = c[0] * base[3 * k] + c[1] * base[3 * k + 1] + c[2] * base[3 * k + 2];
        double r, phi, z;
        cart_to_cyl(p[0], p[1], p[2], &r, &phi, &z);
        out[3 * i] = (float)r;
        out[3 * i + 1] = (float)phi;
        out[3 * i + 2] = (float)z;
    }
}

/* MPCylinderScatterDevice::scatter -- multipole_scatter_device.cpp:905-985.  coords float [NF][NA][3] = (r, phi, z).
 * Boost's cyl_bessel_j is restated with the C library's jn() (both are accurate to a few ulp). */
void orc_scatter_mpcylinder(const float *coords, size_t NA, size_t f0, size_t f1, const double *sfs, const double q[3],
                            const double axis[3], long l, long m, double *at /* [NF][2] */) {
    double base[9];
    orc_vector_base(axis, base);
    double qp[3];
    for (int k = 0; k < 3; k++) qp[k] = q[0] * base[3 * k] + q[1] * base[3 * k + 1] + q[2] * base[3 * k + 2];
    double qr, qphi, qz;
    cart_to_cyl(qp[0], qp[1], qp[2], &qr, &qphi, &qz);
    for (size_t fi = f0; fi < f1; ++fi) {
        const float *p_data = &coords[fi * NA * 3];
        cplx A = 0;
        for (size_t j = 0; j < NA; ++j) {
            double r = p_data[3 * j];
            double phi = p_data[3 * j + 1];
            double z = p_data[3 * j + 2];
            double esf = sfs[j];
            double parallel_sign = 1.0;
            if ((z != 0) && (qz != 0)) parallel_sign = (z * qz) / (fabs(z) * fabs(qz));
            cplx expi = cexp(I * (parallel_sign * z * qz));
            double p = r * qr;
            double psiphi = phi - qphi;
            if (m == 0 && l == 0) {
                A += expi * jn(0, p) * esf;
            } else if (m == 0) {
                cplx fac1 = 2.0 * powf(-1.0, l) * jn((int)(2 * l), p);
                A += sqrt(0.5) * fac1 * expi * cos(2 * l * psiphi) * esf;
            } else if (m == 1) {
                cplx fac1 = 2.0 * powf(-1.0, l) * jn((int)(2 * l), p);
                A += sqrt(0.5) * fac1 * expi * sin(2 * l * psiphi) * esf;
            } else if (m == 2) {
                cplx fac2 = I * (double)(2.0 * powf(-1.0, l - 1) * jn((int)(2 * l - 1), p));
                A += sqrt(0.5) * fac2 * expi * cos((2 * l - 1) * psiphi) * esf;
            } else if (m == 3) {
                cplx fac2 = I * (double)(2.0 * powf(-1.0, l - 1) * jn((int)(2 * l - 1), p));
                A += sqrt(0.5) * fac2 * expi * sin((2 * l - 1) * psiphi) * esf;
            }
        }
        double norm = sqrt(2 * M_PI);
        at[2 * fi] = norm * creal(A);
        at[2 * fi + 1] = norm * cimag(A);
    }
}

/* MPCylinderScatterDevice::compute (NNPP==1) -- multipole_scatter_device.cpp:700-870; norm 1/(2 pi) (:866).
 * Moment validity: parameters.cpp:1082-1102. */
int orc_compute_mpcylinder(const float *coords_cyl, size_t NF, size_t NA, const double *sfs, const double q[3],
                           const double axis[3], const long *moments /* [NM][2] */, size_t NM, int dsp_type, int dsp_method,
                           int nthreads, double *atfinal, double afinal[2], double a2final[2], double *at_out) {
    memset(atfinal, 0, NF * 2 * sizeof(double));
    cplx af = 0, a2f = 0;
    if (nthreads < 1) nthreads = 1;
    size_t NT = (size_t)nthreads;
    for (size_t i = 0; i < NM; i++) {
        long l = moments[2 * i], m = moments[2 * i + 1];
        if (l < 0 || m < 0 || m > 3 || (l == 0 && m != 0)) return 2;
    }
    double *at_ = (double *)malloc(NF * NT * 2 * sizeof(double));
    double *nat = (double *)malloc(2 * NF * 2 * sizeof(double));
    int err = 0;
    for (size_t i = 0; i < NM; i += NT) {
        size_t cnt = (NM - i < NT) ? NM - i : NT;
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
        for (size_t j = 0; j < cnt; j++) {
            orc_scatter_mpcylinder(coords_cyl, NA, 0, NF, sfs, q, axis, moments[2 * (i + j)], moments[2 * (i + j) + 1],
                                   &at_[((j + i) % NT) * NF * 2]);
        }
        for (size_t j = 0; j < cnt; ++j) {
            const double *pat = &at_[((j + i) % NT) * NF * 2];
            if (at_out) memcpy(&at_out[(i + j) * NF * 2], pat, NF * 2 * sizeof(double));
            memcpy(nat, pat, NF * 2 * sizeof(double));
            memset(&nat[2 * NF], 0, NF * 2 * sizeof(double));
            err |= dsp(nat, NF, dsp_type, dsp_method);
            store(nat, NF, atfinal, &af, &a2f);
        }
    }
    free(at_);
    free(nat);
    double factor = 1.0 / (2 * M_PI);
    for (size_t n = 0; n < NF; n++) {
        atfinal[2 * n] *= factor;
        atfinal[2 * n + 1] *= factor;
    }
    af *= factor;
    a2f *= factor;
    afinal[0] = creal(af);
    afinal[1] = cimag(af);
    a2final[0] = creal(a2f);
    a2final[1] = cimag(a2f);
    return err;
}

/* moments generator, cylinder branch -- parameters.cpp:1062-1070 */
size_t orc_moments_cylinder(long resolution, long *out /* [n][2] */) {
    size_t n = 0;
    if (out) { out[0] = 0; out[1] = 0; }
    n++;
    for (long l = 1; l <= resolution; ++l)
        for (long m = 0; m <= 3; ++m) {
            if (out) { out[2 * n] = l; out[2 * n + 1] = m; }
            n++;
        }
    return n;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
