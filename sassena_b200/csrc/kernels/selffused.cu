// selffused.cu — fused amplitude + FFT autocorrelation for incoherent self scattering (K2+K3 fused).
//
// Restates, per (atom n, q-vector m), SelfVectorsScatterDevice::scatter (reference
// src/scatter_devices/self_vectors_scatter_device.cpp:288-322) followed by smath::auto_correlate_fftw
// (src/math/smath.cpp:141-156) and store() (self...:137-142), for ~1e8 timelines per |q|.  Writing the timelines to HBM
// and transforming them there costs >1 MB of traffic each; here a timeline never leaves the SM:
//
//   * padded length L = R * N  (N = 2^k <= 4096, R = ceil((2NF-1)/N), any integer): the L-point spectrum splits into
//     R residues  X[R k + j] = FFT_N(y_j)[k],   y_j[n'] = sum_p x[n' + N p] * exp(-2 pi i (n'+N p) j / L);
//   * a CTA owns ONE residue j and loops over its share of timelines.  It generates y_j straight into shared memory:
//     the residue twiddle is folded into the phase (u = q'.r - 4 n j / L quarter turns, one extra FMA), so y_j costs
//     the same 21 FP64 instructions per frame as a bare amplitude;
//   * N-point in-place radix-4 DIF FFT in shared memory (digit-reversed output order, which is irrelevant for a
//     power spectrum as long as the weights and the final inverse use the same order);
//   * |X|^2 accumulates in registers (N/256 doubles per thread) across all timelines of the CTA; the store() mean
//     a = mean_tau C[tau] is a dot product with precomputed weights What (see correlate.cu), reduced per timeline.
//   * one inverse transform per |q| (self_finalize) turns the summed power spectrum into sum C[tau].
// No atomics; partials are combined in a fixed order.
#include "kernels.hpp"
#include "sincos_qt.cuh"

#include <algorithm>
#include <vector>

namespace sass {

namespace {

constexpr int SF_THREADS = 256;
constexpr int SF_MAX_LOG2N = 12;  // N <= 4096 (64 KB of shared memory per CTA)

__device__ __forceinline__ double2 cmul2(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// In-place radix-4 DIF FFT of s[0..N) (natural order in, digit-reversed out).  tw[k] = exp(-2 pi i k / N), k < N.
// sign = -1 forward, +1 inverse (conjugated twiddles).  All SF_THREADS threads participate.
__device__ __forceinline__ void fft_dif_r4(double2 *s, int log2N, const double2 *__restrict__ tw, int sign) {
    const int N = 1 << log2N;
    int span_log = log2N;
    for (; span_log >= 2; span_log -= 2) {
        const int q = 1 << (span_log - 2);
        const int tstep = log2N - span_log;  // twiddle stride N/span = 2^tstep
        __syncthreads();
        for (int id = threadIdx.x; id < (N >> 2); id += SF_THREADS) {
            const int k = id & (q - 1);
            const int base = ((id >> (span_log - 2)) << span_log) + k;
            const double2 x0 = s[base], x1 = s[base + q], x2 = s[base + 2 * q], x3 = s[base + 3 * q];
            const double2 a = make_double2(x0.x + x2.x, x0.y + x2.y);
            const double2 b = make_double2(x0.x - x2.x, x0.y - x2.y);
            const double2 c = make_double2(x1.x + x3.x, x1.y + x3.y);
            const double2 dd = make_double2(x1.x - x3.x, x1.y - x3.y);
            // forward: (x1-x3) * (-i) = (im, -re); inverse: * (+i) = (-im, re)
            const double2 d = (sign < 0) ? make_double2(dd.y, -dd.x) : make_double2(-dd.y, dd.x);
            const double2 y0 = make_double2(a.x + c.x, a.y + c.y);
            double2 y1 = make_double2(b.x + d.x, b.y + d.y);
            double2 y2 = make_double2(a.x - c.x, a.y - c.y);
            double2 y3 = make_double2(b.x - d.x, b.y - d.y);
            if (k != 0) {
                double2 w1 = __ldg(&tw[(size_t)k << tstep]);
                double2 w2 = __ldg(&tw[(size_t)(2 * k) << tstep]);
                double2 w3 = __ldg(&tw[(size_t)(3 * k) << tstep]);
                if (sign > 0) {
                    w1.y = -w1.y;
                    w2.y = -w2.y;
                    w3.y = -w3.y;
                }
                y1 = cmul2(y1, w1);
                y2 = cmul2(y2, w2);
                y3 = cmul2(y3, w3);
            }
            s[base] = y0;
            s[base + q] = y1;
            s[base + 2 * q] = y2;
            s[base + 3 * q] = y3;
        }
    }
    if (span_log == 1) {  // odd log2N: final radix-2 stage on adjacent pairs
        __syncthreads();
        for (int id = threadIdx.x; id < (N >> 1); id += SF_THREADS) {
            const double2 x0 = s[2 * id], x1 = s[2 * id + 1];
            s[2 * id] = make_double2(x0.x + x1.x, x0.y + x1.y);
            s[2 * id + 1] = make_double2(x0.x - x1.x, x0.y - x1.y);
        }
    }
    __syncthreads();
}

enum { GEN_AMPLITUDE = 0, GEN_WEIGHTS = 1 };

// ---- register-resident radix-16 passes -------------------------------------------------------------------------------
// The N-point transform (N = 256 .. 4096) is done in at most three passes: radix 16, radix 16, radix N/256.  In a pass a
// thread loads the R inputs of one butterfly (stride q) into registers, runs an R-point DIF network with compile-time
// twiddles, applies the inter-pass twiddles W_S^{k c} (c = 1..R-1, built from one table load by a multiplication tree)
// and stores the results back in place.  Half the shared-memory traffic and barriers of radix-4 passes, and 16
// independent loads per thread instead of 4.  Shared-memory indices are padded by one element per 16 (phys()) so that
// the stride-1 pass, where a thread owns 16 consecutive elements, is bank-conflict free.
__device__ __forceinline__ int phys(int i) { return i + (i >> 4); }

__device__ __forceinline__ constexpr int bitrev_r(int i, int log2r) {
    int r = 0;
    for (int b = 0; b < log2r; b++) r |= ((i >> b) & 1) << (log2r - 1 - b);
    return r;
}

// cos/sin of 2 pi m / 16
__device__ constexpr double kCos16[16] = {1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977,
                                          0.0, -0.38268343236508977, -0.70710678118654752, -0.92387953251128674,
                                          -1.0, -0.92387953251128674, -0.70710678118654752, -0.38268343236508977,
                                          0.0, 0.38268343236508977, 0.70710678118654752, 0.92387953251128674};
__device__ constexpr double kSin16[16] = {0.0, 0.38268343236508977, 0.70710678118654752, 0.92387953251128674,
                                          1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977,
                                          0.0, -0.38268343236508977, -0.70710678118654752, -0.92387953251128674,
                                          -1.0, -0.92387953251128674, -0.70710678118654752, -0.38268343236508977};

// R-point DIF network in registers; on exit x[i] holds frequency bitrev(i).  sign = -1 forward, +1 inverse.
template <int R, int SIGN>
__device__ __forceinline__ void fft_regs(double2 (&x)[R]) {
#pragma unroll
    for (int len = R; len >= 2; len >>= 1) {
        const int half = len >> 1;
#pragma unroll
        for (int blk = 0; blk < R; blk += len) {
#pragma unroll
            for (int m = 0; m < half; m++) {
                const double2 a = x[blk + m], c = x[blk + m + half];
                x[blk + m] = make_double2(a.x + c.x, a.y + c.y);
                const double2 d = make_double2(a.x - c.x, a.y - c.y);
                const int e = m * (16 / len);  // twiddle W_len^m = W_16^e, e in [0, 8)
                if (e == 0) {
                    x[blk + m + half] = d;
                } else if (e == 4) {  // -i (forward) / +i (inverse)
                    x[blk + m + half] = (SIGN < 0) ? make_double2(d.y, -d.x) : make_double2(-d.y, d.x);
                } else {
                    const double wr = kCos16[e], wi = (SIGN < 0) ? -kSin16[e] : kSin16[e];
                    x[blk + m + half] = make_double2(fma(d.x, wr, -d.y * wi), fma(d.x, wi, d.y * wr));
                }
            }
        }
    }
}

__device__ __forceinline__ double2 csqr(double2 a) { return make_double2(fma(a.x, a.x, -a.y * a.y), 2.0 * a.x * a.y); }

// one DIF pass over blocks of span S = 2^LOG2S with radix R (q = S/R); N = 2^LOG2N
template <int LOG2N, int LOG2S, int R, int SIGN>
__device__ __forceinline__ void fft_pass(double2 *s, const double2 *__restrict__ tw) {
    constexpr int N = 1 << LOG2N, S = 1 << LOG2S, q = S / R, NB = N / R;
    constexpr int LOG2R = (R == 16) ? 4 : (R == 8) ? 3 : (R == 4) ? 2 : 1;
    for (int id = threadIdx.x; id < NB; id += SF_THREADS) {
        const int k = id & (q - 1);
        const int base = (id / q) * S + k;
        double2 x[R];
#pragma unroll
        for (int m = 0; m < R; m++) x[m] = s[phys(base + m * q)];
        fft_regs<R, SIGN>(x);
        if (q > 1) {
            double2 w[R];  // w[c] = W_S^{k c}
            w[1] = __ldg(&tw[k * (N / S)]);
            if (SIGN > 0) w[1].y = -w[1].y;
#pragma unroll
            for (int c = 2; c < R; c++) w[c] = (c & (c - 1)) == 0 ? csqr(w[c >> 1]) : cmul2(w[c & (c - 1)], w[c & -c]);
#pragma unroll
            for (int i = 1; i < R; i++) x[i] = cmul2(x[i], w[bitrev_r(i, LOG2R)]);
        }
#pragma unroll
        for (int i = 0; i < R; i++) s[phys(base + bitrev_r(i, LOG2R) * q)] = x[i];
    }
    __syncthreads();
}

// full transform: radices 16, 16, N/256 (N >= 256); entry assumes the buffer is complete (does a barrier first)
template <int LOG2N, int SIGN>
__device__ __forceinline__ void fft_r16(double2 *s, const double2 *__restrict__ tw) {
    __syncthreads();
    fft_pass<LOG2N, LOG2N, 16, SIGN>(s, tw);
    fft_pass<LOG2N, LOG2N - 4, 16, SIGN>(s, tw);
    if (LOG2N == 9) fft_pass<LOG2N, 1, 2, SIGN>(s, tw);
    if (LOG2N == 10) fft_pass<LOG2N, 2, 4, SIGN>(s, tw);
    if (LOG2N == 11) fft_pass<LOG2N, 3, 8, SIGN>(s, tw);
    if (LOG2N == 12) fft_pass<LOG2N, 4, 16, SIGN>(s, tw);
}

// grid = (R, G).  CTA (j, g) handles a contiguous share of the timelines tl < ntl, tl = atom_rel*NM + m.
//   GEN_AMPLITUDE: accumulate |X|^2 into Ppart[g][j][pos] and a_part[tl][j] = sum_pos |X|^2 * What[j][pos]
//   GEN_WEIGHTS  : single "timeline" w[tau] = 1/(NF-tau), inverse sign, store X to Wout[j][pos]
// dynamic shared memory: padded FFT buffer (N + N/16 complex) followed by the power accumulator (N doubles)
template <int LOG2N, int GEN>
__global__ void __launch_bounds__(SF_THREADS, 2) self_fused_kernel(
    const float *__restrict__ xyz, const double *__restrict__ b, const double *__restrict__ qs, int NF, int NM,
    size_t atom0, size_t ntl, int R, const double2 *__restrict__ tw, const double2 *__restrict__ What,
    double *__restrict__ Ppart, double2 *__restrict__ a_part, double2 *__restrict__ Wout) {
    extern __shared__ double2 s[];
    __shared__ double2 red[SF_THREADS / 32];
    constexpr int N = 1 << LOG2N;
    double *s_acc = reinterpret_cast<double *>(s + N + N / 16);
    const int j = blockIdx.x;
    const size_t g = blockIdx.y, G = gridDim.y;
    const double L = (double)R * (double)N;
    // residue twiddle in quarter turns per frame index: forward -4 j / L, inverse (weights) +4 j / L
    const double cj = ((GEN == GEN_AMPLITUDE) ? -4.0 : 4.0) * (double)j / L;

    if (GEN == GEN_AMPLITUDE)
        for (int pos = threadIdx.x; pos < N; pos += SF_THREADS) s_acc[pos] = 0.0;

    // contiguous share of the timelines: a CTA stays on one atom for up to NM consecutive jobs (coordinates cache-hot)
    const size_t per = (ntl + G - 1) / G;
    const size_t tl_begin = g * per, tl_end = min(ntl, tl_begin + per);
    for (size_t tl = tl_begin; tl < tl_end; tl++) {
        double qx = 0, qy = 0, qz = 0, bn = 1.0;
        const float *p = nullptr;
        if (GEN == GEN_AMPLITUDE) {
            const size_t atom = atom0 + tl / NM;
            const int m = (int)(tl % NM);
            qx = __ldg(&qs[3 * m]);
            qy = __ldg(&qs[3 * m + 1]);
            qz = __ldg(&qs[3 * m + 2]);
            bn = __ldg(&b[atom]);
            p = xyz + atom * (size_t)NF * 3;
        }
        __syncthreads();  // previous timeline's spectrum fully consumed
        // y_j[n mod N] += x[n] * exp(-+2 pi i n j / L).  n = tid + 256 i: all passes of one n' belong to the same
        // thread (256 | N), so the shared-memory accumulation is thread-private.  Flat loop, unrolled for ILP.
        if (NF < N)
            for (int np = NF + threadIdx.x; np < N; np += SF_THREADS) s[phys(np)] = make_double2(0.0, 0.0);
#pragma unroll 4
        for (int n = threadIdx.x; n < NF; n += SF_THREADS) {
            double u, amp;
            if (GEN == GEN_AMPLITUDE) {
                const double x = (double)__ldg(&p[3 * n]), y = (double)__ldg(&p[3 * n + 1]), z = (double)__ldg(&p[3 * n + 2]);
                u = fma(z, qz, fma(y, qy, fma(x, qx, (double)n * cj)));
                amp = bn;
            } else {
                u = (double)n * cj;
                amp = 1.0 / (double)(NF - n);
            }
            double sn, cs;
            sincos_qt(u, sn, cs);
            const int np = phys(n & (N - 1));
            double2 v = make_double2(amp * cs, amp * sn);
            if (n >= N) {
                const double2 o = s[np];
                v.x += o.x;
                v.y += o.y;
            }
            s[np] = v;
        }
        fft_r16<LOG2N, (GEN == GEN_AMPLITUDE) ? -1 : +1>(s, tw);
        if (GEN == GEN_AMPLITUDE) {
            double2 ap = make_double2(0.0, 0.0);
#pragma unroll 4
            for (int pos = threadIdx.x; pos < N; pos += SF_THREADS) {
                const double2 v = s[phys(pos)];
                const double pw = fma(v.x, v.x, v.y * v.y);
                s_acc[pos] += pw;
                const double2 w = __ldg(&What[(size_t)j * N + pos]);
                ap.x = fma(pw, w.x, ap.x);
                ap.y = fma(pw, w.y, ap.y);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ap.x += __shfl_xor_sync(0xffffffffu, ap.x, o);
                ap.y += __shfl_xor_sync(0xffffffffu, ap.y, o);
            }
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ap;
            __syncthreads();
            if (threadIdx.x == 0) {
                double2 t = red[0];
#pragma unroll
                for (int w = 1; w < SF_THREADS / 32; w++) {
                    t.x += red[w].x;
                    t.y += red[w].y;
                }
                a_part[tl * R + j] = t;
            }
        } else {
            for (int pos = threadIdx.x; pos < N; pos += SF_THREADS) Wout[(size_t)j * N + pos] = s[phys(pos)];
        }
    }
    if (GEN == GEN_AMPLITUDE) {
        __syncthreads();
        for (int pos = threadIdx.x; pos < N; pos += SF_THREADS) Ppart[(g * R + j) * (size_t)N + pos] = s_acc[pos];
    }
}

// P[i] += sum_g Ppart[g][i]
__global__ void sf_reduce_ppart_kernel(const double *__restrict__ Ppart, size_t G, size_t len, double *__restrict__ P) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= len) return;
    double sum = 0.0;
    for (size_t g = 0; g < G; g++) sum += Ppart[g * len + i];
    P[i] += sum;
}

// a_tl[tl] = norm * sum_j a_part[tl][j]
__global__ void sf_reduce_apart_kernel(const double2 *__restrict__ a_part, size_t ntl, int R, double norm,
                                       double2 *__restrict__ a_tl) {
    size_t tl = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (tl >= ntl) return;
    double2 sum = make_double2(0.0, 0.0);
    for (int j = 0; j < R; j++) {
        const double2 v = a_part[tl * R + j];
        sum.x += v.x;
        sum.y += v.y;
    }
    a_tl[tl] = make_double2(sum.x * norm, sum.y * norm);
}

// acc[0..2] += {sum re a, sum im a, sum |a|^2} — single CTA, fixed order
__global__ void __launch_bounds__(1024) sf_reduce_atl_kernel(const double2 *__restrict__ a_tl, size_t n,
                                                             double *__restrict__ acc) {
    __shared__ double sh[3][32];
    double r = 0.0, i = 0.0, q = 0.0;
    for (size_t k = threadIdx.x; k < n; k += blockDim.x) {
        const double2 a = a_tl[k];
        r += a.x;
        i += a.y;
        q += fma(a.x, a.x, a.y * a.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o);
        i += __shfl_xor_sync(0xffffffffu, i, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = r;
        sh[1][threadIdx.x >> 5] = i;
        sh[2][threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double Rr = 0, I = 0, Q = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            Rr += sh[0][w];
            I += sh[1][w];
            Q += sh[2][w];
        }
        acc[0] += Rr;
        acc[1] += I;
        acc[2] += Q;
    }
}

// finalize step 1: per residue j, inverse N-FFT of P_j (stored by digit-reversed position) -> Q_j[t], t < N, natural
// (P is stored in the radix-16 kernel's position order, freq16; the radix-4 transform used here has its own, freq4)
__global__ void __launch_bounds__(SF_THREADS) sf_inv_residue_kernel(const double *__restrict__ P, int log2N,
                                                                    const int *__restrict__ freq16,
                                                                    const int *__restrict__ freq4,
                                                                    const double2 *__restrict__ tw,
                                                                    double2 *__restrict__ Q) {
    extern __shared__ double2 s[];
    const int N = 1 << log2N;
    const int j = blockIdx.x;
    for (int pos = threadIdx.x; pos < N; pos += SF_THREADS)
        s[freq16[pos]] = make_double2(P[(size_t)j * N + pos], 0.0);  // natural frequency order for the DIF input
    fft_dif_r4(s, log2N, tw, +1);
    for (int pos = threadIdx.x; pos < N; pos += SF_THREADS) Q[(size_t)j * N + freq4[pos]] = s[pos];
}

// finalize step 2: c[tau] = sum_j exp(+2 pi i j tau / L) Q_j[tau mod N]; out = scale * c / (L (NF - tau))
__global__ void sf_combine_kernel(const double2 *__restrict__ Q, int log2N, int R, size_t NF, double scale, int conj_out,
                                  double2 *__restrict__ out) {
    const size_t tau = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (tau >= NF) return;
    const size_t N = (size_t)1 << log2N;
    const size_t L = (size_t)R * N;
    const size_t t = tau & (N - 1);
    double cr = 0.0, ci = 0.0;
    for (int j = 0; j < R; j++) {
        const double2 v = Q[(size_t)j * N + t];
        double sn, cs;
        sincospi(2.0 * (double)(((size_t)j * tau) % L) / (double)L, &sn, &cs);
        cr += v.x * cs - v.y * sn;
        ci += v.x * sn + v.y * cs;
    }
    const double f = scale / ((double)L * (double)(NF - tau));
    out[tau] = make_double2(cr * f, conj_out ? -ci * f : ci * f);
}

__global__ void sf_twiddle_kernel(double2 *tw, size_t N) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= N) return;
    double sn, cs;
    sincospi(-2.0 * (double)k / (double)N, &sn, &cs);
    tw[k] = make_double2(cs, sn);
}

int freq_of_pos_host(int pos, int log2N) {
    // mirrors fft_dif_r4: radix-4 stages while span >= 4, then an optional radix-2 stage
    if (log2N == 0) return 0;
    if (log2N == 1) return pos;
    const int quarter = pos >> (log2N - 2);
    const int rest = pos & ((1 << (log2N - 2)) - 1);
    return 4 * freq_of_pos_host(rest, log2N - 2) + quarter;
}

// position -> frequency of the radix-16 pass sequence (16, 16, N/256)
int freq16_of_pos_host(int pos, int log2N) {
    int radices[3] = {16, 16, 1 << (log2N - 8)};
    int span = 1 << log2N;
    int freq = 0, mult = 1;
    for (int pass = 0; pass < 3; pass++) {
        const int r = radices[pass];
        if (r == 1) break;
        const int q = span / r;
        const int c = pos / q;
        pos = pos % q;
        freq += mult * c;
        mult *= r;
        span = q;
    }
    return freq;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t sf_smem_bytes(int log2N) {
    const size_t N = (size_t)1 << log2N;
    return (N + N / 16) * sizeof(double2) + N * sizeof(double);
}

template <int GEN>
void launch_fused(int log2N, dim3 grid, cudaStream_t st, const float *xyz, const double *b, const double *qs, int NF, int NM,
                  size_t atom0, size_t ntl, int R, const double2 *tw, const double2 *What, double *Ppart,
                  double2 *a_part, double2 *Wout) {
    const size_t smem = sf_smem_bytes(log2N);
#define SF_CASE(LN)                                                                                                    \
    case LN: {                                                                                                         \
        static bool attr = false;                                                                                      \
        if (!attr) {                                                                                                   \
            cudaFuncSetAttribute(self_fused_kernel<LN, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
            attr = true;                                                                                               \
        }                                                                                                              \
        self_fused_kernel<LN, GEN><<<grid, SF_THREADS, smem, st>>>(xyz, b, qs, NF, NM, atom0, ntl, R, tw, What, Ppart, \
                                                                   a_part, Wout);                                     \
        break;                                                                                                         \
    }
    switch (log2N) {
        SF_CASE(8)
        SF_CASE(9)
        SF_CASE(10)
        SF_CASE(11)
        SF_CASE(12)
        default: break;
    }
#undef SF_CASE
}

size_t pick_groups(const SelfPlan *p, size_t ntl) {
    size_t G = (2 * 148 + p->R - 1) / p->R;  // two resident CTAs per SM
    if (G > ntl) G = ntl;
    if (G < 1) G = 1;
    if (G > 65535) G = 65535;
    return G;
}

}  // namespace

int self_plan_create(SelfPlan *p, size_t NF, cudaStream_t st, uint64_t *launches) {
    if (NF == 0) return 1;
    p->NF = NF;
    const size_t need = 2 * NF - 1;
    int log2N = 8;  // N >= 256 = SF_THREADS: the generation loop relies on 256 | N
    while (log2N < SF_MAX_LOG2N && ((size_t)1 << log2N) < need) log2N++;
    p->log2N = log2N;
    p->N = (size_t)1 << log2N;
    p->R = (int)((need + p->N - 1) / p->N);
    if (p->R > 4096) return 1;  // NF > ~8e6 frames
    p->L = (size_t)p->R * p->N;
    if (cudaMalloc(&p->d_tw, sizeof(double2) * p->N) != cudaSuccess) return 2;
    if (cudaMalloc(&p->d_w, sizeof(double2) * p->L) != cudaSuccess) return 2;
    if (cudaMalloc(&p->d_freq, sizeof(int) * 2 * p->N) != cudaSuccess) return 2;
    std::vector<int> f(2 * p->N);
    for (size_t i = 0; i < p->N; i++) {
        f[i] = freq16_of_pos_host((int)i, log2N);       // order of the fused forward kernel
        f[p->N + i] = freq_of_pos_host((int)i, log2N);  // order of the radix-4 transform used at finalize
    }
    cudaMemcpyAsync(p->d_freq, f.data(), sizeof(int) * 2 * p->N, cudaMemcpyHostToDevice, st);
    sf_twiddle_kernel<<<(unsigned)((p->N + 255) / 256), 256, 0, st>>>(p->d_tw, p->N);
    cudaFuncSetAttribute(sf_inv_residue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    launch_fused<GEN_WEIGHTS>(log2N, dim3(p->R, 1), st, nullptr, nullptr, nullptr, (int)NF, 1, 0, 1, p->R, p->d_tw, nullptr,
                              nullptr, nullptr, p->d_w);
    cudaStreamSynchronize(st);
    if (launches) *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

void self_plan_destroy(SelfPlan *p) {
    if (p->d_tw) cudaFree(p->d_tw);
    if (p->d_w) cudaFree(p->d_w);
    if (p->d_freq) cudaFree(p->d_freq);
    p->d_tw = nullptr;
    p->d_w = nullptr;
    p->d_freq = nullptr;
    p->NF = p->L = 0;
}

size_t self_work_bytes(const SelfPlan *p, size_t ntl) {
    const size_t G = pick_groups(p, ntl);
    return align256(G * p->L * sizeof(double)) + align256(ntl * p->R * sizeof(double2)) + align256(ntl * sizeof(double2)) +
           align256(p->L * sizeof(double2));
}

int self_power_accumulate(const SelfPlan *p, const float *d_xyz_by_atom, const double *d_b, const double *d_qs,
                          size_t NM, size_t atom0, size_t natoms, void *d_work, double *d_P, double *d_acc,
                          cudaStream_t st) {
    const size_t ntl = natoms * NM;
    if (ntl == 0) return 0;
    const size_t G = pick_groups(p, ntl);
    char *w = reinterpret_cast<char *>(d_work);
    double *Ppart = reinterpret_cast<double *>(w);
    w += align256(G * p->L * sizeof(double));
    double2 *a_part = reinterpret_cast<double2 *>(w);
    w += align256(ntl * p->R * sizeof(double2));
    double2 *a_tl = reinterpret_cast<double2 *>(w);
    launch_fused<GEN_AMPLITUDE>(p->log2N, dim3(p->R, (unsigned)G), st, d_xyz_by_atom, d_b, d_qs, (int)p->NF, (int)NM, atom0,
                                ntl, p->R, p->d_tw, p->d_w, Ppart, a_part, nullptr);
    sf_reduce_ppart_kernel<<<(unsigned)((p->L + 255) / 256), 256, 0, st>>>(Ppart, G, p->L, d_P);
    const double norm = 1.0 / ((double)p->NF * (double)p->L);
    sf_reduce_apart_kernel<<<(unsigned)((ntl + 255) / 256), 256, 0, st>>>(a_part, ntl, p->R, norm, a_tl);
    sf_reduce_atl_kernel<<<1, 1024, 0, st>>>(a_tl, ntl, d_acc);
    return 4;
}

int self_finalize(const SelfPlan *p, const double *d_P, void *d_work, double2 *d_out, double scale, int conj_out,
                  cudaStream_t st) {
    double2 *Q = reinterpret_cast<double2 *>(d_work);
    sf_inv_residue_kernel<<<p->R, SF_THREADS, sizeof(double2) * p->N, st>>>(d_P, p->log2N, p->d_freq, p->d_freq + p->N,
                                                                            p->d_tw, Q);
    sf_combine_kernel<<<(unsigned)((p->NF + 127) / 128), 128, 0, st>>>(Q, p->log2N, p->R, p->NF, scale, conj_out, d_out);
    return 2;
}

}  // namespace sass
