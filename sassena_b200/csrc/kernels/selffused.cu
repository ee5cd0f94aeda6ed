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
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace sass {

namespace {

constexpr int SF_THREADS = 256;
constexpr int SF_MAX_LOG2N = 12;  // N <= 4096 (64 KB of shared memory per CTA)
constexpr int SF_2S_THREADS = 64;  // positions per CTA of the two-stage combine kernel (see there)

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

enum { GEN_AMPLITUDE = 0, GEN_WEIGHTS = 1, GEN_LOAD = 2 };

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

// cos/sin of 2 pi m / 16.  In constant memory: the butterflies take them as c[bank][offset] operands; as constexpr
// literals ptxas materialises them in registers outside the loops (spilled in the split kernel: 30 reloads per transform)
__constant__ double kCos16[16] = {1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977,
                                          0.0, -0.38268343236508977, -0.70710678118654752, -0.92387953251128674,
                                          -1.0, -0.92387953251128674, -0.70710678118654752, -0.38268343236508977,
                                          0.0, 0.38268343236508977, 0.70710678118654752, 0.92387953251128674};
__constant__ double kSin16[16] = {0.0, 0.38268343236508977, 0.70710678118654752, 0.92387953251128674,
                                          1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977,
                                          0.0, -0.38268343236508977, -0.70710678118654752, -0.92387953251128674,
                                          -1.0, -0.92387953251128674, -0.70710678118654752, -0.38268343236508977};

// R-point DIF network in registers; on exit x[i] holds frequency bitrev(i).  sign = -1 forward, +1 inverse.
template <int R, int SIGN, int LEN0 = R>
__device__ __forceinline__ void fft_regs(double2 (&x)[R]) {
#pragma unroll
    for (int len = LEN0; len >= 2; len >>= 1) {
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
struct GlobalTwiddles {  // exp(-2 pi i t / N) from an N-entry table in global memory
    const double2 *__restrict__ tw;
    __device__ __forceinline__ double2 operator()(int t) const { return __ldg(&tw[t]); }
};
template <int LOG2N, int LOG2S, int R, int SIGN, class TW>
__device__ __forceinline__ void fft_pass(double2 *s, const TW &tw) {
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
            w[1] = tw(k * (N / S));
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
template <int LOG2N, int SIGN, class TW>
__device__ __forceinline__ void fft_r16(double2 *s, const TW &tw) {
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
//   GEN_LOAD     : like GEN_AMPLITUDE, but timeline tl is read from Ain[tl * ldA + n], n < NF (R == 1: no residue twiddle)
// dynamic shared memory: padded FFT buffer (N + N/16 complex) followed by the power accumulator (N doubles)
template <int LOG2N, int GEN>
__global__ void __launch_bounds__(SF_THREADS, 2) self_fused_kernel(
    const float *__restrict__ xyz, const double *__restrict__ b, const double *__restrict__ qs, int NF, int NM,
    size_t atom0, size_t ntl, int R, const double2 *__restrict__ tw, const double2 *__restrict__ What,
    double *__restrict__ Ppart, double2 *__restrict__ a_part, double2 *__restrict__ Wout, const double2 *__restrict__ Ain,
    size_t ldA) {
    constexpr bool ACC = (GEN != GEN_WEIGHTS);  // accumulate power spectra and the store() means
    extern __shared__ double2 s[];
    __shared__ double2 red[SF_THREADS / 32];
    constexpr int N = 1 << LOG2N;
    double *s_acc = reinterpret_cast<double *>(s + N + N / 16);
    const int j = blockIdx.x;
    const size_t g = blockIdx.y, G = gridDim.y;
    const double L = (double)R * (double)N;
    // residue twiddle in quarter turns per frame index: forward -4 j / L, inverse (weights) +4 j / L
    const double cj = ((GEN != GEN_WEIGHTS) ? -4.0 : 4.0) * (double)j / L;

    if (ACC)
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
        if (GEN == GEN_LOAD) {
            const double2 *src = Ain + tl * ldA;
#pragma unroll 4
            for (int n = threadIdx.x; n < NF; n += SF_THREADS) s[phys(n)] = __ldg(&src[n]);
        }
#pragma unroll 4
        for (int n = threadIdx.x; n < (GEN == GEN_LOAD ? 0 : NF); n += SF_THREADS) {
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
        fft_r16<LOG2N, ACC ? -1 : +1>(s, GlobalTwiddles{tw});
        if (ACC) {
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
    if (ACC) {
        __syncthreads();
        for (int pos = threadIdx.x; pos < N; pos += SF_THREADS) Ppart[(g * R + j) * (size_t)N + pos] = s_acc[pos];
    }
}

// ---- split path ---------------------------------------------------------------------------------------------------
// The fused kernel above evaluates exp(i q.r) R times per frame (once per residue).  For R >= 5 the transform is split the
// other way round: time index n = R m + r, frequency k = k1 + N k2,
//     X[k1 + N k2] = sum_r exp(-2 pi i r k2 / R) * ( exp(-2 pi i r k1 / L) * Z_r[k1] ),   Z_r = FFT_N( x[R m + r] )_m
// so every frame is evaluated ONCE (it belongs to exactly one decimated sequence), the R sub-FFTs of a timeline run back to
// back in one CTA (kernel A, coordinates stay L1/L2-hot), the twiddled Z_r go to a batch buffer in HBM/L2 ([tl][r][pos],
// coalesced on both sides), and kernel B does the R-point DFT across r for a slice of positions, accumulates |X|^2 per
// (k2, pos) in shared memory over all its timelines and the store() mean as the weighted sum with What (split layout).
// The result is permuted back to the residue-major layout at the reduction, so finalize is unchanged.

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- split path, kernel A ----------------------------------------------------------------------------------------------
// N = 4096 = 16^3 (the split path only runs with R >= 2, i.e. 2NF-1 > 4096).  Index algebra: n = 256 n1 + 16 n2 + n3,
// k = k1 + 16 k2 + 256 k3,
//     X[k] = sum_n3 W_16^{n3 k3} W_256^{n3 k2} [ sum_n2 W_16^{n2 k2} W_N^{k1 (16 n2 + n3)} [ sum_n1 W_16^{n1 k1} x[n] ] ].
//   pass 1  thread (n2, n3) = tid: the amplitudes x[256 n1 + tid], n1 < 8, are evaluated straight into registers (a
//           decimated sub-sequence has at most N/2 frames: n1 >= 8 is the zero pad, so the first DIF stage of the 16-point
//           network degenerates to 8 twiddle products), twiddled by W_N^{k1 tid} and written to shared memory;
//   pass 2  thread (k1, n3): 16-point transform over n2 in place, twiddled by W_256^{n3 k2};
//   pass 3  thread (k1, k2) = tid: 16-point transform over n3, times the split twiddle W_L^{r k} (and the atom's factor),
//           stored straight to Z in NATURAL frequency order: Z[tid + 256 k3] is coalesced.
// Shared-memory layout  (k1 ^ (n3 & 7)) + 16 n3 + 256 (n2 | k2): every access of the three passes is conflict free
// without padding.  Pass 3 of one sub-transform and pass 1 of the next touch the same 512 elements per warp and every warp
// fetches its own coordinates, so a sub-transform costs two CTA barriers (after pass 1 and after pass 2).  All twiddle powers w^k, k < 16, come from two interleaved
// three-term recurrences w_{k+2} = 2 cos(2 theta) w_k - w_{k-2} (one FMA per component; error <= 1.2e-14, checked in
// tools/check_twiddle_recurrence.py) seeded by thread constants, instead of table loads and complex products.
__device__ __forceinline__ constexpr int bitrev4(int i) {
    return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);
}

// emit(k, x[bitrev4(k)] * z_k), k = 0..15, where z_k = z0 * g^k, |g| = 1, given z0, z1 = z0 * g and cg2 = 2 Re g.
// UNIT0: z0 == 1.  Every product is handed to `emit` (a store) as soon as it exists, so the stores of a pass are spread
// over the recurrence instead of queueing up behind it.
template <bool UNIT0, class Emit>
__device__ __forceinline__ void twiddle16(const double2 (&x)[16], double2 z0, double2 z1, double cg2, Emit &&emit) {
    double2 pa = z0, pb = z1;
    double2 wa = make_double2(fma(cg2, z1.x, -z0.x), fma(cg2, z1.y, -z0.y));  // z2
    double2 wb = make_double2(fma(cg2, wa.x, -z1.x), fma(cg2, wa.y, -z1.y));  // z3
    const double C2 = fma(cg2, cg2, -2.0);                                    // 2 cos(2 theta)
    emit(0, UNIT0 ? x[0] : cmul2(x[0], z0));
    emit(1, cmul2(x[bitrev4(1)], z1));
    emit(2, cmul2(x[bitrev4(2)], wa));
    emit(3, cmul2(x[bitrev4(3)], wb));
#pragma unroll
    for (int k = 4; k < 16; k += 2) {
        const double2 na = make_double2(fma(C2, wa.x, -pa.x), fma(C2, wa.y, -pa.y));
        pa = wa;
        wa = na;
        emit(k, cmul2(x[bitrev4(k)], wa));
        const double2 nb = make_double2(fma(C2, wb.x, -pb.x), fma(C2, wb.y, -pb.y));
        pb = wb;
        wb = nb;
        emit(k + 1, cmul2(x[bitrev4(k + 1)], wb));
    }
}

// 16-point forward DIF network whose inputs 8..15 are zero: the first stage is x[m + 8] = x[m] W_16^m
__device__ __forceinline__ void fft16_upper_zero(double2 (&x)[16]) {
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const double2 d = x[m];
        if (m == 0) {
            x[8] = d;
        } else if (m == 4) {
            x[12] = make_double2(d.y, -d.x);
        } else {
            const double wr = kCos16[m], wi = -kSin16[m];
            x[m + 8] = make_double2(fma(d.x, wr, -d.y * wi), fma(d.x, wi, d.y * wr));
        }
    }
    fft_regs<16, -1, 8>(x);
}

__global__ void __launch_bounds__(SF_THREADS, 2) self_split_fft12_kernel(
    const float *__restrict__ xyz, const double *__restrict__ b, const double *__restrict__ qs, int NF, int NM, size_t atom0,
    size_t tl_first, size_t ntl, int R, int dec, double2 *__restrict__ Zt) {
    extern __shared__ double2 s[];
    constexpr int N = 4096;
    // exp(-2 pi i t / L) = Thi[t >> 8] * Tlo[t & 255]
    double2 *Tlo = s + N;
    double2 *Thi = Tlo + 256;
    const int L = R * N;
    double *qb = reinterpret_cast<double *>(Thi + (L >> 8));  // (q', b) of the current and the next timeline, 2 x 4 doubles
    float *cbuf = reinterpret_cast<float *>(qb + 8);           // coordinates of one sub-sequence, [m][3]
    const int tid = threadIdx.x;
    for (int i = tid; i < 256 + (L >> 8); i += SF_THREADS) {
        const int t = (i < 256) ? i : ((i - 256) << 8);
        double sn, cs;
        sincospi(-2.0 * (double)t / (double)L, &sn, &cs);
        Tlo[i] = make_double2(cs, sn);  // Thi follows Tlo
    }
    // thread constants: W_N^{tid} (pass 1) and W_256^{tid >> 4} (pass 2)
    double2 w1, w2;
    sincospi(-2.0 * (double)tid / 4096.0, &w1.y, &w1.x);
    sincospi(-2.0 * (double)(tid >> 4) / 256.0, &w2.y, &w2.x);
    const int lo4 = tid & 15, hi4 = tid >> 4;
    const int i1 = (16 * tid) | (tid & 7);            // pass 1 stores s[i1 ^ k1]
    const int i2 = (lo4 ^ (hi4 & 7)) + 16 * hi4;      // pass 2: s[i2 + 256 n2], in place
    const int i3 = 256 * hi4;                         // pass 3: s[i3 + 16 n3 + (k1 ^ (n3 & 7))]

    const int base = NF / R, rem = NF % R;
    // the (timeline, r) sub-transforms of the launch are dealt out evenly, CTA g takes [g P / G, (g + 1) P / G): whole
    // timelines per CTA cost up to 26 % at R = 25 (3.16 timelines per CTA means 4 for some)
    const size_t P = ntl * (size_t)R;
    const size_t p_begin = (size_t)(((unsigned __int128)blockIdx.x * P) / gridDim.x);
    const size_t p_end = (size_t)(((unsigned __int128)(blockIdx.x + 1) * P) / gridDim.x);
    if (p_begin >= p_end) return;
    // The sub-sequence lands at cbuf + mis, mis = its misalignment (in floats) against 16 bytes in global memory, so that
    // source and destination are congruent and the body moves as 16-byte cp.async; returns mis (0 for the gather).
    // Every WARP fetches exactly the frames it evaluates in pass 1 (rows 256 n1 + 32 warp .. + 32, i.e. the 16-byte chunks
    // 192 n1 + 24 warp + lane, lane < 24, plus one more when the rows straddle chunk boundaries), so the arrival of the
    // coordinates is a warp-local matter (cp.async.wait_group + __syncwarp) and needs no CTA barrier.  (Reads of pass 1
    // beyond the sub-sequence's last frame are redirected to a frame of the warp's own first row; their values are discarded.)
    const int lane = tid & 31;
    const int gl = 96 * (tid >> 5) + 4 * lane;
    auto prefetch = [&](const float *p, int r) -> int {
        const int Mr = base + (r < rem ? 1 : 0);
        int mis = 0;
        if (dec) {
            const float *pr = p + 3 * (size_t)(r * base + min(r, rem));
            mis = (int)((reinterpret_cast<size_t>(pr) >> 2) & 3);
            const float *pa = pr - mis;      // 16-byte aligned: shifted index g is pa[g] in global memory and cbuf[g] here
            const int g_end = mis + 3 * Mr;  // valid shifted indices: [mis, g_end)
            if (lane < 24 + (mis != 0 ? 1 : 0)) {
#pragma unroll
                for (int n1 = 0; n1 < 8; n1++) {
                    const int gi = gl + 768 * n1;
                    // this warp's floats of the row: shifted indices [lo, hi); a chunk that straddles lo or hi (the first and
                    // the last of the row when the sub-sequence is not 16-byte aligned) is copied float by float, its other
                    // floats belong to the neighbouring warp, which copies them itself: no byte is written twice
                    const int lo = gi - 4 * lane + mis, hi = min(lo + 96, g_end);
                    if (gi >= lo && gi + 4 <= hi) {
                        cp_async16(&cbuf[gi], &pa[gi]);
                    } else if (gi < hi) {
#pragma unroll 1
                        for (int i = 0; i < 4; i++)
                            if (gi + i >= lo && gi + i < hi) cp_async4(&cbuf[gi + i], &pa[gi + i]);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int n1 = 0; n1 < 8; n1++) {
                const int f = 256 * n1 + tid;
                if (f < Mr) {
                    const float *src = p + 3 * ((size_t)R * f + r);
                    cp_async4(&cbuf[3 * f], src);
                    cp_async4(&cbuf[3 * f + 1], src + 1);
                    cp_async4(&cbuf[3 * f + 2], src + 2);
                }
            }
        }
        return mis;
    };
    // (q', b) of timeline (arel, m) into slot `slot` of qb: threads 0..3 (warp 0), 8 bytes each, in the current cp.async group
    auto prefetch_qb = [&](size_t arel_, int m_, int slot) {
        if (tid < 4) {
            const double *src = (tid < 3) ? (qs + 3 * m_ + tid) : (b + atom0 + arel_);
            const unsigned d = (unsigned)__cvta_generic_to_shared(qb + 4 * slot + tid);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
        }
    };
    // (atom, q-vector) of the CTA's first timeline; advanced incrementally afterwards (no division inside the loop)
    const size_t t_first = p_begin / (size_t)R;
    int r = (int)(p_begin - t_first * (size_t)R);
    size_t arel = (tl_first + t_first) / (size_t)NM;
    int m = (int)((tl_first + t_first) - arel * (size_t)NM);
    const float *p = xyz + (atom0 + arel) * (size_t)NF * 3;
    prefetch_qb(arel, m, 0);
    int mis_next = prefetch(p, r);
    cp_async_commit_group();
    cp_async_wait_all();
    __syncthreads();  // tables and the first timeline's (q', b) are visible to every warp
    double2 *out = Zt + (p_begin << 12);
    int slot = 0;
    double qx = 0.0, qy = 0.0, qz = 0.0, bn = 0.0;
    for (size_t pair = p_begin; pair < p_end; pair++, out += N) {
        if (pair == p_begin || r == 0) {  // the CTA's first pair of a timeline
            if (pair != p_begin + 1) {    // staged: by the blocking fetch above, or two pairs ago (see below)
                qx = qb[4 * slot], qy = qb[4 * slot + 1], qz = qb[4 * slot + 2], bn = qb[4 * slot + 3];
            } else {  // the CTA's share began with the last sub-transform of the previous timeline: nothing was staged yet
                qx = __ldg(&qs[3 * m]), qy = __ldg(&qs[3 * m + 1]), qz = __ldg(&qs[3 * m + 2]), bn = __ldg(&b[atom0 + arel]);
            }
        }
        const float *cb = cbuf + mis_next;
        const int Mr = base + (r < rem ? 1 : 0);
        double2 x[16];
        // the twiddle powers below are rebuilt per sub-transform on purpose (2 x 28 FMAs): hoisted out of the loop they
        // are 60 live doubles per thread, i.e. spills whose reloads miss the L1 left over beside 2 x 94 KB of shared memory
        asm volatile("" : "+d"(w1.x), "+d"(w1.y), "+d"(w2.x), "+d"(w2.y));
        // this warp's coordinates are in cbuf (it fetched them itself), and pass 3 of the previous pair has read the
        // part of the FFT buffer pass 1 is about to overwrite: both read and written by this warp only (elements
        // [512 warp, 512 warp + 512)), so a warp-level synchronisation is all that is needed here
        cp_async_wait_all();
        __syncwarp();
        // ---- pass 1: amplitudes of frames tid + 256 n1 in registers, pruned 16-point transform over n1.  Branch free
        // (clamped index + select) so that the eight evaluations interleave; rows beyond the sub-sequence are skipped
        // by CTA-uniform tests
        const int rows = (Mr + 255) >> 8;
        if (rows == 8) {
#pragma unroll
            for (int n1 = 0; n1 < 8; n1++) {
                const int mm = tid + 256 * n1, mc = mm < Mr ? mm : tid;  // dummy: a frame of this warp's own (Mr >= 1024)
                const double cx = (double)cb[3 * mc], cy = (double)cb[3 * mc + 1], cz = (double)cb[3 * mc + 2];
                double2 v;
                sincos_qt(fma(cz, qz, fma(cy, qy, cx * qx)), v.y, v.x);
                x[n1] = (n1 < 7 || mm < Mr) ? v : make_double2(0.0, 0.0);
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 8; n1++) {
                const int mm = tid + 256 * n1;
                double2 v = make_double2(0.0, 0.0);
                if (n1 < rows) {
                    const int mc = mm < Mr ? mm : tid;
                    const double cx = (double)cb[3 * mc], cy = (double)cb[3 * mc + 1], cz = (double)cb[3 * mc + 2];
                    sincos_qt(fma(cz, qz, fma(cy, qy, cx * qx)), v.y, v.x);
                    if (mm >= Mr) v = make_double2(0.0, 0.0);
                }
                x[n1] = v;
            }
        }
        fft16_upper_zero(x);
        twiddle16<true>(x, make_double2(1.0, 0.0), w1, w1.x + w1.x, [&](int k1, double2 v) { s[i1 ^ k1] = v; });
        __syncthreads();  // cbuf consumed by every warp, pass 1 complete
        const int r_this = r;
        if (pair + 1 < p_end) {  // the next pair's coordinates land while passes 2 and 3 run
            if (++r == R) {      // ... it opens the next timeline
                r = 0;
                slot ^= 1;
                if (++m == NM) {
                    m = 0;
                    arel++;
                }
                p = xyz + (atom0 + arel) * (size_t)NF * 3;
            } else if (r + 1 == R && pair + 2 < p_end) {
                // this timeline's last pair comes next and the CTA goes on to the timeline after it: that one's (q', b)
                // join this cp.async group, are waited for by warp 0 at the top of the next pair and published to the
                // other warps by that pair's barrier above
                const bool wrap = (m + 1 == NM);
                prefetch_qb(arel + (wrap ? 1 : 0), wrap ? 0 : m + 1, slot ^ 1);
            }
            mis_next = prefetch(p, r);
        }
        cp_async_commit_group();
        // ---- pass 2: over n2, in place
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) x[n2] = s[i2 + 256 * n2];
        fft_regs<16, -1>(x);
        twiddle16<true>(x, make_double2(1.0, 0.0), w2, w2.x + w2.x, [&](int k2, double2 v) { s[i2 + 256 * k2] = v; });
        __syncthreads();
        // ---- pass 3: over n3; split twiddle b * W_L^{r (tid + 256 k3)} = h * g^k3, g = W_L^{256 r} = Thi[r]
#pragma unroll
        for (int n3 = 0; n3 < 16; n3++) x[n3] = s[i3 + 16 * n3 + (lo4 ^ (n3 & 7))];
        fft_regs<16, -1>(x);
        {
            const int th = r_this * tid;  // < 64 * 256 <= L
            double2 h = cmul2(Thi[th >> 8], Tlo[th & 255]);
            h.x *= bn;
            h.y *= bn;
            const double2 gg = Thi[r_this];
            twiddle16<false>(x, h, cmul2(h, gg), gg.x + gg.x, [&](int k3, double2 v) { out[tid + 256 * k3] = v; });
        }
    }
}

// natural [n] <-> decimated [r][m] order of one atom's frames, out of place: dst[atom][...] = src[atom][...]
__global__ void sf_decimate_kernel(const float *__restrict__ src, float *__restrict__ dst, size_t natoms, int NF, int R,
                                   int forward) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= natoms * (size_t)NF) return;
    const size_t atom = i / NF;
    const int n = (int)(i - atom * NF);
    const int base = NF / R, rem = NF % R;
    const int r = n % R, m = n / R;
    const size_t d = atom * (size_t)NF + (size_t)(r * base + min(r, rem)) + m;
    const size_t from = forward ? i : d, to = forward ? d : i;
    dst[3 * to] = src[3 * from];
    dst[3 * to + 1] = src[3 * from + 1];
    dst[3 * to + 2] = src[3 * from + 2];
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// grid = (C, G).  CTA (c, g): positions [c*S, (c+1)*S) of the timelines in its share of the batch.
// dynamic shared memory: Zs[2][R*S] double2 | Wh[R*S] double2 | acc[R*S] double | Wr[R] double2 | red[2][8] double2
__global__ void __launch_bounds__(SF_THREADS) self_split_combine_kernel(const double2 *__restrict__ Zt, int N, int R, int S,
                                                                        size_t ntl, size_t tl_first,
                                                                        const double2 *__restrict__ What2,
                                                                        double *__restrict__ Ppart2,
                                                                        double2 *__restrict__ a_part) {
    extern __shared__ double2 sm2[];
    const int RS = R * S;
    double2 *Zs = sm2;
    double2 *Wh = Zs + 2 * RS;
    double *acc = reinterpret_cast<double *>(Wh + RS);
    double2 *Wr = reinterpret_cast<double2 *>(acc + RS + (RS & 1));
    double2 *red = Wr + R;
    const int c = blockIdx.x, C = gridDim.x;
    const size_t g = blockIdx.y, G = gridDim.y;
    const int tid = threadIdx.x;
    for (int e = tid; e < RS; e += SF_THREADS) {
        const int k2 = e / S, pos = e - k2 * S;
        Wh[e] = __ldg(&What2[(size_t)k2 * N + c * S + pos]);
        acc[e] = 0.0;
    }
    for (int r = tid; r < R; r += SF_THREADS) {
        double sn, cs;
        sincospi(-2.0 * (double)r / (double)R, &sn, &cs);
        Wr[r] = make_double2(cs, sn);
    }
    const size_t per = (ntl + G - 1) / G;
    const size_t t_begin = g * per, t_end = min(ntl, t_begin + per);
    auto prefetch = [&](size_t t, int buf) {
        const double2 *src = Zt + t * (size_t)R * N + (size_t)c * S;
        double2 *dst = Zs + buf * RS;
        for (int e = tid; e < RS; e += SF_THREADS) {
            const int r = e / S, pos = e - r * S;
            cp_async16(&dst[e], &src[(size_t)r * N + pos]);
        }
        cp_async_commit();
    };
    if (t_begin < t_end) prefetch(t_begin, 0);
    int buf = 0;
    for (size_t t = t_begin; t < t_end; t++, buf ^= 1) {
        if (t + 1 < t_end) {
            prefetch(t + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();  // slice t visible to all; everyone is done with iteration t-1 (its red[] slot is complete)
        if (tid == 0 && t > t_begin) {  // deferred block sum of the previous timeline
            const double2 *rd = red + ((buf ^ 1) * (SF_THREADS / 32));
            double2 sum = rd[0];
#pragma unroll
            for (int w = 1; w < SF_THREADS / 32; w++) {
                sum.x += rd[w].x;
                sum.y += rd[w].y;
            }
            a_part[(tl_first + t - 1) * C + c] = sum;
        }
        const double2 *Z = Zs + buf * RS;
        double2 ap = make_double2(0.0, 0.0);
        for (int e = tid; e < RS; e += SF_THREADS) {
            const int k2 = e / S, pos = e - k2 * S;
            double xr = 0.0, xi = 0.0;
            int idx = 0;
            for (int r = 0; r < R; r++) {
                const double2 z = Z[r * S + pos];
                const double2 w = Wr[idx];
                xr = fma(z.x, w.x, fma(-z.y, w.y, xr));
                xi = fma(z.x, w.y, fma(z.y, w.x, xi));
                idx += k2;
                if (idx >= R) idx -= R;
            }
            const double pw = fma(xr, xr, xi * xi);
            acc[e] += pw;
            const double2 wh = Wh[e];
            ap.x = fma(pw, wh.x, ap.x);
            ap.y = fma(pw, wh.y, ap.y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ap.x += __shfl_xor_sync(0xffffffffu, ap.x, o);
            ap.y += __shfl_xor_sync(0xffffffffu, ap.y, o);
        }
        if ((tid & 31) == 0) red[buf * (SF_THREADS / 32) + (tid >> 5)] = ap;
        __syncthreads();  // buffer `buf` may be overwritten by the prefetch of iteration t+1
    }
    if (tid == 0 && t_begin < t_end) {
        const double2 *rd = red + ((buf ^ 1) * (SF_THREADS / 32));
        double2 sum = rd[0];
#pragma unroll
        for (int w = 1; w < SF_THREADS / 32; w++) {
            sum.x += rd[w].x;
            sum.y += rd[w].y;
        }
        a_part[(tl_first + t_end - 1) * C + c] = sum;
    }
    for (int e = tid; e < RS; e += SF_THREADS) {
        const int k2 = e / S, pos = e - k2 * S;
        Ppart2[(g * R + k2) * (size_t)N + c * S + pos] = acc[e];
    }
}

// Compile-time R (combine kernels below): a thread owns one position, keeps its R inputs in registers and evaluates the R
// outputs with the twiddles exp(-2 pi i k / R) read as constant-bank operands (c_Wr, set at plan creation) -- no index
// arithmetic and no shared-memory traffic for the matrix.
__constant__ double2 c_Wr[64];

// R <= 8: one R-point DFT per position and timeline, grid = (C, G) with 256 positions per CTA.  The inputs of the next D-1
// timelines are in flight as 16-byte cp.async into a shared-memory ring of thread-private columns [stage][r][tid] (a thread
// reads only what it copied itself, so the ring needs no barrier) -- the kernel is HBM-latency bound and, with the ring, runs
// at the HBM floor (R = 5: 248 us for 1.61 GB).  (A register-prefetch version served R = 9...16 until the two-stage kernel below
// took them over: 465 / 516 us at R = 13 / 15, one 256-thread CTA per SM.)
template <int R>
__global__ void __launch_bounds__(SF_THREADS) self_split_combine_ring_kernel(const double2 *__restrict__ Zt, int N, size_t ntl,
                                                                             size_t tl_first,
                                                                             const double2 *__restrict__ What2,
                                                                             double *__restrict__ Ppart2,
                                                                             double2 *__restrict__ a_part) {
    constexpr int D = (R <= 6) ? 4 : 3;
    constexpr int SLOTS = 8;
    extern __shared__ double2 ring[];  // [D][R][SF_THREADS]
    __shared__ double2 red[SLOTS][SF_THREADS / 32];
    const int c = blockIdx.x, C = gridDim.x;
    const size_t g = blockIdx.y, G = gridDim.y;
    const int tid = threadIdx.x;
    const int pos = c * SF_THREADS + tid;
    double acc[R];
    double2 wh[R];
#pragma unroll
    for (int k2 = 0; k2 < R; k2++) {
        acc[k2] = 0.0;
        wh[k2] = __ldg(&What2[(size_t)k2 * N + pos]);
    }
    const size_t per = (ntl + G - 1) / G;
    const size_t t_begin = g * per, t_end = min(ntl, t_begin + per);
    auto prefetch = [&](size_t t, int stage) {
        if (t < t_end) {
#pragma unroll
            for (int r = 0; r < R; r++) cp_async16(&ring[(stage * R + r) * SF_THREADS + tid], &Zt[(t * R + r) * (size_t)N + pos]);
        }
        cp_async_commit();  // (possibly empty) group: keeps the wait count uniform
    };
#pragma unroll
    for (int d = 0; d < D - 1; d++) prefetch(t_begin + d, d);
    auto flush = [&](size_t t_first, int count) {
        __syncthreads();
        if (tid < count) {
            double2 sum = red[tid][0];
#pragma unroll
            for (int w = 1; w < SF_THREADS / 32; w++) {
                sum.x += red[tid][w].x;
                sum.y += red[tid][w].y;
            }
            a_part[(tl_first + t_first + tid) * C + c] = sum;
        }
        __syncthreads();
    };
    int slot = 0, stage = 0;
    for (size_t t = t_begin; t < t_end; t++) {
        // stage (stage + D - 1) % D was consumed in the previous iteration (its values went through the arithmetic below)
        prefetch(t + D - 1, stage == 0 ? D - 1 : stage - 1);
        cp_async_wait<D - 1>();  // the group of timeline t is complete
        double2 z[R];
#pragma unroll
        for (int r = 0; r < R; r++) z[r] = ring[(stage * R + r) * SF_THREADS + tid];
        stage = (stage + 1 == D) ? 0 : stage + 1;
        double2 ap = make_double2(0.0, 0.0);
#pragma unroll
        for (int k2 = 0; k2 < R; k2++) {
            double xr = z[0].x, xi = z[0].y;
#pragma unroll
            for (int r = 1; r < R; r++) {
                const int idx = (r * k2) % R;
                if (idx == 0) {
                    xr += z[r].x;
                    xi += z[r].y;
                } else {
                    xr = fma(z[r].x, c_Wr[idx].x, fma(-z[r].y, c_Wr[idx].y, xr));
                    xi = fma(z[r].x, c_Wr[idx].y, fma(z[r].y, c_Wr[idx].x, xi));
                }
            }
            const double pw = fma(xr, xr, xi * xi);
            acc[k2] += pw;
            ap.x = fma(pw, wh[k2].x, ap.x);
            ap.y = fma(pw, wh[k2].y, ap.y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ap.x += __shfl_xor_sync(0xffffffffu, ap.x, o);
            ap.y += __shfl_xor_sync(0xffffffffu, ap.y, o);
        }
        if ((tid & 31) == 0) red[slot][tid >> 5] = ap;
        if (++slot == SLOTS) {
            flush(t + 1 - SLOTS, SLOTS);
            slot = 0;
        }
    }
    if (slot) flush(t_end - slot, slot);
#pragma unroll
    for (int k2 = 0; k2 < R; k2++) Ppart2[(g * R + k2) * (size_t)N + pos] = acc[k2];
}

// Two-stage version for composite R = R1*R2 (17..64): r = R2 a + b, k2 = c + R1 d,
//     X[c + R1 d] = sum_b w_R2^{b d} ( w_R^{b c} sum_a z[R2 a + b] w_R1^{a c} )
// in place in the thread's R registers (R1^2 R2 + R1 R2^2 complex multiply-adds instead of R^2).  The power accumulators and
// the weights live in shared memory as thread-private columns [k2][tid].
// THREADS positions per CTA.  The kernel keeps R weights and R power accumulators per position in shared memory (24 R bytes per
// thread) and holds a timeline's R inputs in registers, so a CTA alternates between a burst of loads and arithmetic; with 256
// threads only ONE such CTA fits an SM and nothing covers its load latency (captured at R = 25: 483 us for 1.6 GB = 3.3 TB/s,
// long-scoreboard stalls 45 % of the samples).  CTAs of 64 threads put several independent ones on an SM, and the next
// timeline's inputs are prefetched with cp.async under the arithmetic.
template <int R1, int R2, int THREADS>
__global__ void __launch_bounds__(THREADS) self_split_combine_2s_kernel(const double2 *__restrict__ Zt, int N, size_t ntl,
                                                                           size_t tl_first,
                                                                           const double2 *__restrict__ What2,
                                                                           double *__restrict__ Ppart2,
                                                                           double2 *__restrict__ a_part) {
    constexpr int R = R1 * R2;
    extern __shared__ double2 sm3[];
    double2 *wh = sm3;               // [R][THREADS] weights
    double2 *ring = wh + R * THREADS;  // [R][THREADS] inputs of the next timeline (thread-private columns)
    double acc[R];                   // power accumulators of this position, registers
    __shared__ double2 red[2][THREADS / 32];
    const int c_ = blockIdx.x, C = gridDim.x;
    const size_t g = blockIdx.y, G = gridDim.y;
    const int tid = threadIdx.x;
    const int pos = c_ * THREADS + tid;
#pragma unroll
    for (int k2 = 0; k2 < R; k2++) {
        acc[k2] = 0.0;
        wh[k2 * THREADS + tid] = __ldg(&What2[(size_t)k2 * N + pos]);
    }
    const size_t per = (ntl + G - 1) / G;
    const size_t t_begin = g * per, t_end = min(ntl, t_begin + per);
    // the inputs of the NEXT timeline are in flight as 16-byte cp.async into thread-private columns of shared memory (no
    // barrier: a thread reads only what it copied itself) while this timeline's R values are transformed
    auto prefetch = [&](size_t t) {
        if (t < t_end) {
#pragma unroll
            for (int r = 0; r < R; r++) cp_async16(&ring[r * THREADS + tid], &Zt[(t * R + r) * (size_t)N + pos]);
        }
        cp_async_commit();
    };
    prefetch(t_begin);
    int buf = 0;
    for (size_t t = t_begin; t < t_end; t++, buf ^= 1) {
        cp_async_wait<0>();
        double2 z[R];
#pragma unroll
        for (int r = 0; r < R; r++) z[r] = ring[r * THREADS + tid];
        // stage 1: R1-point DFT over a for every b, then the twiddle w_R^{b c}
#pragma unroll
        for (int bb = 0; bb < R2; bb++) {
            double2 y[R1];
#pragma unroll
            for (int c = 0; c < R1; c++) {
                double xr = z[bb].x, xi = z[bb].y;
#pragma unroll
                for (int a = 1; a < R1; a++) {
                    const int idx = (R2 * a * c) % R;
                    const double2 v = z[R2 * a + bb];
                    if (idx == 0) {
                        xr += v.x;
                        xi += v.y;
                    } else {
                        xr = fma(v.x, c_Wr[idx].x, fma(-v.y, c_Wr[idx].y, xr));
                        xi = fma(v.x, c_Wr[idx].y, fma(v.y, c_Wr[idx].x, xi));
                    }
                }
                const int ti = (bb * c) % R;
                if (ti == 0) {
                    y[c] = make_double2(xr, xi);
                } else {
                    y[c] = make_double2(fma(xr, c_Wr[ti].x, -xi * c_Wr[ti].y), fma(xr, c_Wr[ti].y, xi * c_Wr[ti].x));
                }
            }
#pragma unroll
            for (int c = 0; c < R1; c++) z[R2 * c + bb] = y[c];  // Y_b[c] stored at slot R2 c + b
        }
        prefetch(t + 1);  // stage 1 has consumed every input: the ring column is free (its loads have completed)
        // stage 2: R2-point DFT over b for every c; output k2 = c + R1 d
        double2 ap = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < R1; c++) {
#pragma unroll
            for (int d = 0; d < R2; d++) {
                double xr = z[R2 * c].x, xi = z[R2 * c].y;
#pragma unroll
                for (int bb = 1; bb < R2; bb++) {
                    const int idx = (R1 * bb * d) % R;
                    const double2 v = z[R2 * c + bb];
                    if (idx == 0) {
                        xr += v.x;
                        xi += v.y;
                    } else {
                        xr = fma(v.x, c_Wr[idx].x, fma(-v.y, c_Wr[idx].y, xr));
                        xi = fma(v.x, c_Wr[idx].y, fma(v.y, c_Wr[idx].x, xi));
                    }
                }
                const int k2 = c + R1 * d;
                const double pw = fma(xr, xr, xi * xi);
                acc[k2] += pw;
                const double2 w = wh[k2 * THREADS + tid];
                ap.x = fma(pw, w.x, ap.x);
                ap.y = fma(pw, w.y, ap.y);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ap.x += __shfl_xor_sync(0xffffffffu, ap.x, o);
            ap.y += __shfl_xor_sync(0xffffffffu, ap.y, o);
        }
        if ((tid & 31) == 0) red[buf][tid >> 5] = ap;
        __syncthreads();
        if (tid == 0) {
            double2 sum = red[buf][0];
#pragma unroll
            for (int w = 1; w < THREADS / 32; w++) {
                sum.x += red[buf][w].x;
                sum.y += red[buf][w].y;
            }
            a_part[(tl_first + t) * C + c_] = sum;
        }
    }
#pragma unroll
    for (int k2 = 0; k2 < R; k2++) Ppart2[(g * R + k2) * (size_t)N + pos] = acc[k2];
}

// Timeline groups (grid.y) of a combine kernel with C position slices: as many CTAs as are resident at once, rounded DOWN
// to whole groups -- a few CTAs more than one wave would double (or, at one CTA per SM, add half to) the kernel's duration.
// (The process drives one device; residency is cached per kernel instantiation.)
template <typename Kernel>
size_t resident_ctas(Kernel kernel, size_t smem, int threads = SF_THREADS) {
    int per_sm = 1, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return (size_t)per_sm * (size_t)sms;
}
inline size_t combine_groups(size_t resident, size_t C, size_t cap) { return std::max<size_t>(1, std::min(resident / C, cap)); }

template <int R1, int R2>
size_t launch_combine_2s(size_t C, size_t gcap, cudaStream_t st, const double2 *Zt, int N, size_t nt, size_t t0, const double2 *w2,
                         double *Ppart2, double2 *a_part) {
    constexpr int T = SF_2S_THREADS;
    constexpr size_t smem = (size_t)R1 * R2 * T * 2 * sizeof(double2);  // weights + the prefetched inputs of the next timeline
    static size_t resident = 0;
    if (!resident) {
        cudaFuncSetAttribute(self_split_combine_2s_kernel<R1, R2, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        resident = resident_ctas(self_split_combine_2s_kernel<R1, R2, T>, smem, T);
    }
    const size_t G = combine_groups(resident, C, gcap);
    self_split_combine_2s_kernel<R1, R2, T><<<dim3((unsigned)C, (unsigned)G), T, smem, st>>>(Zt, N, nt, t0, w2, Ppart2, a_part);
    return G;
}

// compile-time R <= 8: ring kernel
template <int R>
size_t launch_combine_reg(size_t C, size_t gcap, cudaStream_t st, const double2 *Zt, int N, size_t nt, size_t t0, const double2 *w2,
                          double *Ppart2, double2 *a_part) {
    static_assert(R <= 8, "R > 8 goes through the two-stage kernel");
    static size_t resident = 0;
    constexpr size_t smem = (size_t)((R <= 6) ? 4 : 3) * R * SF_THREADS * sizeof(double2);
    if (!resident) {
        cudaFuncSetAttribute(self_split_combine_ring_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        resident = resident_ctas(self_split_combine_ring_kernel<R>, smem);
    }
    const size_t G = combine_groups(resident, C, gcap);
    self_split_combine_ring_kernel<R><<<dim3((unsigned)C, (unsigned)G), SF_THREADS, smem, st>>>(Zt, N, nt, t0, w2, Ppart2, a_part);
    return G;
}

// composite R the two-stage kernel is instantiated for: (R1, R2) with R1 <= R2, both <= 8, 9 <= R <= 32
struct SplitFactor {
    int R, R1, R2;
};
constexpr SplitFactor kSplitFactors[] = {{9, 3, 3}, {10, 2, 5}, {12, 3, 4}, {14, 2, 7}, {15, 3, 5}, {16, 4, 4}, {18, 3, 6}, {20, 4, 5}, {21, 3, 7}, {24, 4, 6}, {25, 5, 5}, {28, 4, 7}, {30, 5, 6}, {32, 4, 8}};

// P[perm[i]] += sum_g Ppart2[g][i]   (split layout -> residue-major layout; perm is a bijection)
__global__ void sf_reduce_ppart_perm_kernel(const double *__restrict__ Ppart, size_t G, size_t len, const int *__restrict__ perm,
                                            double *__restrict__ P) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= len) return;
    double sum = 0.0;
    for (size_t g = 0; g < G; g++) sum += Ppart[g * len + i];
    P[perm[i]] += sum;
}

__global__ void sf_gather_weights_kernel(const double2 *__restrict__ w, const int *__restrict__ perm, size_t len,
                                         double2 *__restrict__ w2) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < len) w2[i] = w[perm[i]];
}

// P[i] += sum_g Ppart[g][i].  Block (64, 4): thread (x, y) sums the partials g = y, y + 4, ... of entry i = 64 blockIdx.x + x
// with four loads in flight, the four strands are combined in a fixed order (bit-reproducible, no atomics).  (One thread
// per entry walking all G partials was 43 us for G = 296, L = 2048: 8 CTAs of dependent loads.)
__global__ void __launch_bounds__(256) sf_reduce_ppart_kernel(const double *__restrict__ Ppart, size_t G, size_t len,
                                                              double *__restrict__ P) {
    __shared__ double sh[4][64];
    const size_t i = blockIdx.x * (size_t)64 + threadIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (i < len) {
        size_t g = threadIdx.y;
        for (; g + 12 < G; g += 16) {
            s0 += Ppart[g * len + i];
            s1 += Ppart[(g + 4) * len + i];
            s2 += Ppart[(g + 8) * len + i];
            s3 += Ppart[(g + 12) * len + i];
        }
        for (; g < G; g += 4) s0 += Ppart[g * len + i];
    }
    sh[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.y == 0 && i < len) P[i] += (sh[0][threadIdx.x] + sh[1][threadIdx.x]) + (sh[2][threadIdx.x] + sh[3][threadIdx.x]);
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
                  double2 *a_part, double2 *Wout, const double2 *Ain = nullptr, size_t ldA = 0) {
    const size_t smem = sf_smem_bytes(log2N);
#define SF_CASE(LN)                                                                                                    \
    case LN: {                                                                                                         \
        cudaFuncSetAttribute(self_fused_kernel<LN, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        self_fused_kernel<LN, GEN><<<grid, SF_THREADS, smem, st>>>(xyz, b, qs, NF, NM, atom0, ntl, R, tw, What, Ppart, \
                                                                   a_part, Wout, Ain, ldA);                           \
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
    // two resident CTAs per SM (102 KB shared memory, 128 registers); R * G rounded DOWN to one wave: a handful of CTAs
    // beyond it would run as a second wave and double the kernel's duration (R = 5: 300 CTAs on 296 slots)
    size_t G = (2 * 148) / p->R;
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
    // any L >= 2NF-1 yields the same correlation: between 9 and 32 round R up to a value the two-stage combine kernel
    // of the split path is instantiated for (at most 2 more sub-transforms)
    if (p->R > 8 && p->R <= 32 && !getenv("SASSENA_SELF_EXACT_R")) {
        for (const SplitFactor &sf : kSplitFactors)
            if (sf.R >= p->R) {
                p->R = sf.R;
                break;
            }
    }
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
    if (launches) *launches += 2;
    // split path tables.  Measured (tools/probe_self.py): at R = 4 (NF = 7000) the fused kernel still wins (73 vs 77 ms per
    // 4.1e5 timelines), from R = 5 on the split path does (NF = 10000: 97 vs 162 ms; R = 25: 325 vs 1177 ms).
    // SASSENA_SELF_PATH=fused|split overrides for tests and experiments.
    p->split = p->R >= 5 && p->R <= 64;
    if (const char *e = getenv("SASSENA_SELF_PATH")) {
        if (!strcmp(e, "fused")) p->split = false;
        if (!strcmp(e, "split") && p->R >= 2 && p->R <= 64) p->split = true;
    }
    if (p->split) {
        const bool generic = getenv("SASSENA_SELF_GENERIC_COMBINE") != nullptr;
        p->reg_combine = p->R <= 8 && !generic;
        p->two_stage = false;
        for (const SplitFactor &sf : kSplitFactors) p->two_stage = p->two_stage || (sf.R == p->R && !generic);
        int S = p->two_stage ? SF_2S_THREADS : 256;
        while (!p->reg_combine && !p->two_stage && S > 8 && (size_t)S * p->R > 2048) S >>= 1;
        if ((size_t)S > p->N) S = (int)p->N;
        p->S = S;
        p->C = (int)(p->N / S);
        if (p->reg_combine || p->two_stage) {
            std::vector<double2> wr(64, make_double2(0.0, 0.0));
            for (int k = 0; k < p->R; k++) {
                // exact quadrant values, libm elsewhere (accurate to 1 ulp)
                const double a = -2.0 * 3.14159265358979323846 * (double)k / (double)p->R;
                wr[k] = make_double2(cos(a), sin(a));
                if ((4 * k) % p->R == 0) {
                    const int qd = (4 * k) / p->R;  // multiples of a quarter turn
                    const double cs[4] = {1, 0, -1, 0}, sn[4] = {0, -1, 0, 1};
                    wr[k] = make_double2(cs[qd & 3], sn[qd & 3]);
                }
            }
            cudaMemcpyToSymbolAsync(c_Wr, wr.data(), sizeof(double2) * 64, 0, cudaMemcpyHostToDevice, st);
        }
        if (cudaMalloc(&p->d_w2, sizeof(double2) * p->L) != cudaSuccess) return 2;
        if (cudaMalloc(&p->d_perm, sizeof(int) * p->L) != cudaSuccess) return 2;
        std::vector<int> inv(p->N), perm(p->L);
        for (size_t i = 0; i < p->N; i++) inv[f[i]] = (int)i;  // frequency -> position of the radix-16 order
        for (int k2 = 0; k2 < p->R; k2++)
            for (size_t pos = 0; pos < p->N; pos++) {
                const size_t freq = pos + p->N * k2;  // X[k1 + N k2]: kernel A stores Z in natural frequency order
                const size_t j = freq % p->R, k = freq / p->R;   // = X[R k + j] of the residue-major layout
                perm[(size_t)k2 * p->N + pos] = (int)(j * p->N + inv[k]);
            }
        cudaMemcpyAsync(p->d_perm, perm.data(), sizeof(int) * p->L, cudaMemcpyHostToDevice, st);
        sf_gather_weights_kernel<<<(unsigned)((p->L + 255) / 256), 256, 0, st>>>(p->d_w, p->d_perm, p->L, p->d_w2);
        cudaStreamSynchronize(st);  // perm[] is a stack buffer
        if (launches) *launches += 2;
    }
    cudaStreamSynchronize(st);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

void self_plan_destroy(SelfPlan *p) {
    if (p->d_tw) cudaFree(p->d_tw);
    if (p->d_w) cudaFree(p->d_w);
    if (p->d_freq) cudaFree(p->d_freq);
    if (p->d_w2) cudaFree(p->d_w2);
    if (p->d_perm) cudaFree(p->d_perm);
    p->d_w2 = nullptr;
    p->d_perm = nullptr;
    p->split = false;
    p->d_tw = nullptr;
    p->d_w = nullptr;
    p->d_freq = nullptr;
    p->NF = p->L = 0;
}

namespace {
constexpr size_t kSplitZBytes = (size_t)1536 << 20;  // batch buffer of twiddled sub-transforms

size_t split_batch(const SelfPlan *p, size_t ntl) {
    const size_t per_tl = p->L * sizeof(double2);
    return std::max<size_t>(1, std::min(ntl, kSplitZBytes / per_tl));
}
size_t split_groups_b(const SelfPlan *p, size_t tb) {
    // upper bound of the timeline groups (sizes the partial spectra): two combine CTAs per SM, the two-stage kernel's small
    // CTAs up to seven
    size_t G = ((p->two_stage ? 7 : 2) * 148 + p->C - 1) / p->C;
    if (G > tb) G = tb;
    return std::max<size_t>(G, 1);
}
size_t split_smem_b(const SelfPlan *p) {
    const size_t RS = (size_t)p->R * p->S;
    return 3 * RS * sizeof(double2) + (RS + (RS & 1)) * sizeof(double) + p->R * sizeof(double2) +
           2 * (SF_THREADS / 32) * sizeof(double2);
}

void launch_split_fft(size_t G, cudaStream_t st, const float *xyz, const double *b, const double *qs, int NF, int NM, size_t atom0,
                      size_t tl_first, size_t ntl, const SelfPlan *p, int dec, double2 *Zt) {
    constexpr size_t N_ = 4096;
    const size_t coords = ((p->NF / p->R + 1) * 3 * sizeof(float) + 16 + 15) & ~(size_t)15;  // + the alignment offset
    // tables grow with R, the coordinate buffer with NF/R <= N/2; 113 KB lets two CTAs share an SM (per-device attribute)
    const size_t smem = (N_ + 256 + (p->L >> 8)) * sizeof(double2) + 8 * sizeof(double) + coords;
    cudaFuncSetAttribute(self_split_fft12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    self_split_fft12_kernel<<<(unsigned)G, SF_THREADS, smem, st>>>(xyz, b, qs, NF, NM, atom0, tl_first, ntl, p->R, dec, Zt);
}
}  // namespace

size_t self_work_bytes(const SelfPlan *p, size_t ntl) {
    if (p->split) {
        const size_t tb = split_batch(p, ntl);
        return align256(tb * p->L * sizeof(double2)) + align256(split_groups_b(p, tb) * p->L * sizeof(double)) +
               align256(ntl * p->C * sizeof(double2)) + align256(ntl * sizeof(double2)) + align256(p->L * sizeof(double2));
    }
    const size_t G = pick_groups(p, ntl);
    return align256(G * p->L * sizeof(double)) + align256(ntl * p->R * sizeof(double2)) + align256(ntl * sizeof(double2)) +
           align256(p->L * sizeof(double2));
}

static int self_power_accumulate_split(const SelfPlan *p, const float *d_xyz_by_atom, const double *d_b, const double *d_qs,
                                       size_t NM, size_t atom0, size_t natoms, void *d_work, double *d_P, double *d_acc,
                                       int dec, cudaStream_t st) {
    const size_t ntl = natoms * NM;
    const size_t tb = split_batch(p, ntl);
    const size_t GB = split_groups_b(p, tb);
    char *w = reinterpret_cast<char *>(d_work);
    double2 *Zt = reinterpret_cast<double2 *>(w);
    w += align256(tb * p->L * sizeof(double2));
    double *Ppart2 = reinterpret_cast<double *>(w);
    w += align256(GB * p->L * sizeof(double));
    double2 *a_part = reinterpret_cast<double2 *>(w);
    w += align256(ntl * p->C * sizeof(double2));
    double2 *a_tl = reinterpret_cast<double2 *>(w);
    const size_t smem_b = split_smem_b(p);
    static size_t configured = 0;
    if (!p->reg_combine && !p->two_stage && smem_b > configured) {
        cudaFuncSetAttribute(self_split_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
        configured = smem_b;
    }
    int launches = 0;
    for (size_t t0 = 0; t0 < ntl; t0 += tb) {
        const size_t nt = std::min(tb, ntl - t0);
        // two sub-transform CTAs per SM (128 registers each); the (timeline, r) pairs are dealt out evenly
        const size_t GA = std::min<size_t>(nt * (size_t)p->R, 2 * 148);
        launch_split_fft(GA, st, d_xyz_by_atom, d_b, d_qs, (int)p->NF, (int)NM, atom0, t0, nt, p, dec, Zt);
        size_t G = split_groups_b(p, nt);  // upper bound (Ppart2 is sized for it); the launchers pick whole waves
        if (p->reg_combine) {
#define SF_RCASE(RR)                                                                                        \
    case RR:                                                                                                \
        G = launch_combine_reg<RR>(p->C, G, st, Zt, (int)p->N, nt, t0, p->d_w2, Ppart2, a_part);            \
        break;
            switch (p->R) {
                SF_RCASE(2) SF_RCASE(3) SF_RCASE(4) SF_RCASE(5) SF_RCASE(6) SF_RCASE(7) SF_RCASE(8)
                default: break;
            }
#undef SF_RCASE
        } else if (p->two_stage) {
#define SF_2CASE(RR, A, B)                                                                                  \
    case RR:                                                                                                \
        G = launch_combine_2s<A, B>(p->C, G, st, Zt, (int)p->N, nt, t0, p->d_w2, Ppart2, a_part);           \
        break;
            switch (p->R) {
                SF_2CASE(9, 3, 3) SF_2CASE(10, 2, 5) SF_2CASE(12, 3, 4) SF_2CASE(14, 2, 7) SF_2CASE(15, 3, 5) SF_2CASE(16, 4, 4)
                SF_2CASE(18, 3, 6) SF_2CASE(20, 4, 5) SF_2CASE(21, 3, 7) SF_2CASE(24, 4, 6) SF_2CASE(25, 5, 5) SF_2CASE(28, 4, 7)
                SF_2CASE(30, 5, 6) SF_2CASE(32, 4, 8)
                default: break;
            }
#undef SF_2CASE
        } else {
            self_split_combine_kernel<<<dim3((unsigned)p->C, (unsigned)G), SF_THREADS, smem_b, st>>>(Zt, (int)p->N, p->R, p->S, nt,
                                                                                                t0, p->d_w2, Ppart2, a_part);
        }
        sf_reduce_ppart_perm_kernel<<<(unsigned)((p->L + 255) / 256), 256, 0, st>>>(Ppart2, G, p->L, p->d_perm, d_P);
        launches += 3;
    }
    const double norm = 1.0 / ((double)p->NF * (double)p->L);
    sf_reduce_apart_kernel<<<(unsigned)((ntl + 255) / 256), 256, 0, st>>>(a_part, ntl, p->C, norm, a_tl);
    sf_reduce_atl_kernel<<<1, 1024, 0, st>>>(a_tl, ntl, d_acc);
    return launches + 2;
}

int self_decimate_layout(const float *d_src, float *d_dst, size_t natoms, size_t NF, int R, int forward, cudaStream_t st) {
    const size_t n = natoms * NF;
    if (n == 0) return 0;
    sf_decimate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_src, d_dst, natoms, (int)NF, R, forward);
    return 1;
}

int self_power_accumulate(const SelfPlan *p, const float *d_xyz_by_atom, const double *d_b, const double *d_qs,
                          size_t NM, size_t atom0, size_t natoms, void *d_work, double *d_P, double *d_acc, int dec,
                          cudaStream_t st) {
    const size_t ntl = natoms * NM;
    if (ntl == 0) return 0;
    if (p->split) return self_power_accumulate_split(p, d_xyz_by_atom, d_b, d_qs, NM, atom0, natoms, d_work, d_P, d_acc, dec, st);
    const size_t G = pick_groups(p, ntl);
    char *w = reinterpret_cast<char *>(d_work);
    double *Ppart = reinterpret_cast<double *>(w);
    w += align256(G * p->L * sizeof(double));
    double2 *a_part = reinterpret_cast<double2 *>(w);
    w += align256(ntl * p->R * sizeof(double2));
    double2 *a_tl = reinterpret_cast<double2 *>(w);
    launch_fused<GEN_AMPLITUDE>(p->log2N, dim3(p->R, (unsigned)G), st, d_xyz_by_atom, d_b, d_qs, (int)p->NF, (int)NM, atom0,
                                ntl, p->R, p->d_tw, p->d_w, Ppart, a_part, nullptr);
    sf_reduce_ppart_kernel<<<(unsigned)((p->L + 63) / 64), dim3(64, 4), 0, st>>>(Ppart, G, p->L, d_P);
    const double norm = 1.0 / ((double)p->NF * (double)p->L);
    sf_reduce_apart_kernel<<<(unsigned)((ntl + 255) / 256), 256, 0, st>>>(a_part, ntl, p->R, norm, a_tl);
    sf_reduce_atl_kernel<<<1, 1024, 0, st>>>(a_tl, ntl, d_acc);
    return 4;
}

// DSP of timelines that already exist in memory (the coherent and multipole devices' A[nt][ldA]) with the in-SM transform:
// every timeline is read once (16 NF bytes, the algorithmic minimum), nothing else touches HBM.  R == 1 only (2NF-1 <= 4096).
static size_t loaded_groups(size_t nt) {
    // whole waves of resident CTAs (two per SM by the launch bounds); a CTA keeps its power accumulator across its timelines
    return std::max<size_t>(1, std::min<size_t>(nt, 2 * 148));
}
size_t self_loaded_work_bytes(const SelfPlan *p, size_t nt) {
    const size_t G = loaded_groups(nt);
    return align256(G * p->L * sizeof(double)) + align256(nt * sizeof(double2)) + align256(nt * sizeof(double2)) +
           align256(p->L * sizeof(double2));
}
int self_power_accumulate_loaded(const SelfPlan *p, const double2 *d_A, size_t ldA, size_t nt, void *d_work, double *d_P,
                                 double *d_acc, cudaStream_t st) {
    if (nt == 0) return 0;
    if (p->R != 1) return -1;
    const size_t G = loaded_groups(nt);
    char *w = reinterpret_cast<char *>(d_work);
    double *Ppart = reinterpret_cast<double *>(w);
    w += align256(G * p->L * sizeof(double));
    double2 *a_part = reinterpret_cast<double2 *>(w);
    w += align256(nt * sizeof(double2));
    double2 *a_tl = reinterpret_cast<double2 *>(w);
    launch_fused<GEN_LOAD>(p->log2N, dim3(1, (unsigned)G), st, nullptr, nullptr, nullptr, (int)p->NF, 1, 0, nt, 1, p->d_tw, p->d_w,
                           Ppart, a_part, nullptr, d_A, ldA);
    sf_reduce_ppart_kernel<<<(unsigned)((p->L + 63) / 64), dim3(64, 4), 0, st>>>(Ppart, G, p->L, d_P);
    const double norm = 1.0 / ((double)p->NF * (double)p->L);
    sf_reduce_apart_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(a_part, nt, 1, norm, a_tl);
    sf_reduce_atl_kernel<<<1, 1024, 0, st>>>(a_tl, nt, d_acc);
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
