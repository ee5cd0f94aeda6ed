// correlate.cu — batched FFT autocorrelation and the store()/reduce step (K3/K4).
//
// Restates smath::auto_correlate_fftw (reference src/math/smath.cpp:141-156) + store()
// (src/scatter_devices/all_vectors_scatter_device.cpp:231-236) for a batch of timelines:
//     X = FFT_L(x zero-padded),  P = |X|^2,  c = IFFT_L(P),  C[tau] = c[tau] / (L (NF - tau)),  tau < NF
// with L the power of two >= 2 NF (the reference pads to exactly 2 NF; any L >= 2NF-1 gives the same
// linear autocorrelation, the 1/(2NF) in smath.cpp:151 is FFTW's unnormalised-transform factor).
//
// Two algebraic moves keep the batch HBM-light and deterministic, both exact up to FP64 rounding:
//   (1) the inverse transform is linear, so  sum_m C_m = IFFT(sum_m P_m): one inverse per |q|;
//   (2) store() needs a_m = mean_tau C_m[tau] = (1/(NF L)) sum_k P_m[k] What[k] with the fixed weights
//       What[k] = sum_{tau<NF} e^{+2 pi i k tau / L} / (NF - tau), so no per-timeline inverse at all.
//
// The length-L transform is a four-step FFT, L = N1*N2:
//   pass A (corr_colfft_kernel): for each n2, FFT over n1 of x[n1*N2+n2] (stride N2), times W_L^{n2 k1},
//          written to Y[k1*N2+n2]; a CTA handles a tile of adjacent columns so global accesses stay in
//          128-byte lines; rows beyond NF are implicit zeros (the padded input is never materialised);
//   pass B (corr_rowfft_kernel): for each k1, FFT over n2 (contiguous), result X[k1 + N1*k2].
// Sub-FFTs run in shared memory (radix-2 DIT, FP64).  Power spectra are kept in the "internal order"
// i = k1*N2 + k2; the inverse pass reads them through the matching permutation.
// No atomics anywhere: partial sums go to per-chunk buffers that are reduced in a fixed order, so
// results are bit-reproducible run to run.
#include "kernels.hpp"

#include <algorithm>
#include <cstdlib>

namespace sass {

namespace {

constexpr int FFT_THREADS = 256;
constexpr int MAX_LOG2_SUB = 11;  // sub-FFT length <= 2048

enum { LOAD_TIMELINE = 0, LOAD_REALPERM = 1, LOAD_WEIGHTS = 2 };
enum { OUT_POWER = 0, OUT_FINAL = 1, OUT_INTERNAL = 2 };

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return bits == 0 ? 0u : (__brev(v) >> (32 - bits)); }

// In-place radix-2 DIT FFT over the rows of s[N][cols] (input rows already bit-reversed).
// tw holds W_Nmax^k (forward sign, k < Nmax/2); sign=+1 conjugates.
__device__ __forceinline__ void smem_fft_rows(double2 *s, int log2N, int cols, const double2 *__restrict__ tw,
                                              int log2Nmax, int sign) {
    const int N = 1 << log2N;
    const int nbf = (N >> 1) * cols;
    for (int st = 1; st <= log2N; st++) {
        const int half = 1 << (st - 1);
        __syncthreads();
        for (int id = threadIdx.x; id < nbf; id += blockDim.x) {
            const int c = id % cols;
            const int bf = id / cols;
            const int pos = bf & (half - 1);
            const int i0 = ((bf >> (st - 1)) << st) + pos;
            const int i1 = i0 + half;
            double2 w = __ldg(&tw[(size_t)pos << (log2Nmax - st)]);
            if (sign > 0) w.y = -w.y;
            const double2 u = s[i0 * cols + c];
            const double2 v = cmul(s[i1 * cols + c], w);
            s[i0 * cols + c] = make_double2(u.x + v.x, u.y + v.y);
            s[i1 * cols + c] = make_double2(u.x - v.x, u.y - v.y);
        }
    }
    __syncthreads();
}

// pass A.  grid.x = ntl * ntiles (tile fastest), dynamic smem = N1*cols*16 bytes.
template <int LOAD>
__global__ void __launch_bounds__(FFT_THREADS) corr_colfft_kernel(const void *__restrict__ in, size_t ld_in, size_t NF,
                                                                  int log2N1, int log2N2, int cols,
                                                                  const double2 *__restrict__ tw, int log2Nmax,
                                                                  int sign, double2 *__restrict__ Y) {
    extern __shared__ double2 s[];
    const int N1 = 1 << log2N1, N2 = 1 << log2N2;
    const size_t L = (size_t)1 << (log2N1 + log2N2);
    const unsigned ntiles = N2 / cols;
    const unsigned tile = blockIdx.x % ntiles;
    const size_t tl = blockIdx.x / ntiles;
    const int col0 = tile * cols;

    for (int idx = threadIdx.x; idx < N1 * cols; idx += blockDim.x) {
        const int n1 = idx / cols, c = idx % cols;
        const size_t n = (size_t)n1 * N2 + col0 + c;
        double2 v = make_double2(0.0, 0.0);
        if (LOAD == LOAD_TIMELINE) {
            if (n < NF) v = reinterpret_cast<const double2 *>(in)[tl * ld_in + n];
        } else if (LOAD == LOAD_REALPERM) {
            // natural index n -> internal location (n mod N1)*N2 + n / N1
            v.x = reinterpret_cast<const double *>(in)[tl * ld_in + (((n & (size_t)(N1 - 1)) << log2N2) + (n >> log2N1))];
        } else {
            if (n < NF) v.x = 1.0 / (double)(NF - n);
        }
        s[bitrev(n1, log2N1) * cols + c] = v;
    }
    smem_fft_rows(s, log2N1, cols, tw, log2Nmax, sign);
    const double inv = (double)sign * 2.0 / (double)L;
    for (int idx = threadIdx.x; idx < N1 * cols; idx += blockDim.x) {
        const int k1 = idx / cols, c = idx % cols;
        const int n2 = col0 + c;
        double sn, cs;
        sincospi(inv * (double)((size_t)n2 * k1), &sn, &cs);  // W_L^{-sign... } = exp(sign*2*pi*i*n2*k1/L)
        Y[tl * L + (size_t)k1 * N2 + n2] = cmul(s[idx], make_double2(cs, sn));
    }
}

// pass B.  grid = (N1, nchunks); each CTA loops over the timelines of its chunk.
//   OUT_POWER   : acc |X|^2 into Ppart[chunk][k1*N2+k2]; a_part[tl][k1] = sum_k2 |X|^2 * What[k1*N2+k2]
//   OUT_FINAL   : out[tau = k1+N1*k2] = scale * X / (L (NF-tau)) for tau < NF (conj optional)   (single timeline)
//   OUT_INTERNAL: out[k1*N2+k2] = X                                                             (single timeline)
template <int OUT>
__global__ void __launch_bounds__(FFT_THREADS) corr_rowfft_kernel(const double2 *__restrict__ Y, size_t ntl,
                                                                  size_t tl_per_chunk, size_t NF, int log2N1,
                                                                  int log2N2, const double2 *__restrict__ tw,
                                                                  int log2Nmax, int sign,
                                                                  const double2 *__restrict__ What,
                                                                  double *__restrict__ Ppart,
                                                                  double2 *__restrict__ a_part, double2 *__restrict__ out,
                                                                  double scale, int conj_out) {
    extern __shared__ double2 s[];
    __shared__ double2 red[FFT_THREADS / 32];
    const int N2 = 1 << log2N2;
    const size_t N1 = (size_t)1 << log2N1;
    const size_t L = (size_t)1 << (log2N1 + log2N2);
    const size_t k1 = blockIdx.x;
    const size_t chunk = blockIdx.y;
    const size_t tl0 = chunk * tl_per_chunk;
    const size_t tl1 = min(ntl, tl0 + tl_per_chunk);
    constexpr int SLOTS = (1 << MAX_LOG2_SUB) / FFT_THREADS;
    double acc[SLOTS];
#pragma unroll
    for (int i = 0; i < SLOTS; i++) acc[i] = 0.0;

    for (size_t tl = tl0; tl < tl1; tl++) {
        const double2 *row = Y + tl * L + k1 * N2;
        __syncthreads();
        for (int n2 = threadIdx.x; n2 < N2; n2 += blockDim.x) s[bitrev(n2, log2N2)] = row[n2];
        smem_fft_rows(s, log2N2, 1, tw, log2Nmax, sign);
        if (OUT == OUT_POWER) {
            double2 ap = make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < SLOTS; i++) {
                const int k2 = threadIdx.x + i * FFT_THREADS;
                if (k2 < N2) {
                    const double2 x = s[k2];
                    const double p = fma(x.x, x.x, x.y * x.y);
                    acc[i] += p;
                    const double2 w = __ldg(&What[k1 * N2 + k2]);
                    ap.x = fma(p, w.x, ap.x);
                    ap.y = fma(p, w.y, ap.y);
                }
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
                for (int w = 1; w < FFT_THREADS / 32; w++) {
                    t.x += red[w].x;
                    t.y += red[w].y;
                }
                a_part[tl * N1 + k1] = t;
            }
        } else if (OUT == OUT_FINAL) {
            for (int k2 = threadIdx.x; k2 < N2; k2 += blockDim.x) {
                const size_t tau = k1 + N1 * (size_t)k2;
                if (tau < NF) {
                    const double f = scale / ((double)L * (double)(NF - tau));
                    double2 x = s[k2];
                    out[tau] = make_double2(x.x * f, conj_out ? -x.y * f : x.y * f);
                }
            }
        } else {
            for (int k2 = threadIdx.x; k2 < N2; k2 += blockDim.x) out[k1 * N2 + k2] = s[k2];
        }
    }
    if (OUT == OUT_POWER) {
#pragma unroll
        for (int i = 0; i < SLOTS; i++) {
            const int k2 = threadIdx.x + i * FFT_THREADS;
            if (k2 < N2) Ppart[chunk * L + k1 * N2 + k2] = acc[i];
        }
    }
}

// P[i] += sum_chunk Ppart[chunk][i], fixed order
__global__ void reduce_ppart_kernel(const double *__restrict__ Ppart, size_t nchunks, size_t L, double *__restrict__ P) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= L) return;
    double sum = 0.0;
    for (size_t c = 0; c < nchunks; c++) sum += Ppart[c * L + i];
    P[i] += sum;
}

// a_tl[tl] = norm * sum_k a_part[tl][k]   (one warp per timeline, fixed order)
__global__ void reduce_apart_kernel(const double2 *__restrict__ a_part, size_t ntl, size_t N1, double norm,
                                    double2 *__restrict__ a_tl) {
    const size_t tl = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tl >= ntl) return;
    double2 sum = make_double2(0.0, 0.0);
    for (size_t k = lane; k < N1; k += 32) {
        const double2 v = a_part[tl * N1 + k];
        sum.x += v.x;
        sum.y += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
        sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
    }
    if (lane == 0) a_tl[tl] = make_double2(sum.x * norm, sum.y * norm);
}

// acc[0..2] += {sum re a, sum im a, sum |a|^2} over a_tl[0..n) — single CTA, fixed tree
__global__ void __launch_bounds__(1024) reduce_atl_kernel(const double2 *__restrict__ a_tl, size_t n,
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
        double R = 0, I = 0, Q = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            R += sh[0][w];
            I += sh[1][w];
            Q += sh[2][w];
        }
        acc[0] += R;
        acc[1] += I;
        acc[2] += Q;
    }
}

// DSP square / plain (smath.cpp:168-177 / identity) + the atfinal accumulation of store().
// thread per frame, sequential over timelines (fixed order).
__global__ void dsp_elementwise_at_kernel(const double2 *__restrict__ A, size_t ldA, size_t ntl, size_t NF, int square,
                                          double2 *__restrict__ at) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= NF) return;
    double2 sum = make_double2(0.0, 0.0);
    for (size_t m = 0; m < ntl; m++) {
        const double2 v = A[m * ldA + t];
        if (square) sum.x += fma(v.x, v.x, v.y * v.y);
        else {
            sum.x += v.x;
            sum.y += v.y;
        }
    }
    at[t].x += sum.x;
    at[t].y += sum.y;
}

// a_tl[m] = mean_t f(A[m][t])  (one CTA per timeline)
__global__ void __launch_bounds__(256) dsp_elementwise_mean_kernel(const double2 *__restrict__ A, size_t ldA, size_t NF,
                                                                    int square, double2 *__restrict__ a_tl) {
    __shared__ double2 red[8];
    const size_t m = blockIdx.x;
    double2 sum = make_double2(0.0, 0.0);
    for (size_t t = threadIdx.x; t < NF; t += blockDim.x) {
        const double2 v = A[m * ldA + t];
        if (square) sum.x += fma(v.x, v.x, v.y * v.y);
        else {
            sum.x += v.x;
            sum.y += v.y;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
        sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double2 t = red[0];
        for (int w = 1; w < 8; w++) {
            t.x += red[w].x;
            t.y += red[w].y;
        }
        const double inv = 1.0 / (double)NF;
        a_tl[m] = make_double2(t.x * inv, t.y * inv);
    }
}

__global__ void scale_complex_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n, double scale) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 v = in[i];
    out[i] = make_double2(v.x * scale, v.y * scale);
}

__global__ void twiddle_table_kernel(double2 *tw, size_t Nmax) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= Nmax / 2) return;
    double sn, cs;
    sincospi(-2.0 * (double)k / (double)Nmax, &sn, &cs);
    tw[k] = make_double2(cs, sn);
}

int ilog2_ceil(size_t v) {
    int l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}

int pick_cols(int log2N1, int log2N2) {
    // adjacent columns per pass-A CTA: 8 (one 128-byte line of complex128) unless shared memory says otherwise
    int cols = 8;
    while (cols > 1 && ((size_t)cols << log2N1) * sizeof(double2) > (size_t)160 * 1024) cols >>= 1;
    if (cols > (1 << log2N2)) cols = 1 << log2N2;
    return cols;
}

template <int LOAD>
int run_colfft(const CorrPlan *p, const void *in, size_t ld_in, size_t ntl, int sign, double2 *Y, cudaStream_t st) {
    const int cols = pick_cols(p->log2N1, p->log2N2);
    const size_t smem = ((size_t)cols << p->log2N1) * sizeof(double2);
    const unsigned ntiles = (1u << p->log2N2) / cols;
    // per-device attribute, set before every launch (see amplitude.cu allow_smem)
    cudaFuncSetAttribute(corr_colfft_kernel<LOAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int launches = 0;
    const size_t max_tl = (size_t)0x7fffffff / ntiles;
    for (size_t done = 0; done < ntl;) {
        size_t cnt = std::min(ntl - done, max_tl);
        const char *inp = reinterpret_cast<const char *>(in) + done * ld_in * (LOAD == LOAD_TIMELINE ? 16 : 8);
        corr_colfft_kernel<LOAD><<<(unsigned)(cnt * ntiles), FFT_THREADS, smem, st>>>(
            inp, ld_in, p->NF, p->log2N1, p->log2N2, cols, p->d_tw, ilog2_ceil(p->Nmax), sign, Y + done * p->L);
        launches++;
        done += cnt;
    }
    return launches;
}

size_t pick_chunks(const CorrPlan *p, size_t ntl) {
    const size_t N1 = (size_t)1 << p->log2N1;
    size_t want = (4 * 148 + N1 - 1) / N1;
    if (want < 1) want = 1;
    if (want > ntl) want = ntl;
    if (want > 65535) want = 65535;
    return want ? want : 1;
}

}  // namespace

int corr_plan_create(CorrPlan *p, size_t NF, cudaStream_t st, uint64_t *launches) {
    if (NF == 0) return 1;
    p->NF = NF;
    p->in_sm = false;
    if (2 * NF - 1 <= 4096 && !getenv("SASSENA_DSP_FOURSTEP")) {
        // the padded timeline fits one SM: single-kernel DSP (selffused.cu), see CorrPlan::in_sm
        int rc = self_plan_create(&p->sm, NF, st, launches);
        if (rc) return rc;
        if (p->sm.R == 1) {
            p->in_sm = true;
            p->L = p->sm.L;
            p->d_tw = p->sm.d_tw;  // aliases: the context tests d_tw to see whether a plan exists
            p->d_w = p->sm.d_w;
            return 0;
        }
        self_plan_destroy(&p->sm);
    }
    int log2L = ilog2_ceil(2 * NF);
    if (log2L < 1) log2L = 1;
    if (log2L > 2 * MAX_LOG2_SUB) return 1;  // NF > 2^21 frames not supported by the two-pass plan
    p->L = (size_t)1 << log2L;
    p->log2N2 = (log2L + 1) / 2;
    p->log2N1 = log2L - p->log2N2;
    p->Nmax = (size_t)1 << std::max(p->log2N1, p->log2N2);
    if (cudaMalloc(&p->d_tw, sizeof(double2) * std::max<size_t>(p->Nmax / 2, 1)) != cudaSuccess) return 2;
    if (cudaMalloc(&p->d_w, sizeof(double2) * p->L) != cudaSuccess) return 2;
    twiddle_table_kernel<<<(unsigned)((p->Nmax / 2 + 255) / 256 + 1), 256, 0, st>>>(p->d_tw, p->Nmax);
    // weights: What = unnormalised inverse DFT (sign +1) of w[tau] = 1/(NF-tau), stored in internal order
    double2 *Y = nullptr;
    if (cudaMalloc(&Y, sizeof(double2) * p->L) != cudaSuccess) return 2;
    int n = 1;
    n += run_colfft<LOAD_WEIGHTS>(p, nullptr, 0, 1, +1, Y, st);
    const size_t smem = sizeof(double2) << p->log2N2;
    cudaFuncSetAttribute(corr_rowfft_kernel<OUT_INTERNAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    corr_rowfft_kernel<OUT_INTERNAL><<<dim3(1u << p->log2N1, 1), FFT_THREADS, smem, st>>>(
        Y, 1, 1, NF, p->log2N1, p->log2N2, p->d_tw, ilog2_ceil(p->Nmax), +1, nullptr, nullptr, nullptr, p->d_w, 1.0, 0);
    n++;
    cudaStreamSynchronize(st);
    cudaFree(Y);
    if (launches) *launches += n;
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

void corr_plan_destroy(CorrPlan *p) {
    if (p->in_sm) {
        self_plan_destroy(&p->sm);
        p->in_sm = false;
        p->d_tw = nullptr;
        p->d_w = nullptr;
        p->NF = p->L = 0;
        return;
    }
    if (p->d_tw) cudaFree(p->d_tw);
    if (p->d_w) cudaFree(p->d_w);
    p->d_tw = nullptr;
    p->d_w = nullptr;
    p->NF = p->L = 0;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t corr_work_bytes(const CorrPlan *p, size_t nt) {
    if (p->in_sm) return std::max(self_loaded_work_bytes(&p->sm, nt), self_work_bytes(&p->sm, 1));
    const size_t N1 = (size_t)1 << p->log2N1;
    const size_t chunks = pick_chunks(p, nt);
    return align256(nt * p->L * sizeof(double2)) + align256(chunks * p->L * sizeof(double)) +
           align256(nt * N1 * sizeof(double2)) + align256(nt * sizeof(double2));
}

int corr_power_accumulate(const CorrPlan *p, const double2 *d_A, size_t ldA, size_t nt, void *d_work, double *d_P,
                          double *d_acc, cudaStream_t st) {
    if (nt == 0) return 0;
    if (p->in_sm) return self_power_accumulate_loaded(&p->sm, d_A, ldA, nt, d_work, d_P, d_acc, st);
    const size_t N1 = (size_t)1 << p->log2N1;
    const size_t chunks = pick_chunks(p, nt);
    char *w = reinterpret_cast<char *>(d_work);
    double2 *Y = reinterpret_cast<double2 *>(w);
    w += align256(nt * p->L * sizeof(double2));
    double *Ppart = reinterpret_cast<double *>(w);
    w += align256(chunks * p->L * sizeof(double));
    double2 *a_part = reinterpret_cast<double2 *>(w);
    w += align256(nt * N1 * sizeof(double2));
    double2 *a_tl = reinterpret_cast<double2 *>(w);

    int n = run_colfft<LOAD_TIMELINE>(p, d_A, ldA, nt, -1, Y, st);
    const size_t per_chunk = (nt + chunks - 1) / chunks;
    const size_t smem = sizeof(double2) << p->log2N2;
    cudaFuncSetAttribute(corr_rowfft_kernel<OUT_POWER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    corr_rowfft_kernel<OUT_POWER><<<dim3((unsigned)N1, (unsigned)chunks), FFT_THREADS, smem, st>>>(
        Y, nt, per_chunk, p->NF, p->log2N1, p->log2N2, p->d_tw, ilog2_ceil(p->Nmax), -1, p->d_w, Ppart, a_part, nullptr,
        1.0, 0);
    reduce_ppart_kernel<<<(unsigned)((p->L + 255) / 256), 256, 0, st>>>(Ppart, chunks, p->L, d_P);
    const double norm = 1.0 / ((double)p->NF * (double)p->L);
    reduce_apart_kernel<<<(unsigned)((nt * 32 + 255) / 256), 256, 0, st>>>(a_part, nt, N1, norm, a_tl);
    reduce_atl_kernel<<<1, 1024, 0, st>>>(a_tl, nt, d_acc);
    return n + 4;
}

int corr_finalize(const CorrPlan *p, const double *d_P, void *d_work, double2 *d_out, double scale, int conj_out,
                  cudaStream_t st) {
    if (p->in_sm) return self_finalize(&p->sm, d_P, d_work, d_out, scale, conj_out, st);
    double2 *Y = reinterpret_cast<double2 *>(d_work);
    int n = run_colfft<LOAD_REALPERM>(p, d_P, p->L, 1, +1, Y, st);
    const size_t smem = sizeof(double2) << p->log2N2;
    cudaFuncSetAttribute(corr_rowfft_kernel<OUT_FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    corr_rowfft_kernel<OUT_FINAL><<<dim3(1u << p->log2N1, 1), FFT_THREADS, smem, st>>>(
        Y, 1, 1, p->NF, p->log2N1, p->log2N2, p->d_tw, ilog2_ceil(p->Nmax), +1, nullptr, nullptr, nullptr, d_out, scale,
        conj_out);
    return n + 1;
}

size_t dsp_elementwise_work_bytes(size_t nt) { return align256(nt * sizeof(double2)); }

int dsp_elementwise_accumulate(const double2 *d_A, size_t ldA, size_t nt, size_t NF, int square, double2 *d_at,
                               double *d_acc, void *d_work, cudaStream_t st) {
    if (nt == 0) return 0;
    double2 *a_tl = reinterpret_cast<double2 *>(d_work);
    dsp_elementwise_at_kernel<<<(unsigned)((NF + 127) / 128), 128, 0, st>>>(d_A, ldA, nt, NF, square, d_at);
    int n = 1;
    for (size_t done = 0; done < nt;) {
        size_t cnt = std::min<size_t>(nt - done, 0x7fffffff);
        dsp_elementwise_mean_kernel<<<(unsigned)cnt, 256, 0, st>>>(d_A + done * ldA, ldA, NF, square, a_tl + done);
        done += cnt;
        n++;
    }
    reduce_atl_kernel<<<1, 1024, 0, st>>>(a_tl, nt, d_acc);
    return n + 1;
}

int launch_scale_complex(const double2 *d_in, double2 *d_out, size_t n, double scale, cudaStream_t st) {
    if (n == 0) return 0;
    scale_complex_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, d_out, n, scale);
    return 1;
}

}  // namespace sass
