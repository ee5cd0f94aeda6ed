// scan_sym.cu — K1s, coherent amplitudes of a |q| scan in the symmetric (Chebyshev) form.
//
// A Sassena run evaluates the SAME orientation vectors at many |q| (scattering.vectors.scans, reference
// parameters.cpp:1125-1189; init_subvectors scales unit vectors by |q|, abstract_vectors_scatter_device.cpp:96-175), so the
// q-vectors of the runner loop (abstract_scatter_device.cpp:162-173) are q_{n,m} = s_n v_m and, for equally spaced
// s_n = s_c + (n-K) ds, n = 0..2K, the phases of one (atom, direction) pair are symmetric about the centre |q|:
//
//     b exp(i s_{K+-k} sigma) = z0 (C_k +- i S_k),   z0 = b exp(i s_c sigma),  C_k = cos(k ds sigma),  S_k = sin(k ds sigma)
//
// with sigma = v.r.  C_k and S_k are REAL and obey the Chebyshev recurrence X_{k+1} = 2 cos(ds sigma) X_k - X_{k-1}, so one
// pair of |q| values costs two real recurrence steps and four real accumulations
//
//     P_k += C_k z0,   Q_k += S_k z0      (complex P_k, Q_k;  A_{K+k} = P_k + i Q_k,  A_{K-k} = P_k - i Q_k)
//
// = 3 FP64 instructions per evaluation (the complex three-term recurrence on z_n itself needs 4; one sincos per
// evaluation, as the reference does at all_vectors_scatter_device.cpp:433-434, needs 21).  Per (atom, direction, pass) the
// set-up is the dot product and two sincos.  The recurrences run k <= K <= 14 steps from exact seeds, so a rounding error
// grows to at most ~K^2 ulp (2e-14).
//
// Float-rounded scans (CORR): the reference computes the scan fractions in float (parameters.cpp:1151), so real scans are
// s_n + e_n with |e_n| ~ 3e-8 (to - from).  exp(i (s_n + e_n) sigma) = z_n (1 + i th - th^2/2 + O(th^3)), th = kappa_n sigma:
// the first-order term needs D_n = sum b sigma z_n — the same P/Q accumulation with w0 = sigma z0 in place of z0 (2 more
// FP64 instructions per evaluation) — and the second-order term E_n = sum b sigma^2 z_n carries a weight ~1e-10, so it
// runs in FP32 as PACKED f32x2 instructions (sm_100 FFMA2): one for the (C_k, S_k) recurrence and two for the four
// accumulations of a pair.  The FP32 recurrence is written on y_k = sg_k X_k with sg = + + - - + + ..., which turns the
// subtraction into y_{k+1} = +-c2 y_k + y_{k-1} (FFMA2 has no negated operands); the signs are undone at the end.
// CORR = 2 also takes D_n from the FP32 recurrence (two more FFMA2 per pair instead of four DFMAs): the inner loop is then
// the plain kernel's 3 FP64 instructions per evaluation, and with D in packed floats the registers allow K = 12 (25 |q| per
// pass, two passes for a 50-point scan instead of three).  Its error is bounded by theta_max (K^2 ulp32 + accumulation),
// theta_max = max |kappa_n sigma|; plan_scan (sgpu_capi.cu) takes it only where that bound stays below 5e-10.
//
// Mapping (as the general kernel K1, amplitude.cu): one CTA = one frame x WARPS directions, one direction per warp, lanes
// stride over the atoms of a TILE-atom tile; tiles stream through a STAGES-deep shared-memory ring filled by 1-D TMA bulk
// copies (cp.async.bulk, SASS UBLKCP) tracked with full/empty mbarriers; warp-shuffle tree at the end.
#include "kernels.hpp"
#include "ptx.cuh"
#include "sincos_qt.cuh"

#include <algorithm>

namespace sass {

namespace {

template <int K, int WARPS, int TILE, int STAGES, int CORR>
__global__ void __launch_bounds__(WARPS * 32, 1) amplitude_scan_sym_kernel(
    const float *__restrict__ xyz, const double *__restrict__ b, const double *__restrict__ vs, double sc, double ds,
    double2 *__restrict__ A, size_t ldA, size_t strideQ, int NA, int NM, int nq_valid, unsigned ngroups, size_t f0,
    int use_bulk, const ScanKappa kap) {
    constexpr int B = 2 * K + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_xyz = reinterpret_cast<float *>(smem_raw);                                  // [STAGES][TILE*3]
    double *s_b = reinterpret_cast<double *>(smem_raw + (size_t)STAGES * TILE * 3 * 4);  // [STAGES][TILE]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * TILE * (3 * 4 + 8));
    uint64_t *empty = full + STAGES;

    const unsigned group = blockIdx.x % ngroups;
    const size_t frame = f0 + blockIdx.x / ngroups;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = group * WARPS + warp;
    const float *p = xyz + frame * (size_t)NA * 3;
    const int ntiles = (NA + TILE - 1) / TILE;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full[s], use_bulk ? 1u : (unsigned)(WARPS * 32));
            ptx::mbar_init(&empty[s], (unsigned)WARPS);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](int t) {  // called by every thread at a converged point
        const int s = t % STAGES;
        const int a0 = t * TILE;
        const int cnt = min(TILE, NA - a0);
        if (use_bulk) {
            if (tid == 0) {
                if (t >= STAGES) ptx::mbar_wait(&empty[s], (unsigned)((t / STAGES - 1) & 1));
                ptx::mbar_expect_tx(&full[s], (unsigned)cnt * 20u);
                ptx::bulk_g2s(s_xyz + (size_t)s * TILE * 3, p + (size_t)a0 * 3, (unsigned)cnt * 12u, &full[s]);
                ptx::bulk_g2s(s_b + (size_t)s * TILE, b + a0, (unsigned)cnt * 8u, &full[s]);
            }
        } else {
            if (t >= STAGES) ptx::mbar_wait(&empty[s], (unsigned)((t / STAGES - 1) & 1));
            float *dx = s_xyz + (size_t)s * TILE * 3;
            const float *sx = p + (size_t)a0 * 3;
            for (int i = tid; i < cnt * 3; i += WARPS * 32) ptx::cp_async4(dx + i, sx + i);
            float *db = reinterpret_cast<float *>(s_b + (size_t)s * TILE);
            const float *sb = reinterpret_cast<const float *>(b + a0);
            for (int i = tid; i < cnt * 2; i += WARPS * 32) ptx::cp_async4(db + i, sb + i);
            ptx::cp_async_mbar_arrive_noinc(&full[s]);
        }
    };
    for (int t = 0; t < STAGES - 1 && t < ntiles; t++) issue(t);

    const bool active = m0 < NM;
    const double vx = __ldg(&vs[3 * m0]), vy = __ldg(&vs[3 * m0 + 1]), vz = __ldg(&vs[3 * m0 + 2]);  // zero padded past NM

    // lane sums: centre term and the K pairs, for A (and D in FP64, E in FP32 as (P, Q) pairs per component)
    double a0r = 0.0, a0i = 0.0, pr[K], pi[K], qr[K], qi[K];
    double d0r = 0.0, d0i = 0.0, dpr[CORR == 1 ? K : 1], dpi[CORR == 1 ? K : 1], dqr[CORR == 1 ? K : 1], dqi[CORR == 1 ? K : 1];
    float2 e0 = make_float2(0.f, 0.f), epq_r[CORR ? K : 1], epq_i[CORR ? K : 1];
    float2 fd0 = make_float2(0.f, 0.f), dpqf_r[CORR == 2 ? K : 1], dpqf_i[CORR == 2 ? K : 1];  // D in FP32, as (P, Q) pairs
#pragma unroll
    for (int k = 0; k < K; k++) pr[k] = pi[k] = qr[k] = qi[k] = 0.0;
    if (CORR) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (CORR == 1) dpr[k] = dpi[k] = dqr[k] = dqi[k] = 0.0;
            if (CORR == 2) dpqf_r[k] = dpqf_i[k] = make_float2(0.f, 0.f);
            epq_r[k] = epq_i[k] = make_float2(0.f, 0.f);
        }
    }

    for (int t = 0; t < ntiles; t++) {
        if (t + STAGES - 1 < ntiles) issue(t + STAGES - 1);
        const int s = t % STAGES;
        const int cnt = min(TILE, NA - t * TILE);
        ptx::mbar_wait(&full[s], (unsigned)((t / STAGES) & 1));
        if (active) {
            const float *sx = s_xyz + (size_t)s * TILE * 3;
            const double *sb = s_b + (size_t)s * TILE;
#pragma unroll 1
            for (int j = lane; j < cnt; j += 32) {
                const double x = (double)sx[3 * j], y = (double)sx[3 * j + 1], z = (double)sx[3 * j + 2];
                const double bj = sb[j];
                const double sigma = fma(z, vz, fma(y, vy, x * vx));  // quarter turns per unit |q|
                double sn, cs, s1, c1;
                sincos_qt(sc * sigma, sn, cs);
                sincos_qt(ds * sigma, s1, c1);
                const double zr = bj * cs, zi = bj * sn;  // z0
                a0r += zr;
                a0i += zi;
                const double c2 = c1 + c1;
                double Cp = 1.0, C = c1, Sp = 0.0, S = s1;
                double wr = 0.0, wi = 0.0;
                float2 yk = make_float2(0.f, 0.f), ykp = make_float2(0.f, 0.f), er_b = yk, ei_b = yk, c2p = yk, c2m = yk;
                float2 dr_b = yk, di_b = yk;
                if (CORR) {
                    const float fs = (float)sigma, fs2 = fs * fs, fzr = (float)zr, fzi = (float)zi;
                    if (CORR == 1) {
                        wr = sigma * zr;  // w0 = sigma z0
                        wi = sigma * zi;
                        d0r += wr;
                        d0i += wi;
                    } else {
                        const float fwr = fs * fzr, fwi = fs * fzi;
                        fd0.x += fwr;
                        fd0.y += fwi;
                        dr_b = make_float2(fwr, fwr);
                        di_b = make_float2(fwi, fwi);
                    }
                    const float fer = fs2 * fzr, fei = fs2 * fzi;  // sigma^2 z0
                    e0.x += fer;
                    e0.y += fei;
                    er_b = make_float2(fer, fer);
                    ei_b = make_float2(fei, fei);
                    const float fc2 = (float)c2;
                    c2p = make_float2(fc2, fc2);
                    c2m = make_float2(-fc2, -fc2);
                    yk = make_float2((float)c1, (float)s1);  // (C_1, S_1), sign +
                    ykp = make_float2(1.f, 0.f);             // (C_0, S_0), sign +
                }
#pragma unroll
                for (int k = 1; k <= K; k++) {
                    if (k >= 2) {
                        const double Cn = fma(c2, C, -Cp);
                        const double Sn = fma(c2, S, -Sp);
                        Cp = C;
                        C = Cn;
                        Sp = S;
                        S = Sn;
                        if (CORR) {
                            // y_k = rho_{k-1} c2 y_{k-1} + y_{k-2}, rho_{k-1} = (-1)^(k-1)
                            const float2 yn = __ffma2_rn((k & 1) ? c2p : c2m, yk, ykp);
                            ykp = yk;
                            yk = yn;
                        }
                    }
                    // operand order chosen so that consecutive DFMAs share one source (register reuse cache): a DFMA
                    // with three register-file reads costs 3 issue cycles, with two it costs 2 (tools/micro)
                    pr[k - 1] = fma(C, zr, pr[k - 1]);
                    pi[k - 1] = fma(C, zi, pi[k - 1]);
                    qi[k - 1] = fma(S, zi, qi[k - 1]);
                    qr[k - 1] = fma(S, zr, qr[k - 1]);
                    if (CORR == 1) {
                        dqr[k - 1] = fma(S, wr, dqr[k - 1]);
                        dqi[k - 1] = fma(S, wi, dqi[k - 1]);
                        dpi[k - 1] = fma(C, wi, dpi[k - 1]);
                        dpr[k - 1] = fma(C, wr, dpr[k - 1]);
                    }
                    if (CORR == 2) {
                        dpqf_r[k - 1] = __ffma2_rn(yk, dr_b, dpqf_r[k - 1]);
                        dpqf_i[k - 1] = __ffma2_rn(yk, di_b, dpqf_i[k - 1]);
                    }
                    if (CORR) {
                        epq_r[k - 1] = __ffma2_rn(yk, er_b, epq_r[k - 1]);  // (P, Q) of the real part, sign sg_k
                        epq_i[k - 1] = __ffma2_rn(yk, ei_b, epq_i[k - 1]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);
    }
    if (!active) return;

    // lane sums of the B |q| values: n = K is the centre, n = K +- k the pairs
    double re[B], im[B];
    re[K] = a0r;
    im[K] = a0i;
#pragma unroll
    for (int k = 1; k <= K; k++) {
        re[K + k] = pr[k - 1] - qi[k - 1];
        im[K + k] = pi[k - 1] + qr[k - 1];
        re[K - k] = pr[k - 1] + qi[k - 1];
        im[K - k] = pi[k - 1] - qr[k - 1];
    }
    if (CORR) {
        // fold the corrections in: A_n + i kappa_n D_n - (kappa_n^2 / 2) E_n
        double dre[B], dim[B], ere[B], eim[B];
        dre[K] = CORR == 1 ? d0r : (double)fd0.x;
        dim[K] = CORR == 1 ? d0i : (double)fd0.y;
        ere[K] = (double)e0.x;
        eim[K] = (double)e0.y;
#pragma unroll
        for (int k = 1; k <= K; k++) {
            const double sg = ((k & 3) >= 2) ? -1.0 : 1.0;  // sign carried by the FP32 recurrence
            double Dpr, Dpi, Dqr, Dqi;
            if (CORR == 1) {
                Dpr = dpr[k - 1], Dpi = dpi[k - 1], Dqr = dqr[k - 1], Dqi = dqi[k - 1];
            } else {
                Dpr = sg * (double)dpqf_r[k - 1].x, Dqr = sg * (double)dpqf_r[k - 1].y;
                Dpi = sg * (double)dpqf_i[k - 1].x, Dqi = sg * (double)dpqf_i[k - 1].y;
            }
            dre[K + k] = Dpr - Dqi;
            dim[K + k] = Dpi + Dqr;
            dre[K - k] = Dpr + Dqi;
            dim[K - k] = Dpi - Dqr;
            const double epr = sg * (double)epq_r[k - 1].x, eqr = sg * (double)epq_r[k - 1].y;
            const double epi = sg * (double)epq_i[k - 1].x, eqi = sg * (double)epq_i[k - 1].y;
            ere[K + k] = epr - eqi;
            eim[K + k] = epi + eqr;
            ere[K - k] = epr + eqi;
            eim[K - k] = epi - eqr;
        }
#pragma unroll
        for (int n = 0; n < B; n++) {
            const double kn = kap.k[n], hk = 0.5 * kn * kn;
            re[n] = fma(-hk, ere[n], fma(-kn, dim[n], re[n]));
            im[n] = fma(-hk, eim[n], fma(kn, dre[n], im[n]));
        }
    }
#pragma unroll
    for (int n = 0; n < B; n++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            re[n] += __shfl_xor_sync(0xffffffffu, re[n], o);
            im[n] += __shfl_xor_sync(0xffffffffu, im[n], o);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < B; n++)
            if (n < nq_valid) A[(size_t)n * strideQ + (size_t)m0 * ldA + frame] = make_double2(re[n], im[n]);
    }
}

constexpr int SYM_TILE = 512, SYM_STAGES = 4;
constexpr int SYM_WARPS_PLAIN = 12, SYM_WARPS_CORR = 8;

struct SymArgs {
    const float *d_xyz;
    const double *d_b, *d_vs;
    double sc, ds;
    int valid;
    double2 *A;
    size_t ldA, strideQ, NA, NM, f0, nf;
    cudaStream_t st;
    ScanKappa kap;
};

template <int K, int WARPS, int CORR>
int launch_sym(const SymArgs &a) {
    const unsigned ngroups = (unsigned)((a.NM + WARPS - 1) / WARPS);
    const size_t smem = (size_t)SYM_STAGES * SYM_TILE * 20 + 2 * SYM_STAGES * sizeof(uint64_t);
    auto kern = amplitude_scan_sym_kernel<K, WARPS, SYM_TILE, SYM_STAGES, CORR>;
    // the attribute is per device: set it on every launch (a few hundred ns) rather than once per process
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    const int use_bulk = (a.NA % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.d_xyz) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(a.d_b) & 15) == 0);
    int launches = 0;
    const size_t max_frames = (size_t)0x7fffffff / ngroups;
    for (size_t done = 0; done < a.nf;) {
        const size_t cnt = a.nf - done < max_frames ? a.nf - done : max_frames;
        kern<<<(unsigned)(cnt * ngroups), WARPS * 32, smem, a.st>>>(a.d_xyz, a.d_b, a.d_vs, a.sc, a.ds, a.A, a.ldA, a.strideQ,
                                                                    (int)a.NA, (int)a.NM, a.valid, ngroups, a.f0 + done,
                                                                    use_bulk, a.kap);
        launches++;
        done += cnt;
    }
    return launches;
}

template <int CORR, int WARPS>
int dispatch_sym(int K, const SymArgs &a) {
    if constexpr (CORR != 2) {
        switch (K) {
            case 1: return launch_sym<1, WARPS, CORR>(a);
            case 2: return launch_sym<2, WARPS, CORR>(a);
            case 3: return launch_sym<3, WARPS, CORR>(a);
            case 4: return launch_sym<4, WARPS, CORR>(a);
            case 5: return launch_sym<5, WARPS, CORR>(a);
            case 6: return launch_sym<6, WARPS, CORR>(a);
            case 7: return launch_sym<7, WARPS, CORR>(a);
            case 8: return launch_sym<8, WARPS, CORR>(a);
            default: break;
        }
    }
    if constexpr (CORR == 2) {
        switch (K) {
            case 9: return launch_sym<9, WARPS, 2>(a);
            case 10: return launch_sym<10, WARPS, 2>(a);
            case 11: return launch_sym<11, WARPS, 2>(a);
            case 12: return launch_sym<12, WARPS, 2>(a);
            default: break;
        }
    }
    if constexpr (!CORR) {
        switch (K) {
            case 9: return launch_sym<9, WARPS, 0>(a);
            case 10: return launch_sym<10, WARPS, 0>(a);
            case 11: return launch_sym<11, WARPS, 0>(a);
            case 12: return launch_sym<12, WARPS, 0>(a);
            case 13: return launch_sym<13, WARPS, 0>(a);
            case 14: return launch_sym<14, WARPS, 0>(a);
            default: break;
        }
    }
    return -1;
}

}  // namespace

// corrected: 1 = first-order sums in FP64 (K <= 8), 2 = in FP32 (K <= 12; plan_scan checks the error bound)
int amplitude_scan_sym_max_pass(int corrected) { return corrected == 2 ? 25 : corrected ? 17 : 29; }
int amplitude_scan_sym_qpad() { return 24; }  // a multiple of the directions per CTA (8 or 12 warps)

int launch_amplitude_scan_sym_pass(const float *d_xyz, const double *d_b, const double *d_vs, double s0, double ds, int nq,
                                   const double *kappa, double2 *d_A, size_t ldA, size_t strideQ, size_t NA, size_t NM,
                                   size_t f0, size_t nf, cudaStream_t st) {
    if (nf == 0 || NM == 0 || nq <= 0) return 0;
    const int K = std::max(1, nq / 2);  // 2K+1 >= nq; an even nq leaves the top slot masked
    SymArgs a{d_xyz, d_b, d_vs, s0 + (double)K * ds, ds, nq, d_A, ldA, strideQ, NA, NM, f0, nf, st, ScanKappa()};
    if (kappa) {
        for (int n = 0; n < 32; n++) a.kap.k[n] = n < nq ? kappa[n] : 0.0;
        // passes longer than the FP64-D kernel takes are the FP32-D kernel's (plan_scan makes them only within its error bound)
        return K <= 8 ? dispatch_sym<1, SYM_WARPS_CORR>(K, a) : dispatch_sym<2, SYM_WARPS_CORR>(K, a);
    }
    return dispatch_sym<0, SYM_WARPS_PLAIN>(K, a);
}

}  // namespace sass
