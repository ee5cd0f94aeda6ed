// multipole.cu — multipole-sphere amplitude kernel (K5).
//
// Restates MPSphereScatterDevice::scatter (reference src/scatter_devices/multipole_scatter_device.cpp:428-498):
//     A_lm[f] = sum_j 4 pi i^l b_j j_l(q r_j) conj(Y_lm(theta_j, phi_j)),   coordinates (r, phi, theta) as float.
// The reference evaluates boost::math::sph_bessel and spherical_harmonic once per (moment, atom, frame);
// here one pass over the atoms of a frame produces ALL moments:
//   * j_l(x), l = 0..lmax, by one recurrence per atom (series for x<1, Miller downward recurrence for
//     x < lmax, upward recurrence otherwise) kept in shared memory;
//   * the normalised associated Legendre functions by the standard (m, l) recurrence with host-made
//     coefficient tables, and cos(m phi), sin(m phi) by rotation — each term costs a few FP64 ops;
//   * only m >= 0 is evaluated: with U = sum b j_l P_lm cos(m phi), V = sum b j_l P_lm sin(m phi),
//       A_{l,+m} = 4 pi i^l (U - iV),   A_{l,-m} = 4 pi i^l (-1)^m (U + iV)       (Y_{l,-m} = (-1)^m conj Y_{lm}).
// Reduction over atoms: warp shuffle tree per (l,m), accumulated in per-warp shared memory, combined in
// a fixed order (no atomics).
#include "kernels.hpp"
#include <cstdlib>

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

namespace sass {

namespace {

constexpr int MP_THREADS = 128;
constexpr int MP_WARPS = MP_THREADS / 32;
constexpr int MP_SERIES_TERMS = 10;
constexpr int MP_LMAX = 50;

struct MpTables {
    int lmax = -1;
    double *d = nullptr;  // [cMM (lmax+1)] [cM1 (lmax+1)] [A npairs] [B npairs] [series (lmax+1)*TERMS]
};
MpTables g_tab[16];  // per device

__host__ __device__ inline int mp_npairs(int lmax) { return (lmax + 1) * (lmax + 2) / 2; }
__host__ __device__ inline int mp_pair(int lmax, int l, int m) { return m * (lmax + 1) - (m * (m - 1)) / 2 + (l - m); }
// l-major order used by the batched kernel: a warp of consecutive pairs shares (nearly) one l, so its Bessel operands
// are a shared-memory broadcast
__host__ __device__ inline int mp_pair_l(int l, int m) { return (l * (l + 1)) / 2 + m; }

__global__ void __launch_bounds__(MP_THREADS) multipole_sphere_kernel(
    const float *__restrict__ sph, const double *__restrict__ b, double ql, int lmax, int lstart, size_t NA,
    size_t f0, int nsplit, size_t atoms_per_split, const double *__restrict__ tab, double2 *__restrict__ part,
    size_t nf) {
    extern __shared__ double smem[];
    const int npairs = mp_npairs(lmax);
    double *sJ = smem;                                 // [(lmax+1)][MP_THREADS]
    double *sAcc = smem + (lmax + 1) * MP_THREADS;     // [MP_WARPS][npairs*2]
    const double *cMM = tab;
    const double *cM1 = tab + (lmax + 1);
    const double *cA = tab + 2 * (lmax + 1);
    const double *cB = cA + npairs;
    const double *cS = cB + npairs;  // series coefficients [(lmax+1)][TERMS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t fr = blockIdx.x;
    const int split = blockIdx.y;
    for (int i = tid; i < MP_WARPS * npairs * 2; i += MP_THREADS) sAcc[i] = 0.0;
    __syncthreads();

    const size_t a_begin = (size_t)split * atoms_per_split;
    const size_t a_end = min(NA, a_begin + atoms_per_split);
    const float *p = sph + (f0 + fr) * NA * 3;
    double *myAcc = sAcc + warp * npairs * 2;

    for (size_t base = a_begin; base < a_end; base += MP_THREADS) {
        const size_t atom = base + tid;
        double r = 0.0, phi = 0.0, theta = 0.0, bj = 0.0;
        if (atom < a_end) {
            r = (double)__ldg(&p[3 * atom]);
            phi = (double)__ldg(&p[3 * atom + 1]);
            theta = (double)__ldg(&p[3 * atom + 2]);
            bj = __ldg(&b[atom]);
        }
        const double x = ql * r;
        // ---- spherical Bessel ladder into sJ[l][tid] ----
        if (x < 1.0) {
            const double x2 = x * x;
            double pref = 1.0;
            for (int l = 0; l <= lmax; l++) {
                if (l > 0) pref *= x / (double)(2 * l + 1);
                double term = 1.0, sum = 1.0;
#pragma unroll
                for (int k = 0; k < MP_SERIES_TERMS; k++) {
                    term *= x2 * __ldg(&cS[l * MP_SERIES_TERMS + k]);
                    sum += term;
                }
                sJ[l * MP_THREADS + tid] = pref * sum;
            }
        } else {
            double sn, cs;
            sincos(x, &sn, &cs);
            const double invx = 1.0 / x;
            const double j0 = sn * invx;
            const double j1 = (sn * invx - cs) * invx;
            if (x >= (double)lmax) {
                double jm = j0, jc = j1;
                sJ[tid] = j0;
                if (lmax >= 1) sJ[MP_THREADS + tid] = j1;
                for (int l = 1; l < lmax; l++) {
                    const double jn = fma((double)(2 * l + 1) * invx, jc, -jm);
                    jm = jc;
                    jc = jn;
                    sJ[(l + 1) * MP_THREADS + tid] = jn;
                }
            } else {
                double jp = 0.0, jc = 1e-300;
                for (int k = lstart; k >= 1; k--) {
                    const double jm = fma((double)(2 * k + 1) * invx, jc, -jp);  // j_{k-1}
                    jp = jc;
                    jc = jm;
                    if (k - 1 <= lmax) sJ[(k - 1) * MP_THREADS + tid] = jc;
                }
                // jc ~ j_0, jp ~ j_1 (unnormalised)
                const double scale = (fabs(j0) >= fabs(j1)) ? j0 / jc : j1 / jp;
                for (int l = 0; l <= lmax; l++) sJ[l * MP_THREADS + tid] *= scale;
            }
        }
        // ---- Legendre / azimuth recurrences, one (l,m>=0) term at a time ----
        double st, ct, s1, c1;
        sincos(theta, &st, &ct);
        sincos(phi, &s1, &c1);
        double pmm = 0.28209479177387814347;  // sqrt(1/(4 pi))
        double cm = 1.0, sm = 0.0;
        for (int m = 0; m <= lmax; m++) {
            if (m > 0) {
                pmm *= __ldg(&cMM[m]) * st;
                const double cn = cm * c1 - sm * s1;
                sm = fma(sm, c1, cm * s1);
                cm = cn;
            }
            double p2 = 0.0, p1 = pmm;
            const int pbase = mp_pair(lmax, m, m);
            for (int l = m; l <= lmax; l++) {
                double pl;
                if (l == m) pl = pmm;
                else if (l == m + 1) pl = __ldg(&cM1[m]) * ct * pmm;
                else pl = __ldg(&cA[pbase + l - m]) * fma(ct, p1, -__ldg(&cB[pbase + l - m]) * p2);
                p2 = p1;
                p1 = pl;
                const double t = bj * sJ[l * MP_THREADS + tid] * pl;
                double u = t * cm, v = t * sm;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    u += __shfl_xor_sync(0xffffffffu, u, o);
                    v += __shfl_xor_sync(0xffffffffu, v, o);
                }
                if (lane == 0) {
                    myAcc[2 * (pbase + l - m)] += u;
                    myAcc[2 * (pbase + l - m) + 1] += v;
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < npairs; i += MP_THREADS) {
        double u = 0.0, v = 0.0;
#pragma unroll
        for (int w = 0; w < MP_WARPS; w++) {
            u += sAcc[w * npairs * 2 + 2 * i];
            v += sAcc[w * npairs * 2 + 2 * i + 1];
        }
        part[((size_t)split * nf + fr) * npairs + i] = make_double2(u, v);
    }
}

// A[mom][f0+fr] = 4 pi i^l * { U - iV  (m>=0) ; (-1)^m (U + iV) (m<0) }, summed over the atom splits
__global__ void multipole_assemble_kernel(const double2 *__restrict__ part, int nsplit, size_t nf, int lmax,
                                          const int *__restrict__ lm, size_t NM, double2 *__restrict__ A, size_t ldA,
                                          size_t f0) {
    const size_t fr = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t mom = blockIdx.y;
    if (fr >= nf || mom >= NM) return;
    const int l = lm[2 * mom], m = lm[2 * mom + 1];
    const int am = m < 0 ? -m : m;
    const int npairs = mp_npairs(lmax);
    const int pi = mp_pair(lmax, l, am);
    double u = 0.0, v = 0.0;
    for (int s = 0; s < nsplit; s++) {
        const double2 w = part[((size_t)s * nf + fr) * npairs + pi];
        u += w.x;
        v += w.y;
    }
    double wr, wi;
    if (m >= 0) {
        wr = u;
        wi = -v;
    } else {
        const double sg = (am & 1) ? -1.0 : 1.0;
        wr = sg * u;
        wi = sg * v;
    }
    const double FOURPI = 12.566370614359172954;
    double ar, ai;
    switch (l & 3) {
        case 0: ar = wr; ai = wi; break;
        case 1: ar = -wi; ai = wr; break;
        case 2: ar = -wr; ai = -wi; break;
        default: ar = wi; ai = -wr; break;
    }
    A[mom * ldA + f0 + fr] = make_double2(FOURPI * ar, FOURPI * ai);
}

// ---- K5, batched "GEMM" form -------------------------------------------------------------------------------------
// U[q][l,m] = sum_a B[q][l][a] * Yc[l,m][a],  V likewise with Ys, where B = b_a(q) j_l(q r_a) and
// (Yc, Ys) = Pbar_lm(theta_a) (cos m phi_a, sin m phi_a).  Y does not depend on |q| and B does not depend on m, so for a
// tile of atoms both tables are built once in shared memory (Y once per batch of Q |q| values) and the sum over atoms
// becomes a register-blocked product: a thread owns one (l,m) pair and 2Q accumulators, walks the atoms of its half of
// the tile and does 2Q DFMAs per atom from one LDS.128 (Y) and Q LDS.64 (B).  No cross-lane reduction exists any
// more: the sum over atoms is sequential inside the owning thread, across tiles, in a fixed order.
// Per tile: geometry (sincos) -> table tasks (Legendre column pairs (m, lmax-m): equal length; Bessel ladders with an
// x-dependent Miller start) spread over all 512 threads -> product phase on 2 x NP threads.
// Two CTAs of 256 threads per SM (128 registers each, half the shared memory each) rather than one of 512: the table phase
// of a tile is a bundle of serial chains (latency bound), the product phase streams shared memory into DFMAs (throughput
// bound) -- two independent CTAs let one phase run under the other instead of alternating behind the same barriers.
// Measured on config 4 (pass of 8 |q|, 16 frames, largest tile that fits): 19.6 ms against 19.9 ms for one 512-thread CTA.
// 16 |q| per pass does not pay: 39.8 ms against 2 x 19.9 ms (the larger B table shrinks the tile and the accumulators of 16
// |q| leave no registers for loads in flight).
constexpr int MG_CTAS_PER_SM = 2;
constexpr int MG_THREADS = 512 / MG_CTAS_PER_SM;
constexpr int MG_GROUPS = MG_THREADS / 256;  // thread groups of the product phase, A / MG_GROUPS atoms of a tile each
constexpr size_t MG_SMEM_CAP = (MG_CTAS_PER_SM == 1 ? 200 : 110) * 1024;
constexpr int MG_ROUNDS = 4;    // rounds of table tasks per tile at most (limits the tile size)
constexpr int MG_A_MAX = 64;    // atoms per tile: chosen per launch (mg_pick_tile) so that the (lmax+1+Q)*A table tasks fill
                                // whole rounds of the 512 threads (24 atoms at L = 20, Q = 8 left the second round 36 % full)

// one Legendre/azimuth column: col[pair_l(l,m)] = Pbar_lm(theta) (cos m phi, sin m phi), l = m..lmax, col = sY + a*NP.
// cAB[pair(lmax,m,m) + l - m] = (A_lm, B_lm) of the three-term recurrence; the pair index advances by l per step
// (pair_l(l,m) = l(l+1)/2 + m), the first two terms are peeled so that the loop body is branch-free.
// TWO atoms per call (geo = r, ct, st, c1, s1 per atom; columns col0, col1): the recurrence is a serial chain of ~25 cycles per
// step, two independent chains share the coefficient loads and the loop and fill each other's latency.
__device__ __forceinline__ void mg_column2(int m, int lmax, const double *__restrict__ g0, const double *__restrict__ g1,
                                           const double *__restrict__ cK, const double *__restrict__ cM1,
                                           const double2 *__restrict__ cAB, double2 *__restrict__ col0,
                                           double2 *__restrict__ col1) {
    const double ct0 = g0[1], st0 = g0[2], ct1 = g1[1], st1 = g1[2];
    // st^m and (c1 + i s1)^m by binary powering (short dependency chain)
    double pw0 = 1.0, b0 = st0, cm0 = 1.0, sm0 = 0.0, rc0 = g0[3], rs0 = g0[4];
    double pw1 = 1.0, b1 = st1, cm1 = 1.0, sm1 = 0.0, rc1 = g1[3], rs1 = g1[4];
    for (int e = m; e; e >>= 1) {
        if (e & 1) {
            pw0 *= b0;
            pw1 *= b1;
            const double t0 = cm0 * rc0 - sm0 * rs0, t1 = cm1 * rc1 - sm1 * rs1;
            sm0 = fma(sm0, rc0, cm0 * rs0);
            sm1 = fma(sm1, rc1, cm1 * rs1);
            cm0 = t0;
            cm1 = t1;
        }
        b0 *= b0;
        b1 *= b1;
        const double t0 = rc0 * rc0 - rs0 * rs0, t1 = rc1 * rc1 - rs1 * rs1;
        rs0 = 2.0 * rc0 * rs0;
        rs1 = 2.0 * rc1 * rs1;
        rc0 = t0;
        rc1 = t1;
    }
    const double k = cK[m];
    const double pmm0 = k * pw0, pmm1 = k * pw1;
    int idx = mp_pair_l(m, m);
    col0[idx] = make_double2(pmm0 * cm0, pmm0 * sm0);
    col1[idx] = make_double2(pmm1 * cm1, pmm1 * sm1);
    if (m == lmax) return;
    const double k1 = cM1[m];
    double p20 = pmm0, p10 = k1 * ct0 * pmm0, p21 = pmm1, p11 = k1 * ct1 * pmm1;
    idx += m + 1;
    col0[idx] = make_double2(p10 * cm0, p10 * sm0);
    col1[idx] = make_double2(p11 * cm1, p11 * sm1);
    const double2 *ab = cAB + (mp_pair(lmax, m, m) - m);  // ab[l]
    for (int l = m + 2; l <= lmax; l++) {
        const double2 c = ab[l];
        const double pl0 = c.x * fma(ct0, p10, -c.y * p20), pl1 = c.x * fma(ct1, p11, -c.y * p21);
        p20 = p10;
        p10 = pl0;
        p21 = p11;
        p11 = pl1;
        idx += l;
        col0[idx] = make_double2(pl0 * cm0, pl0 * sm0);
        col1[idx] = make_double2(pl1 * cm1, pl1 * sm1);
    }
}

// b * j_l(x), l = 0..lmax, into J[l * JS] (one Bessel ladder): 4-term series below x = 0.05, upward recurrence from (j0, j1) for
// x >= lmax, Miller's downward recurrence started at lmax + 8 + 1.5 x otherwise (measured sufficient for 1e-14)
__device__ __forceinline__ void mg_bessel1(double x, double bq, double *__restrict__ J, int JS, int lmax, int lstart_max,
                                           const double *__restrict__ cS) {
    if (x < 0.05) {
        // power series, 4 terms are exact to 1e-16 below x = 0.05
        const double x2 = x * x;
        double pref = bq;
        for (int l = 0; l <= lmax; l++) {
            if (l > 0) pref *= x / (double)(2 * l + 1);
            double term = 1.0, sum = 1.0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                term *= x2 * __ldg(&cS[l * MP_SERIES_TERMS + k]);
                sum += term;
            }
            J[l * JS] = pref * sum;
        }
        return;
    }
    double sn, cs;
    sincos(x, &sn, &cs);
    const double invx = 1.0 / x;
    const double j0 = sn * invx;
    const double j1 = (sn * invx - cs) * invx;
    // (2l+1)/x advances by 2/x per step: one DADD instead of an int->double conversion and a DMUL
    const double step = 2.0 * invx;
    if (x >= (double)lmax) {
        double jm = j0, jc = j1;
        J[0] = bq * j0;
        if (lmax >= 1) J[JS] = bq * j1;
        double t = 3.0 * invx;
        for (int l = 1; l < lmax; l++) {
            const double jn = fma(t, jc, -jm);
            t += step;
            jm = jc;
            jc = jn;
            J[(l + 1) * JS] = bq * jn;
        }
    } else {
        // two loops: above lmax + 1 nothing is kept, below every value is
        const int lstart = min(lstart_max, lmax + 8 + (int)(1.5 * x));
        double jp = 0.0, jc = 1e-300;
        double t = (double)(2 * lstart + 1) * invx;
        int k = lstart;
        for (; k > lmax + 1; k--) {
            const double jm = fma(t, jc, -jp);
            t -= step;
            jp = jc;
            jc = jm;
        }
        for (; k >= 1; k--) {
            const double jm = fma(t, jc, -jp);
            t -= step;
            jp = jc;
            jc = jm;
            J[(k - 1) * JS] = jc;
        }
        const double scale = bq * ((fabs(j0) >= fabs(j1)) ? j0 / jc : j1 / jp);
        for (int l = 0; l <= lmax; l++) J[l * JS] *= scale;
    }
}

template <int Q>
__global__ void __launch_bounds__(MG_THREADS, MG_CTAS_PER_SM) multipole_gemm_kernel(
    const float *__restrict__ sph, const double *__restrict__ b, size_t b_stride, const double *__restrict__ qlens,
    int lmax, int lstart_max, size_t NA, size_t f0, size_t a_first, size_t a_last, size_t atoms_per_split,
    const double *__restrict__ tab, double2 *__restrict__ part, size_t nf, int q0, int NQ, int MG_A) {
    extern __shared__ double smem[];
    const int MG_APG = MG_A / MG_GROUPS;
    const int L1 = lmax + 1;
    const int NP = mp_npairs(lmax);
    double2 *sY = reinterpret_cast<double2 *>(smem);              // [MG_A][NP]
    double *sB = smem + (size_t)2 * MG_A * NP;                    // [MG_A][Q/2][L1][2]: pairs of |q|, l fastest (see below)
    const int BS = L1 * Q + 2;                                    // per-atom stride of sB, padded against bank conflicts
    double *sGeo = sB + (size_t)MG_A * BS;                        // [2][MG_A][5]: r, ct, st, c1, s1 (double-buffered)
    double *sTab = sGeo + 2 * MG_A * 5;                               // recurrence tables: cM1[L1] cK[L1] (cA, cB)[NP]
    int *sL = reinterpret_cast<int *>(sTab + 2 * L1 + 2 * NP);    // [NP]: l of pair p
    const double *cS = tab + 2 * L1 + 2 * NP;
    const double *cM1 = sTab, *cK = sTab + L1;
    const double2 *cAB = reinterpret_cast<const double2 *>(sTab + 2 * L1);  // (A, B) interleaved: one LDS.128 per step

    const int tid = threadIdx.x;
    // the Legendre recurrence is a serial chain: keep its coefficient tables in shared memory, not behind global loads
    for (int i = tid; i < L1; i += MG_THREADS) {
        sTab[i] = __ldg(&tab[L1 + i]);
        sTab[L1 + i] = __ldg(&tab[2 * L1 + 2 * NP + L1 * MP_SERIES_TERMS + i]);
    }
    for (int i = tid; i < NP; i += MG_THREADS) {
        sTab[2 * L1 + 2 * i] = __ldg(&tab[2 * L1 + i]);
        sTab[2 * L1 + 2 * i + 1] = __ldg(&tab[2 * L1 + NP + i]);
    }
    const size_t fr = blockIdx.x;
    const int split = blockIdx.y;
    for (int m = 0; m <= lmax; m++)
        for (int l = m + tid; l <= lmax; l += MG_THREADS) sL[mp_pair_l(l, m)] = l;
    const size_t a_begin = a_first + (size_t)split * atoms_per_split;
    const size_t a_end = min(a_last, a_begin + atoms_per_split);
    const float *p = sph + (f0 + fr) * NA * 3;
    const int NK = L1;  // one task per Legendre column m; short columns first so that the Bessel tasks that wrap
                        // around to a second round land on threads that finished early
    const int grp = tid >> 8, tp = tid & 255;

    // this thread's table tasks (the same for every tile).  Bessel ladders (one per (|q|, atom), the longest chains) come first,
    // then the Legendre columns, long ones first, TWO atoms per task; odd rounds run over the threads backwards so that the
    // threads that drew the short tasks of one round take the next round's.  row < NK: column m = row of atoms a and a + A/2
    // (consecutive lanes -> consecutive atoms: the column stores stay bank-conflict free);
    // row >= NK: ladder of |q| row - NK for atom a.
    const int AH = MG_A >> 1, nBes = Q * MG_A, nTask = nBes + NK * AH;
    int task_a[MG_ROUNDS], task_row[MG_ROUNDS];
#pragma unroll
    for (int k = 0; k < MG_ROUNDS; k++) {
        const int id = ((k & 1) ? (MG_THREADS - 1 - tid) : tid) + k * MG_THREADS;
        if (id < nBes) {
            task_row[k] = NK + id / MG_A;
            task_a[k] = id % MG_A;
        } else if (id < nTask) {
            task_row[k] = (id - nBes) / AH;
            task_a[k] = (id - nBes) % AH;
        } else {
            task_row[k] = -1;
            task_a[k] = 0;
        }
    }

    double2 acc[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) acc[q] = make_double2(0.0, 0.0);

    // geometry of a tile (r, cos/sin theta, cos/sin phi per atom) into sGeo[buf], by the threads (gi, gi+ng, ...)
    auto geometry = [&](size_t base, int buf, int gi, int ng) {
        double *g = sGeo + (size_t)buf * MG_A * 5;
        for (int i = gi; i < MG_A; i += ng) {
            const size_t atom = base + i;
            double r = 0.0, phi = 0.0, theta = 0.0;
            if (atom < a_end) {
                r = (double)__ldg(&p[3 * atom]);
                phi = (double)__ldg(&p[3 * atom + 1]);
                theta = (double)__ldg(&p[3 * atom + 2]);
            }
            double st, ct, s1, c1;
            sincos(theta, &st, &ct);
            sincos(phi, &s1, &c1);
            g[i * 5] = r;
            g[i * 5 + 1] = ct;
            g[i * 5 + 2] = st;
            g[i * 5 + 3] = c1;
            g[i * 5 + 4] = s1;
        }
    };
    // The product phase leaves 2 * (256 - NP) threads without a pair; they prepare the NEXT tile's geometry meanwhile
    // (double-buffered sGeo), so the serial sincos step and its barrier drop out of the tile loop.
    const int n_idle = MG_GROUPS * (256 - NP);
    const int idle_index = (tp >= NP) ? grp * (256 - NP) + (tp - NP) : -1;
    int cur = 0;
    if (a_begin < a_end) geometry(a_begin, 0, tid, MG_THREADS);

    for (size_t base = a_begin; base < a_end; base += MG_A, cur ^= 1) {
        const double *geo = sGeo + (size_t)cur * MG_A * 5;
        if (n_idle == 0 && base > a_begin) {  // (no spare threads: lmax with NP = 256 -- not reachable for lmax <= 21)
            __syncthreads();
            geometry(base, cur, tid, MG_THREADS);
        }
        __syncthreads();  // previous tile fully consumed; this tile's geometry is in place
        // table tasks: rows [0, NK): Legendre column m = lmax-row of atom a; rows [NK, NK+Q): Bessel ladder (q, a)
#pragma unroll
        for (int k = 0; k < MG_ROUNDS; k++) {
            if (task_row[k] < 0) break;
            const int a = task_a[k], row = task_row[k];
            if (row < NK) {
                mg_column2(row, lmax, geo + a * 5, geo + (a + AH) * 5, cK, cM1, cAB, sY + (size_t)a * NP,
                           sY + (size_t)(a + AH) * NP);
            } else {
                // B table of an atom: [Q/2][L1] pairs (q, q+1), l fastest -- the product phase's warp reads the pair of ONE or
                // two neighbouring l, i.e. one 128-byte line per LDS.128 (with q fastest the two l were 8 Q bytes apart: two
                // wavefronts per load, 14 per warp and atom against 8 SM-cycles of DFMA issue).  Element l of this |q| at J[l * JS].
                // (Two ladders of neighbouring |q| side by side in one task -- two chains per loop -- measured 5 % slower on
                // config 4: half as many, longer tasks lengthen the critical path of the table phase.)
                const int q = row - NK;
                const size_t atom = base + a;
                constexpr int JS = (Q >= 2) ? 2 : 1;
                double *J = sB + (size_t)a * BS + ((Q >= 2) ? (size_t)(q >> 1) * (2 * L1) + (q & 1) : 0);
                double bq = 0.0;
                if (atom < a_end && q0 + q < NQ) bq = __ldg(&b[(size_t)(q0 + q) * b_stride + atom]);
                const double ql = (q0 + q < NQ) ? __ldg(&qlens[q0 + q]) : 0.0;
                mg_bessel1(ql * geo[a * 5], bq, J, JS, lmax, lstart_max, cS);
            }
        }
        __syncthreads();
        if (idle_index >= 0 && base + MG_A < a_end) geometry(base + MG_A, cur ^ 1, idle_index, n_idle);
        if (tp < NP) {
            const int l = sL[tp];
            // (a hand-pipelined version -- operands of atom aa + 1 loaded before the DFMAs of atom aa -- measured 6 % slower
            // than what ptxas makes of the four-fold unrolled loop)
#pragma unroll 4
            for (int aa = 0; aa < MG_APG; aa++) {
                const int a = grp * MG_APG + aa;
                const double2 y = sY[(size_t)a * NP + tp];
                const double *Bp = sB + (size_t)a * BS + (size_t)l * ((Q >= 2) ? 2 : 1);  // pair (q, q+1) at Bp + q L1
                if (Q >= 2) {
#pragma unroll
                    for (int q = 0; q < Q; q += 2) {
                        const double2 b2 = *reinterpret_cast<const double2 *>(Bp + q * L1);
                        acc[q].x = fma(b2.x, y.x, acc[q].x);
                        acc[q].y = fma(b2.x, y.y, acc[q].y);
                        acc[q + (Q >= 2 ? 1 : 0)].x = fma(b2.y, y.x, acc[q + (Q >= 2 ? 1 : 0)].x);
                        acc[q + (Q >= 2 ? 1 : 0)].y = fma(b2.y, y.y, acc[q + (Q >= 2 ? 1 : 0)].y);
                    }
                } else {
                    acc[0].x = fma(Bp[0], y.x, acc[0].x);
                    acc[0].y = fma(Bp[0], y.y, acc[0].y);
                }
            }
        }
    }
    if (tp < NP) {
#pragma unroll
        for (int q = 0; q < Q; q++)
            if (q0 + q < NQ) part[((((size_t)split * MG_GROUPS + grp) * NQ + q0 + q) * nf + fr) * NP + tp] = acc[q];
    }
}

// A_q[mom][f0+fr] from part[split][q][fr][pair] (see multipole_assemble_kernel)
__global__ void multipole_assemble_batch_kernel(const double2 *__restrict__ part, int nsplit, size_t nf, int lmax,
                                                const int *__restrict__ lm, size_t NM, double2 *__restrict__ A,
                                                size_t ldA, size_t f0, int q, int NQ) {
    const size_t fr = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t mom = blockIdx.y;
    if (fr >= nf || mom >= NM) return;
    const int l = lm[2 * mom], m = lm[2 * mom + 1];
    const int am = m < 0 ? -m : m;
    const int NP = mp_npairs(lmax);
    const int pi = mp_pair_l(l, am);
    double u = 0.0, v = 0.0;
    for (int s = 0; s < nsplit; s++) {
        const double2 w = part[(((size_t)s * NQ + q) * nf + fr) * NP + pi];
        u += w.x;
        v += w.y;
    }
    double wr, wi;
    if (m >= 0) {
        wr = u;
        wi = -v;
    } else {
        const double sg = (am & 1) ? -1.0 : 1.0;
        wr = sg * u;
        wi = sg * v;
    }
    const double FOURPI = 12.566370614359172954;
    double ar, ai;
    switch (l & 3) {
        case 0: ar = wr; ai = wi; break;
        case 1: ar = -wi; ai = wr; break;
        case 2: ar = -wr; ai = -wi; break;
        default: ar = wi; ai = -wr; break;
    }
    A[mom * ldA + f0 + fr] = make_double2(FOURPI * ar, FOURPI * ai);
}

int mp_lstart(int lmax) { return lmax + 20 + (int)std::sqrt(60.0 * (lmax + 10)); }

const double *mp_tables(int lmax, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    MpTables &T = g_tab[dev & 15];
    if (T.lmax == lmax && T.d) return T.d;
    if (T.d) {
        cudaStreamSynchronize(st);
        cudaFree(T.d);
        T.d = nullptr;
    }
    const int np = mp_npairs(lmax);
    std::vector<double> h(2 * (lmax + 1) + 2 * np + (lmax + 1) * MP_SERIES_TERMS + (lmax + 1), 0.0);
    double *cMM = h.data(), *cM1 = cMM + (lmax + 1), *cA = cM1 + (lmax + 1), *cB = cA + np, *cS = cB + np;
    double *cK = cS + (lmax + 1) * MP_SERIES_TERMS;  // Pbar_mm = cK[m] * sin^m(theta)
    cK[0] = std::sqrt(1.0 / (4.0 * M_PI));
    for (int m = 1; m <= lmax; m++) cK[m] = cK[m - 1] * -std::sqrt((2.0 * m + 1.0) / (2.0 * m));
    for (int m = 0; m <= lmax; m++) {
        if (m > 0) cMM[m] = -std::sqrt((2.0 * m + 1.0) / (2.0 * m));
        cM1[m] = std::sqrt(2.0 * m + 3.0);
        for (int l = m + 2; l <= lmax; l++) {
            const double dl = l, dm = m;
            cA[mp_pair(lmax, l, m)] = std::sqrt((4.0 * dl * dl - 1.0) / (dl * dl - dm * dm));
            cB[mp_pair(lmax, l, m)] = std::sqrt(((dl - 1.0) * (dl - 1.0) - dm * dm) / (4.0 * (dl - 1.0) * (dl - 1.0) - 1.0));
        }
    }
    for (int l = 0; l <= lmax; l++)
        for (int k = 1; k <= MP_SERIES_TERMS; k++) cS[l * MP_SERIES_TERMS + (k - 1)] = -0.5 / (k * (2.0 * l + 2.0 * k + 1.0));
    if (cudaMalloc(reinterpret_cast<void **>(&T.d), h.size() * sizeof(double)) != cudaSuccess) return nullptr;
    cudaMemcpyAsync(T.d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    T.lmax = lmax;
    return T.d;
}

// Atom splits per frame: grid = nf x nsplit CTAs, one CTA per SM (512 threads, ~170 KB shared memory).  At least four waves
// of the 148 SMs when the atoms allow it, and the count whose last wave is fullest (600 CTAs = 4.05 waves ran as 5).
int mp_nsplit(size_t nf, size_t NA, size_t SMS = 148) {  // SMS: resident CTAs of the device (148 SMs x CTAs per SM)
    size_t cap = (NA + 4 * MP_THREADS - 1) / (4 * MP_THREADS);
    if (cap < 1) cap = 1;
    if (cap > 4096) cap = 4096;
    const size_t lo = std::min(cap, (4 * SMS + nf - 1) / nf);       // first count with >= 4 waves (or all the atoms allow)
    const size_t hi = std::min(cap, std::max(lo, (8 * SMS + nf - 1) / nf));
    size_t best = lo;
    double best_eff = 0.0;
    for (size_t n = lo; n <= hi; n++) {
        const size_t ctas = nf * n, waves = (ctas + SMS - 1) / SMS;
        const double eff = (double)ctas / (double)(waves * SMS);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = n;
        }
    }
    return (int)std::max<size_t>(best, 1);
}

}  // namespace

size_t multipole_work_doubles(size_t nf, int lmax, int *nsplit_out, size_t NA) {
    const int ns = mp_nsplit(nf, NA);
    if (nsplit_out) *nsplit_out = ns;
    return (size_t)ns * nf * mp_npairs(lmax) * 2;
}

int launch_multipole_sphere(const float *d_sph, const double *d_b, double ql, const int *d_lm, size_t NM, int lmax,
                            double2 *d_A, size_t ldA, size_t NA, size_t f0, size_t nf, double *d_work,
                            cudaStream_t st) {
    if (nf == 0 || NM == 0) return 0;
    if (lmax > MP_LMAX) return -1;
    const double *tab = mp_tables(lmax, st);
    if (!tab) return -1;
    const int nsplit = mp_nsplit(nf, NA);
    const size_t per = (NA + nsplit - 1) / nsplit;
    const size_t smem = ((size_t)(lmax + 1) * MP_THREADS + (size_t)MP_WARPS * mp_npairs(lmax) * 2) * sizeof(double);
    cudaFuncSetAttribute(multipole_sphere_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int launches = 0;
    double2 *part = reinterpret_cast<double2 *>(d_work);
    // grid.x = frames (<= 2^31-1), grid.y = splits
    multipole_sphere_kernel<<<dim3((unsigned)nf, (unsigned)nsplit), MP_THREADS, smem, st>>>(
        d_sph, d_b, ql, lmax, mp_lstart(lmax), NA, f0, nsplit, per, tab, part, nf);
    launches++;
    for (size_t m0 = 0; m0 < NM; m0 += 65535) {
        const size_t cnt = NM - m0 < 65535 ? NM - m0 : 65535;
        multipole_assemble_kernel<<<dim3((unsigned)((nf + 127) / 128), (unsigned)cnt), 128, 0, st>>>(
            part, nsplit, nf, lmax, d_lm + 2 * m0, cnt, d_A + m0 * ldA, ldA, f0);
        launches++;
    }
    return launches;
}


int multipole_batch_max() { return 8; }

size_t multipole_batch_work_doubles(size_t nf, int lmax, size_t natoms, int NQ) {
    return (size_t)mp_nsplit(nf, natoms, 148 * MG_CTAS_PER_SM) * MG_GROUPS * NQ * nf * mp_npairs(lmax) * 2;
}

// amplitudes of NQ |q| values at once: d_A[q] has NM timelines of ldA entries, q-th block at d_A + q*NM*ldA.
// Atoms [a_first, a_last) only (atom sharding over GPUs); d_b: [NQ][b_stride] factors (b_stride = 0: shared).
int launch_multipole_sphere_batch(const float *d_sph, const double *d_b, size_t b_stride, const double *d_qlens, int NQ,
                                  const int *d_lm, size_t NM, int lmax, double2 *d_A, size_t ldA, size_t NA,
                                  size_t a_first, size_t a_last, size_t f0, size_t nf, double *d_work, cudaStream_t st) {
    if (nf == 0 || NM == 0 || NQ == 0) return 0;
    if (lmax > MP_LMAX) return -1;
    const double *tab = mp_tables(lmax, st);
    if (!tab) return -1;
    const size_t natoms = a_last > a_first ? a_last - a_first : 0;
    const int nsplit = mp_nsplit(nf, std::max<size_t>(natoms, 1), 148 * MG_CTAS_PER_SM);
    const size_t per = (natoms + nsplit - 1) / nsplit;
    const int L1 = lmax + 1, NP = mp_npairs(lmax);
    double2 *part = reinterpret_cast<double2 *>(d_work);
    int launches = 0;
    auto run = [&](auto qtag, int q0) {
        constexpr int Q = decltype(qtag)::value;
        auto smem_of = [&](int A) {
            return ((size_t)2 * A * NP + (size_t)A * (Q * L1 + 2) + 2 * A * 5 + 2 * L1 + 2 * NP) * sizeof(double) + NP * sizeof(int);
        };
        // tile size: the largest even A (even keeps the double2 tables behind sB 16-byte aligned) whose tables fit the shared
        // memory of a CTA and whose (Q + L1/2) * A table tasks (ladders, column pairs) take at most MG_ROUNDS rounds of the MG_THREADS threads.  Measured
        // on config 4 (Q = 8, two CTAs per SM): A = 8 / 12 / 16 / 20 -> 29.4 / 22.0 / 20.2 / 19.6 ms per pass of 16 frames: the
        // two barriers per tile weigh more than a partly filled last round of table tasks
        int A = 2;
        for (int c = 2; c <= MG_A_MAX; c += 2) {
            if (smem_of(c) > MG_SMEM_CAP) break;
            if (((Q + (L1 + 1) / 2) * c + MG_THREADS - 1) / MG_THREADS > MG_ROUNDS) break;  // table tasks of a tile
            A = c;
        }
        const size_t per_tiles = ((per + A - 1) / A) * A;  // splits start on tile boundaries
        cudaFuncSetAttribute(multipole_gemm_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MG_SMEM_CAP);
        multipole_gemm_kernel<Q><<<dim3((unsigned)nf, (unsigned)nsplit), MG_THREADS, smem_of(A), st>>>(
            d_sph, d_b, b_stride, d_qlens, lmax, mp_lstart(lmax), NA, f0, a_first, a_last, per_tiles, tab, part, nf, q0, NQ, A);
        launches++;
    };
    for (int q0 = 0; q0 < NQ;) {
        const int left = NQ - q0;
        if (left >= 8) { run(std::integral_constant<int, 8>(), q0); q0 += 8; }
        else if (left >= 4) { run(std::integral_constant<int, 4>(), q0); q0 += 4; }
        else if (left >= 2) { run(std::integral_constant<int, 2>(), q0); q0 += 2; }
        else { run(std::integral_constant<int, 1>(), q0); q0 += 1; }
    }
    for (int q = 0; q < NQ; q++)
        for (size_t m0 = 0; m0 < NM; m0 += 65535) {
            const size_t cnt = NM - m0 < 65535 ? NM - m0 : 65535;
            multipole_assemble_batch_kernel<<<dim3((unsigned)((nf + 127) / 128), (unsigned)cnt), 128, 0, st>>>(
                part, nsplit * MG_GROUPS, nf, lmax, d_lm + 2 * m0, cnt, d_A + ((size_t)q * NM + m0) * ldA, ldA, f0, q, NQ);
            launches++;
        }
    return launches;
}

}  // namespace sass
