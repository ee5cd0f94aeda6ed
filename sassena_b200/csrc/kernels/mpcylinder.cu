// mpcylinder.cu — multipole cylinder amplitudes (K6), the GPU side of MPCylinderScatterDevice::scatter
// (reference src/scatter_devices/multipole_scatter_device.cpp:905-985).
//
// The reference evaluates, per (moment, frame),
//     A = sqrt(2 pi) * sum_j b_j * exp(i |z_j q_z|) * f_{l,m}(r_j q_r, phi_j - phi_q)
//     f_{0,0} = J_0(p);   f_{l,0} = sqrt(1/2) 2 (-1)^l J_{2l}(p) cos(2l psi);        f_{l,1} = ... sin(2l psi)
//                         f_{l,2} = sqrt(1/2) 2i (-1)^{l-1} J_{2l-1}(p) cos((2l-1) psi);  f_{l,3} = ... sin((2l-1) psi)
// with one Boost cyl_bessel_j call per (moment, atom, frame).  (The phase is exp(i*sign(z q_z)*z*q_z) = exp(i|z q_z|) in the
// reference, :941-947; reproduced as is.)  All moments of a frame share the per-atom quantities, so the kernel computes
// for every Bessel order n <= nmax the four real sums
//     S_n = sum_j w_j J_n(p_j) (cos n psi_j, sin n psi_j),   w_j = b_j exp(i |z_j q_z|)   (complex)
// and an epilogue maps orders to moments.  Per tile of 128 atoms: 128 threads run the Bessel ladders (upward recurrence
// from j0/j1 for n <= p, continued-fraction ratios J_n/J_{n-1} from a downward sweep for n > p: bounded by 1, no
// overflow, no renormalisation), the other 128 run the rotation ladders cos/sin(n psi) and the phase factor; then every
// thread owns one (order, atom slice) pair and accumulates its four sums from shared memory (lanes = consecutive orders:
// conflict-free with the odd row pitch).  FP64 throughout; coordinates are the staged floats (r, phi, z).
#include "kernels.hpp"

#include <algorithm>

namespace sass {

namespace {

constexpr int CY_THREADS = 256;
constexpr int CY_TILE = 128;

// CylinderCoor3D(base.project(c)) (reference src/math/coor3d.cpp:113-138, 296-298) + float narrowing
// (src/stager/data_stager.cpp:111-113).  sign(M_PI, y) in the reference returns FLOAT (include/math/coor3d.hpp:30), so
// the quadrant offsets are float-rounded pi and pi/2; 2*M_PI in the wrap-around is double.
__global__ void cart_to_cylindrical_kernel(float *xyz, size_t n, double e00, double e01, double e02, double e10, double e11,
                                           double e12, double e20, double e21, double e22) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double PI = 3.14159265358979323846;
    const double PIf = (double)3.14159274101257324f, PI2f = (double)1.57079637050628662f;
    const double cx = xyz[3 * i], cy = xyz[3 * i + 1], cz = xyz[3 * i + 2];
    // vec * base_[k] = x*bx + y*by + z*bz, left to right (CartesianCoor3D::operator*)
    const double x = __dadd_rn(__dadd_rn(__dmul_rn(cx, e00), __dmul_rn(cy, e01)), __dmul_rn(cz, e02));
    const double y = __dadd_rn(__dadd_rn(__dmul_rn(cx, e10), __dmul_rn(cy, e11)), __dmul_rn(cz, e12));
    const double z = __dadd_rn(__dadd_rn(__dmul_rn(cx, e20), __dmul_rn(cy, e21)), __dmul_rn(cz, e22));
    const double r = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
    double phi = 0.0;
    if (x != 0.0) {
        phi = atan(y / x);
        if (x < 0.0) phi = ((y < 0.0) ? -PIf : PIf) + phi;
    } else if (y != 0.0) {
        phi = (y < 0.0) ? -PI2f : PI2f;
    }
    if (phi < 0) phi = 2 * PI + phi;
    xyz[3 * i] = (float)r;
    xyz[3 * i + 1] = (float)phi;
    xyz[3 * i + 2] = (float)z;
}

// grid = (frames, splits).  partial[(split*NF + f)*nord + n] = (sum w_r Jc, sum w_i Jc, sum w_r Js, sum w_i Js)
__global__ void __launch_bounds__(CY_THREADS) mpcylinder_kernel(const float *__restrict__ coords,
                                                                const double *__restrict__ b, size_t NA, size_t a_first,
                                                                size_t a_last, double qr, double qphi, double qz,
                                                                int nmax, int mstart, size_t f0, double4 *__restrict__ partial,
                                                                size_t NF_out) {
    extern __shared__ __align__(16) double sm[];
    const int nord = nmax + 1;
    const int pitch = nord | 1;  // odd row pitch
    double *Jt = sm;                                                      // [CY_TILE][pitch]
    double2 *Ct = reinterpret_cast<double2 *>(sm + CY_TILE * pitch);      // [CY_TILE][pitch]
    double2 *W = Ct + CY_TILE * pitch;                                    // [CY_TILE]
    double4 *red = reinterpret_cast<double4 *>(W + CY_TILE);              // [slices][nord]
    const size_t f = blockIdx.x;
    const int S = gridDim.y;
    const size_t span = a_last - a_first;
    const size_t per = (span + S - 1) / S;
    const size_t s_first = a_first + blockIdx.y * per;
    const size_t s_last = (s_first + per < a_last) ? s_first + per : a_last;
    const float *fr = coords + (f0 + f) * NA * 3;
    const int tid = threadIdx.x;
    const int nslices = CY_THREADS / nord;  // >= 1 (nord <= 256)
    const int my_n = tid % nord, my_slice = tid / nord;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (size_t t0 = s_first; t0 < s_last; t0 += CY_TILE) {
        const int cnt = (int)((s_last - t0 < (size_t)CY_TILE) ? s_last - t0 : CY_TILE);
        __syncthreads();
        if (tid < CY_TILE) {  // Bessel ladder of atom tid
            const int a = tid;
            if (a < cnt) {
                const double x = (double)fr[3 * (t0 + a)] * qr;
                double *J = Jt + a * pitch;
                int n0 = (x >= (double)nmax) ? nmax : (int)x;  // orders <= x: upward recurrence is stable
                double jm = j0(x);
                J[0] = jm;
                if (n0 >= 1) {
                    double jc = j1(x);
                    J[1] = jc;
                    const double tox = 2.0 / x;
                    for (int n = 1; n < n0; n++) {
                        const double jn = (double)n * tox * jc - jm;
                        jm = jc;
                        jc = jn;
                        J[n + 1] = jn;
                    }
                }
                if (n0 < nmax) {
                    // ratios rho_n = J_n / J_{n-1} = x / (2n - x rho_{n+1}), downward from mstart (rho ~ 0 there)
                    double rho = 0.0;
                    for (int n = mstart; n > nmax; n--) rho = x / (2.0 * n - x * rho);
                    for (int n = nmax; n > n0; n--) {
                        rho = x / (2.0 * n - x * rho);
                        J[n] = rho;
                    }
                    double prev = J[n0];
                    for (int n = n0 + 1; n <= nmax; n++) {
                        prev *= J[n];
                        J[n] = prev;
                    }
                }
            }
        } else {  // rotation ladder and phase factor of atom tid - CY_TILE
            const int a = tid - CY_TILE;
            if (a < cnt) {
                const double psi = (double)fr[3 * (t0 + a) + 1] - qphi;
                const double z = (double)fr[3 * (t0 + a) + 2];
                double s1, c1, se, ce;
                sincos(psi, &s1, &c1);
                sincos(fabs(z * qz), &se, &ce);
                const double bb = b[t0 + a];
                W[a] = make_double2(bb * ce, bb * se);
                double2 *C = Ct + a * pitch;
                double c = 1.0, s = 0.0;
                C[0] = make_double2(1.0, 0.0);
                for (int n = 1; n <= nmax; n++) {
                    const double cn = c * c1 - s * s1;
                    s = s * c1 + c * s1;
                    c = cn;
                    C[n] = make_double2(c, s);
                }
            }
        }
        __syncthreads();
        if (my_slice < nslices) {
            for (int a = my_slice; a < cnt; a += nslices) {
                const double j = Jt[a * pitch + my_n];
                const double2 cs = Ct[a * pitch + my_n];
                const double2 w = W[a];
                const double jc = j * cs.x, js = j * cs.y;
                a0 = fma(w.x, jc, a0);
                a1 = fma(w.y, jc, a1);
                a2 = fma(w.x, js, a2);
                a3 = fma(w.y, js, a3);
            }
        }
    }
    __syncthreads();
    if (my_slice < nslices) red[my_slice * nord + my_n] = make_double4(a0, a1, a2, a3);
    __syncthreads();
    if (tid < nord) {
        double4 t = red[tid];
        for (int s = 1; s < nslices; s++) {
            const double4 v = red[s * nord + tid];
            t.x += v.x;
            t.y += v.y;
            t.z += v.z;
            t.w += v.w;
        }
        partial[((size_t)blockIdx.y * NF_out + f) * nord + tid] = t;
    }
}

// A[mom][f] = sqrt(2 pi) * factor(l, m) * (sum over splits of the order's cos or sin sums)
__global__ void mpcylinder_epilogue_kernel(const double4 *__restrict__ partial, int S, size_t NF, int nord,
                                           const int *__restrict__ lm, size_t NM, double2 *__restrict__ A) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= NM * NF) return;
    const size_t mom = i / NF, f = i % NF;
    const int l = lm[2 * mom], m = lm[2 * mom + 1];
    const int n = (l == 0) ? 0 : ((m < 2) ? 2 * l : 2 * l - 1);
    double re = 0.0, im = 0.0;
    for (int s = 0; s < S; s++) {
        const double4 v = partial[((size_t)s * NF + f) * nord + n];
        if ((m & 1) == 0) {
            re += v.x;
            im += v.y;
        } else {
            re += v.z;
            im += v.w;
        }
    }
    const double norm = sqrt(2.0 * 3.14159265358979323846);
    if (l == 0) {  // (0,0): J_0 alone
        A[i] = make_double2(norm * re, norm * im);
        return;
    }
    const double half = sqrt(0.5);
    if (m < 2) {
        const double fac = 2.0 * ((l & 1) ? -1.0 : 1.0);  // 2 (-1)^l
        A[i] = make_double2(norm * (half * fac * re), norm * (half * fac * im));
    } else {
        const double fac = 2.0 * (((l - 1) & 1) ? -1.0 : 1.0);  // i 2 (-1)^(l-1)
        A[i] = make_double2(norm * (-(half * fac) * im), norm * (half * fac * re));
    }
}

size_t cyl_smem_bytes(int nmax) {
    const int nord = nmax + 1, pitch = nord | 1;
    const int nslices = CY_THREADS / nord;
    return (size_t)CY_TILE * pitch * (sizeof(double) + sizeof(double2)) + CY_TILE * sizeof(double2) +
           (size_t)std::max(nslices, 1) * nord * sizeof(double4);
}

// Atom splits per frame: grid = NF x splits CTAs.  Between one and two waves of the CTAs that are resident at once (occupancy
// query: registers and the nmax-dependent shared memory), the count whose last wave is fullest; never less than one tile per
// split.  (A fixed 2 x 148 target left the SMs at 0.68 of one wave: three CTAs fit per SM at nmax = 20.)
int cyl_splits(size_t NF, size_t natoms, int nmax) {
    static int cached_nmax = -1;
    static size_t resident = 2 * 148;
    if (nmax != cached_nmax) {
        int per_sm = 1, dev = 0, sms = 148;
        const size_t smem = cyl_smem_bytes(nmax);
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(mpcylinder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mpcylinder_kernel, CY_THREADS, smem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        resident = (size_t)per_sm * (size_t)sms;
        cached_nmax = nmax;
    }
    const size_t maxs = std::max<size_t>(1, natoms / CY_TILE);
    const size_t lo = std::min(maxs, std::max<size_t>(1, (resident + NF - 1) / NF));
    const size_t hi = std::min(maxs, std::max(lo, (2 * resident + NF - 1) / NF));
    size_t best = lo;
    double best_eff = 0.0;
    for (size_t n = lo; n <= hi; n++) {
        const size_t ctas = NF * n, waves = (ctas + resident - 1) / resident;
        const double eff = (double)ctas / (double)(waves * resident);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = n;
        }
    }
    return (int)best;
}

}  // namespace

int launch_cart_to_cylindrical(float *d_xyz, size_t n, const double base[9], cudaStream_t st) {
    if (n == 0) return 0;
    cart_to_cylindrical_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyz, n, base[0], base[1], base[2], base[3], base[4],
                                                                           base[5], base[6], base[7], base[8]);
    return 1;
}

int mpcylinder_max_order() { return 199; }

size_t mpcylinder_work_doubles(size_t NF, int nmax, size_t natoms) {
    return (size_t)cyl_splits(NF, natoms, nmax) * NF * (nmax + 1) * 4;
}

int launch_mpcylinder(const float *d_coords, const double *d_b, double qr, double qphi, double qz, const int *d_lm, size_t NM,
                      int nmax, double2 *d_A, size_t NF, size_t NA, size_t a_first, size_t a_last, double *d_work,
                      cudaStream_t st) {
    if (NF == 0 || NM == 0) return 0;
    const int nord = nmax + 1;
    const size_t natoms = a_last - a_first;
    const int S = cyl_splits(NF, std::max<size_t>(natoms, 1), nmax);
    const size_t smem = cyl_smem_bytes(nmax);
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(mpcylinder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    // start of the downward ratio sweep: far enough above nmax for 1e-16 (Miller's rule of thumb, sqrt(160 n))
    const int mstart = nmax + 2 + (int)std::sqrt(160.0 * (nmax + 1));
    double4 *partial = reinterpret_cast<double4 *>(d_work);
    int launches = 0;
    for (size_t f0 = 0; f0 < NF; f0 += 65535) {  // grid.x limit is 2^31-1, but keep launches modest
        const size_t nf = std::min<size_t>(NF - f0, 65535);
        dim3 grid((unsigned)nf, (unsigned)S);
        mpcylinder_kernel<<<grid, CY_THREADS, smem, st>>>(d_coords, d_b, NA, a_first, a_last, qr, qphi, qz, nmax, mstart, f0,
                                                         partial + f0 * nord, NF);
        launches++;
    }
    const size_t total = NM * NF;
    mpcylinder_epilogue_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, S, NF, nord, d_lm, NM, d_A);
    return launches + 1;
}

}  // namespace sass
