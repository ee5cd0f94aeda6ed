// kernels.hpp — launchers of the sm_100a kernels behind the C-ABI (include/sassena_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace sass {

// ---- amplitude.cu -------------------------------------------------------------------------------
// K1: A[m][f] = sum_j b_j exp(i q_m . r_j(f)) for frames [f0, f0+nf), all NM vectors.
// d_qs: [NMpad][3] q-vectors pre-scaled by 2/pi, zero padded to a multiple of amplitude_all_qpad().
// Returns the number of kernel launches queued.
int amplitude_all_qpad();
int launch_amplitude_all(const float *d_xyz, const double *d_b, const double *d_qs, double2 *d_A, size_t ldA,
                         size_t NA, size_t NM, size_t f0, size_t nf, cudaStream_t st);
// K2: per-atom timelines a[(n*NM+m)][t] = b_n exp(i q_m . r_n(t)) for local atoms [n0, n0+nn).
// |q|-scan amplitudes, one pass: q_{n,m} = (s0 + n ds) v_m for n < nq (nq <= amplitude_scan_max_pass()).
// d_vs: [NMpad][3] direction vectors pre-scaled by 2/pi, zero padded to a multiple of amplitude_scan_qpad().
// kappa (host, [nq], may be NULL): (pi/2) * (s_n - (s0 + n ds)) for |q| values that are only approximately equally
// spaced; selects the corrected kernel (exact to third order in kappa_n * sigma).
// b_stride != 0: |q|-dependent factors, d_b points at the row of the pass's first |q|, rows b_stride doubles apart
// (plain kernel only).
// d_A: [nq][NM][ldA] with strideQ entries between |q| planes.  Returns the launch count, -1 if nq is too large.
int amplitude_scan_qpad();
int amplitude_scan_max_pass(int kind);  // 0 plain, 1 corrected, 2 |q|-dependent factors
int launch_amplitude_scan_pass(const float *d_xyz, const double *d_b, const double *d_vs, double s0, double ds, int nq,
                               const double *kappa, double2 *d_A, size_t ldA, size_t strideQ, size_t NA, size_t NM,
                               size_t f0, size_t nf, cudaStream_t st, size_t b_stride = 0);
// (pi/2) * (s_n - (s0 + n ds)): first/second-order phase correction per |q| of a pass of the corrected scan kernel
struct ScanKappa {
    double k[32];
};
// symmetric (Chebyshev) form of the scan kernel, scan_sym.cu: passes of up to amplitude_scan_sym_max_pass(corrected) |q|
// values, uniform factors only; same arguments as launch_amplitude_scan_pass
int amplitude_scan_sym_qpad();
int amplitude_scan_sym_max_pass(int corrected);
int launch_amplitude_scan_sym_pass(const float *d_xyz, const double *d_b, const double *d_vs, double s0, double ds, int nq,
                                   const double *kappa, double2 *d_A, size_t ldA, size_t strideQ, size_t NA, size_t NM,
                                   size_t f0, size_t nf, cudaStream_t st);
// max over the buffer of |x| (n floats); result in *d_out (float, device)
int launch_max_abs(const float *d_x, size_t n, float *d_out, cudaStream_t st);
int launch_amplitude_self(const float *d_xyz_by_atom, const double *d_b, const double *d_qs, double2 *d_A,
                          size_t ldA, size_t NF, size_t NM, size_t n0, size_t nn, cudaStream_t st);
// cart -> (r, phi, theta), in place, n points
int launch_cart_to_spherical(float *d_xyz, size_t n, cudaStream_t st);
// multipole cylinder (mpcylinder.cu): in-place cartesian -> (r, phi, z) in the basis `base` (rows e_r, e_phi, e_z);
// amplitudes A[NM][NF] of the cylinder moments lm for one q = (qr, qphi, qz) over atoms [a_first, a_last)
int launch_cart_to_cylindrical(float *d_xyz, size_t n, const double base[9], cudaStream_t st);
int mpcylinder_max_order();
size_t mpcylinder_work_doubles(size_t NF, int nmax, size_t natoms);
int launch_mpcylinder(const float *d_coords, const double *d_b, double qr, double qphi, double qz, const int *d_lm, size_t NM,
                      int nmax, double2 *d_A, size_t NF, size_t NA, size_t a_first, size_t a_last, double *d_work,
                      cudaStream_t st);
// chunk of frames [nf][NA][3] starting at frame f0 -> [NA_out][NF][3] taking atoms atom0 + i*stride
int launch_frames_to_atoms(const float *d_frames, float *d_atoms, size_t NF, size_t nf, size_t f0, size_t NA,
                           size_t atom0, size_t stride, size_t NA_out, cudaStream_t st);
int launch_synth_trajectory(float *d_xyz, size_t NF, size_t NA, size_t atom0, size_t atom_stride, size_t NA_out,
                            float box, float offset, float step_scale, uint64_t seed, int layout, cudaStream_t st);
// SM-driven copy of 4-byte words from mapped pinned host memory (keeps small uploads off the DMA engine, where they
// would queue behind the stager's bulk copies)
int launch_copy_words(void *d_dst, const void *mapped_src, size_t nwords, cudaStream_t st);
// dst[i] += src[i]
int launch_accumulate(double *d_dst, const double *d_src, size_t n, cudaStream_t st);
// DFMA peak probe; returns flops executed
double launch_fp64_peak(double *d_sink, int iters, int blocks, cudaStream_t st);

// ---- multipole.cu -------------------------------------------------------------------------------
// K5: A[mom][f] = sum_j 4pi i^l b_j j_l(q r_j) conj(Y_lm(theta_j, phi_j)), frames [f0,f0+nf).
// d_lm: [NM][2] (l,m) as int; lmax = max l.  d_work: scratch of multipole_work_doubles() doubles.
size_t multipole_work_doubles(size_t nf, int lmax, int *nsplit_out, size_t NA);
int launch_multipole_sphere(const float *d_sph, const double *d_b, double ql, const int *d_lm, size_t NM, int lmax,
                            double2 *d_A, size_t ldA, size_t NA, size_t f0, size_t nf, double *d_work,
                            cudaStream_t st);

// batched form: NQ |q| values per pass (Y_lm table shared), atoms [a_first, a_last) only; d_A block q at d_A + q*NM*ldA;
// d_b: [NQ][b_stride] (b_stride = 0: one factor set for all q); d_work >= multipole_batch_work_doubles() doubles
int multipole_batch_max();
size_t multipole_batch_work_doubles(size_t nf, int lmax, size_t natoms, int NQ);
int launch_multipole_sphere_batch(const float *d_sph, const double *d_b, size_t b_stride, const double *d_qlens, int NQ,
                                  const int *d_lm, size_t NM, int lmax, double2 *d_A, size_t ldA, size_t NA,
                                  size_t a_first, size_t a_last, size_t f0, size_t nf, double *d_work, cudaStream_t st);

// ---- selffused.cu -------------------------------------------------------------------------------
// fused self-scattering path (amplitudes + FFT autocorrelation in shared memory), dsp = autocorrelate
struct SelfPlan {
    size_t NF = 0;
    int log2N = 0;   // sub-FFT length N = 2^log2N (<= 4096)
    size_t N = 0;
    int R = 0;       // residues; padded length L = R*N >= 2NF-1
    size_t L = 0;
    double2 *d_tw = nullptr;  // exp(-2 pi i k / N), k < N
    double2 *d_w = nullptr;   // weights What[j][pos] (residue-major, digit-reversed position)
    int *d_freq = nullptr;    // frequency index of a digit-reversed position
    // split path (R >= 3): R decimated sub-FFTs per timeline + an R-point DFT across them (selffused.cu, "split path")
    bool split = false;
    bool reg_combine = false;    // R <= 16: the R-point DFT runs in registers (one thread per position)
    bool two_stage = false;      // composite R in 17..32: two-stage DFT in registers
    int S = 0, C = 0;            // positions per combine CTA, number of position slices (C*S = N)
    double2 *d_w2 = nullptr;     // weights in the split layout [k2][pos]
    int *d_perm = nullptr;       // split index k2*N + pos -> residue-major index j*N + pos' of the same frequency
};
int self_plan_create(SelfPlan *p, size_t NF, cudaStream_t st, uint64_t *launches);
void self_plan_destroy(SelfPlan *p);
size_t self_work_bytes(const SelfPlan *p, size_t ntl);
// timelines of local atoms [atom0, atom0+natoms) x NM q-vectors; d_P[L] and d_acc[4] are accumulated (+=)
// dec != 0: the coordinates of every atom are in the split path's decimated order (self_decimate_layout, forward)
int self_power_accumulate(const SelfPlan *p, const float *d_xyz_by_atom, const double *d_b, const double *d_qs,
                          size_t NM, size_t atom0, size_t natoms, void *d_work, double *d_P, double *d_acc, int dec,
                          cudaStream_t st);
// natural frame order [n] <-> decimated [r][m] (frames R m + r contiguous per r) for natoms rows, out of place
int self_decimate_layout(const float *d_src, float *d_dst, size_t natoms, size_t NF, int R, int forward, cudaStream_t st);
int self_finalize(const SelfPlan *p, const double *d_P, void *d_work, double2 *d_out, double scale, int conj_out,
                  cudaStream_t st);
// the same DSP for nt timelines that already exist in memory, A[nt][ldA] (first NF entries valid); needs p->R == 1
size_t self_loaded_work_bytes(const SelfPlan *p, size_t nt);
int self_power_accumulate_loaded(const SelfPlan *p, const double2 *d_A, size_t ldA, size_t nt, void *d_work, double *d_P,
                                 double *d_acc, cudaStream_t st);


// ---- correlate.cu -------------------------------------------------------------------------------
struct CorrPlan {
    size_t NF = 0;   // frames per timeline
    size_t L = 0;    // padded FFT length (power of two >= 2*NF)
    int log2N1 = 0;  // L = N1*N2, column FFT length N1 (strided), row FFT length N2 (contiguous)
    int log2N2 = 0;
    double2 *d_tw = nullptr;  // W_Nmax^k, k < Nmax/2  (Nmax = max(N1,N2)), forward sign
    double2 *d_w = nullptr;   // weights What'[k1*N2+k2] for a_m = sum_k P[k] What[k] / (NF*L)
    size_t Nmax = 0;
    // short timelines (2NF-1 <= 4096): the whole transform fits one SM's shared memory, so the DSP runs in ONE kernel that
    // reads every timeline once and never writes the spectrum (the in-SM FFT of the self path, selffused.cu) instead of
    // the two-pass four-step FFT through HBM.  L, the layout of P and finalize are then the embedded self plan's.
    bool in_sm = false;
    SelfPlan sm;
};
int corr_plan_create(CorrPlan *p, size_t NF, cudaStream_t st, uint64_t *launches);
void corr_plan_destroy(CorrPlan *p);
// bytes of scratch needed for nt timelines
size_t corr_work_bytes(const CorrPlan *p, size_t nt);
// DSP=autocorrelate over nt timelines A[nt][ldA] (first NF entries valid).  Accumulates (+=) into
//   d_P[L] (power spectrum, internal [k1][k2] order), d_acc[4] = {sum a_re, sum a_im, sum |a|^2, 0}.
int corr_power_accumulate(const CorrPlan *p, const double2 *d_A, size_t ldA, size_t nt, void *d_work, double *d_P,
                          double *d_acc, cudaStream_t st);
// inverse transform of the summed power spectrum: d_out[tau] = scale * c[tau]/(L*(NF-tau)), tau<NF;
// conj_out negates the imaginary part (dsp.method=direct).  d_work >= corr_work_bytes(p,1).
int corr_finalize(const CorrPlan *p, const double *d_P, void *d_work, double2 *d_out, double scale, int conj_out,
                  cudaStream_t st);
// DSP=square / plain: d_at[NF] += sum_m f(A[m][t]); d_acc += {sum_m mean_t, sum_m |mean_t|^2}
int dsp_elementwise_accumulate(const double2 *d_A, size_t ldA, size_t nt, size_t NF, int square, double2 *d_at,
                               double *d_acc, void *d_work, cudaStream_t st);
size_t dsp_elementwise_work_bytes(size_t nt);
// out[i] = in[i]*scale (complex, n entries); acc_out[0..3] = acc[0..3]*scale
int launch_scale_complex(const double2 *d_in, double2 *d_out, size_t n, double scale, cudaStream_t st);


}  // namespace sass
