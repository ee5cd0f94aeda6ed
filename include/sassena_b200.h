/*
 * sassena_b200.h — C-ABI of the B200-native Sassena scattering hot path (libsassena_b200.so).
 *
 * This is the drop-in boundary: everything a Sassena scatter device does between "coordinates are
 * staged" and "atfinal_/afinal_/a2final_ are ready for write()" (reference:
 * src/scatter_devices/abstract_scatter_device.cpp:105-175,239-244).  Plain pointers and sizes only;
 * no exceptions cross the boundary — every call returns 0 on success and a non-zero SGPU_E* code on
 * failure, with the message available from sgpu_last_error().  There is NO CPU fallback: without a
 * CUDA device (sm_100a) sgpu_init fails.
 *
 * Conventions
 *   coor_t = float (reference include/common.hpp:37-40), all signal math in FP64.
 *   complex values are interleaved (re, im) doubles, as fftw_complex.
 *   One sgpu_ctx drives one GPU; one process (rank) per GPU.
 */
#ifndef SASSENA_B200_H
#define SASSENA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgpu_ctx sgpu_ctx;

/* error codes */
#define SGPU_OK 0
#define SGPU_EINVAL 1   /* bad argument (the reference's Err::write + throw paths) */
#define SGPU_ECUDA 2    /* CUDA runtime failure */
#define SGPU_ESTATE 3   /* call order violated (e.g. compute before stage) */
#define SGPU_ENOMEM 4   /* device or pinned allocation failed (reference: ram_check / terminate_request) */

/* scattering.dsp.type / scattering.dsp.method (reference: src/control/parameters.cpp:502-511,
 * dispatch in all_vectors_scatter_device.cpp:209-229) */
#define SGPU_DSP_AUTOCORRELATE 0
#define SGPU_DSP_SQUARE 1
#define SGPU_DSP_PLAIN 2
#define SGPU_METHOD_FFTW 0   /* smath::auto_correlate_fftw, src/math/smath.cpp:141-156 */
#define SGPU_METHOD_DIRECT 1 /* smath::auto_correlate_direct, smath.cpp:51-76 (= conj of the fftw form) */

/* coordinate representation of staged frames (reference: CoordinateSets::set_representation,
 * multipole_scatter_device.cpp:53; conversion coordinate_set.cpp:303-315) */
#define SGPU_REPR_CARTESIAN 0
#define SGPU_REPR_SPHERICAL 1 /* (r, phi, theta) per atom */
#define SGPU_REPR_CYLINDRICAL 2 /* (r, phi, z) per atom in the basis built on scattering.average.orientation.axis */

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Create a context on CUDA device `device` (-1: the device the calling thread is bound to).
 * Fails (SGPU_ECUDA) if no sm_100 device is present. */
int sgpu_init(int device, sgpu_ctx **out);
void sgpu_destroy(sgpu_ctx *ctx);
/* Message of the last failing call on ctx (ctx may be NULL: message of the last failed sgpu_init). */
const char *sgpu_last_error(const sgpu_ctx *ctx);
/* Library version string, and the count of kernels this context has launched so far. */
const char *sgpu_version(void);
uint64_t sgpu_launch_count(const sgpu_ctx *ctx);
/* Block until all work queued by ctx is done. */
int sgpu_synchronize(sgpu_ctx *ctx);

/* Pinned host memory for the stager (cudaHostAlloc / cudaFreeHost). */
int sgpu_host_alloc(void **ptr, size_t bytes);
int sgpu_host_free(void *ptr);

/* ---- staging: replaces DataStagerByFrame / DataStagerByAtom ------------------------------------ */

/* DataStagerByFrame::stage (src/stager/data_stager.cpp:72-129): xyz is host float [NF][NA][3]
 * (frame-major, the layout all_vectors_scatter_device.cpp:420 / multipole...:473 index).
 * The copy is asynchronous and chunked by frames; compute calls wait per chunk, so staging overlaps
 * the first compute.  xyz must stay valid until sgpu_synchronize() or the first compute returns.
 * Pinned memory (sgpu_host_alloc) gives true overlap; pageable memory works but copies serially. */
int sgpu_stage_frames(sgpu_ctx *ctx, const float *xyz, size_t NF, size_t NA, int repr);
/* Same, but adopts coordinates already resident in device memory (no copy; caller keeps ownership). */
int sgpu_stage_frames_device(sgpu_ctx *ctx, const float *d_xyz, size_t NF, size_t NA, int repr);
/* Device-side cartesian -> spherical conversion of the staged frames (coor3d.cpp:168-215 +
 * float narrowing data_stager.cpp:111-113); used when an MPSphere device is fed cartesian input. */
int sgpu_frames_to_spherical(sgpu_ctx *ctx);

/* DataStagerByAtom::stage (data_stager.cpp:214-349): xyz is host float [NA_local][NF][3]
 * (atom-major, self_vectors_scatter_device.cpp:303).  Asynchronous and chunked by atoms like sgpu_stage_frames: the
 * autocorrelation path of sgpu_compute_self_vectors* evaluates every chunk as it lands (every other consumer waits for the
 * last one on the device).  xyz must stay valid until sgpu_synchronize() or a compute that returns results has returned. */
int sgpu_stage_atoms(sgpu_ctx *ctx, const float *xyz, size_t NA_local, size_t NF);
int sgpu_stage_atoms_device(sgpu_ctx *ctx, const float *d_xyz, size_t NA_local, size_t NF);
/* Frame-major host input [NF][NA][3] -> atom-major device layout for the atoms this rank owns under
 * ModAssignment(nranks, rank, NA) (assignment.cpp:82-118); the transpose runs on the GPU. */
int sgpu_stage_atoms_from_frames(sgpu_ctx *ctx, const float *xyz, size_t NF, size_t NA, size_t nranks,
                                 size_t rank);

/* The same for a WAVE of atoms: stages atoms atom_first + i*atom_stride, i < count (a slice of a rank's ModAssignment
 * list), for trajectories whose per-rank share exceeds the coordinate budget (BASELINE config 5: the stager streams
 * the atoms through the GPU wave by wave; partials are additive over atoms, see sgpu_accumulate). */
int sgpu_stage_atoms_wave(sgpu_ctx *ctx, const float *xyz, size_t NF, size_t NA, size_t atom_first, size_t atom_stride,
                          size_t count);
/* Double-buffered wave streaming (DataStagerByAtom's cyclic buffering, data_stager.cpp:249-338, turned into a copy/compute
 * pipeline): xyz is host float [count][NF][3] — a block of this rank's atoms in atom-major order, in pinned memory
 * (sgpu_host_alloc) for a truly asynchronous copy.  sgpu_stage_atoms_prefetch queues the H2D copy of the block into the
 * context's BACK wave buffer on the copy stream and returns at once; it waits (on the device) only for the compute calls
 * that still read that buffer.  sgpu_stage_atoms_swap makes the prefetched block the staged atoms: compute calls queued
 * after it wait (on the device) for the copy and then see NA_local = count atoms of NF frames.  The usual loop:
 *     prefetch(wave 0); for w: swap(); if (w+1 < waves) prefetch(wave w+1); for q: compute_self_vectors_partial(...)
 * so that wave w+1 travels over PCIe while wave w is evaluated.  Both buffers belong to the context (2 x the largest block
 * seen); xyz of a prefetched block must stay valid until the matching swap's first compute call has returned. */
int sgpu_stage_atoms_prefetch(sgpu_ctx *ctx, const float *xyz, size_t count, size_t NF);
int sgpu_stage_atoms_swap(sgpu_ctx *ctx);
/* What is staged: frames NF and atoms NA of the staged block, NF_total = frames of the timelines the DSP works on (the
 * frame window's total if one is set, else NF).  Output buffers of compute / finalize hold NF_total complex entries. */
int sgpu_staged_shape(const sgpu_ctx *ctx, size_t *NF, size_t *NA, size_t *NF_total);
/* Bytes of device memory currently held by the context's buffers (coordinates, wave buffers, amplitudes, work areas):
 * the per-GPU HBM high-water mark of a streamed run. */
int sgpu_device_bytes(sgpu_ctx *ctx, size_t *bytes);

/* d_dst[i] += d_src[i], i < n doubles in device memory, on the context's compute stream (sums the packed partials of
 * successive atom waves; fixed order, no atomics). */
int sgpu_accumulate(sgpu_ctx *ctx, double *d_dst, const double *d_src, size_t n);

/* ScatterFactors::get_all() for the current |q| (src/scatter_devices/scatter_factors.cpp:56-78,100):
 * b has one entry per staged atom (NA for frames, NA_local for atoms, in staged order). */
int sgpu_set_factors(sgpu_ctx *ctx, const double *b, size_t n);

/* ---- compute(): one |q| ------------------------------------------------------------------------ */
/* All three fill atfinal[NF][2], afinal[2], a2final[2] exactly as the reference leaves atfinal_,
 * afinal_, a2final_ on partition rank 0 at the end of compute(): summed over the NM subvectors /
 * moments (and atoms for self) and scaled by 1/NM (vectors) or 1/(4 pi) (multipole sphere). */

/* AllVectorsScatterDevice::compute (all_vectors_scatter_device.cpp:238-361).  qvecs[NM][3] are the
 * subvectors produced by init_subvectors() (abstract_vectors_scatter_device.cpp:112-175). */
int sgpu_compute_all_vectors(sgpu_ctx *ctx, const double *qvecs, size_t NM, int dsp_type, int dsp_method,
                             double *atfinal, double afinal[2], double a2final[2]);

/* SelfVectorsScatterDevice::compute (self_vectors_scatter_device.cpp:145-239). */
int sgpu_compute_self_vectors(sgpu_ctx *ctx, const double *qvecs, size_t NM, int dsp_type, int dsp_method,
                              double *atfinal, double afinal[2], double a2final[2]);

/* MPSphereScatterDevice::compute (multipole_scatter_device.cpp:278-401), moments lm[NM][2]=(l,m)
 * (parameters.cpp:1037-1075); requires frames staged with SGPU_REPR_SPHERICAL.
 * Returns SGPU_EINVAL if |m|>l for any moment (multipole_scatter_device.cpp:459-465). */
int sgpu_compute_mpsphere(sgpu_ctx *ctx, double qlen, const long *lm, size_t NM, int dsp_type, int dsp_method,
                          double *atfinal, double afinal[2], double a2final[2]);

/* MPCylinderScatterDevice (multipole_scatter_device.cpp:504-985).  sgpu_frames_to_cylindrical converts the staged
 * cartesian frames in place to (r, phi, z) in the basis built on `axis` (CylindricalCoordinateSet,
 * src/sample/coordinate_set.cpp:278-296; basis src/math/coor3d.cpp:278-304), narrowed to float like the stager does.
 * sgpu_compute_mpcylinder: q[3] is the q-vector (projected onto the same basis, :925-929), moments lm[NM][2] = (l, m)
 * with l >= 0, 0 <= m <= 3 and m = 0 for l = 0 (parameters.cpp:1082-1102); result scaled by 1/(2 pi) (:866).
 * sgpu_mpcylinder_amplitudes writes the amplitudes of the atom range to d_amp ([NM][NF][2] doubles, device) for
 * atom-sharded runs; sgpu_mpsphere_dsp_partial(d_amp, 1, NM, ...) turns the summed amplitudes into a packed partial. */
int sgpu_frames_to_cylindrical(sgpu_ctx *ctx, const double axis[3]);
int sgpu_compute_mpcylinder(sgpu_ctx *ctx, const double q[3], const double axis[3], const long *lm, size_t NM, int dsp_type,
                            int dsp_method, double *atfinal, double afinal[2], double a2final[2]);
int sgpu_compute_mpcylinder_partial(sgpu_ctx *ctx, const double q[3], const double axis[3], const long *lm, size_t NM,
                                    int dsp_type, double *d_partial);
int sgpu_mpcylinder_amplitudes(sgpu_ctx *ctx, const double q[3], const double axis[3], const long *lm, size_t NM,
                               size_t atom_first, size_t atom_count, double *d_amp);

/* Batched multipole sphere: NQ |q| values in one pass.  The q-independent Y_lm tables are built once per atom tile and
 * shared by the batch (the reference recomputes them per moment, atom, frame AND |q|).  Outputs are per |q|:
 * atfinal[NQ][NF][2], afinal[NQ][2], a2final[NQ][2].  Per-|q| scattering factors can be supplied with
 * sgpu_set_factors_batch(b[NQ][NA]); otherwise the set given to sgpu_set_factors applies to every |q|. */
int sgpu_set_factors_batch(sgpu_ctx *ctx, const double *b, size_t NQ, size_t n);
int sgpu_compute_mpsphere_batch(sgpu_ctx *ctx, const double *qlens, size_t NQ, const long *lm, size_t NM, int dsp_type,
                                int dsp_method, double *atfinal, double *afinal, double *a2final);
int sgpu_compute_mpsphere_batch_partial(sgpu_ctx *ctx, const double *qlens, size_t NQ, const long *lm, size_t NM,
                                        int dsp_type, double *d_partials);
/* Multi-GPU multipole: A_lm is a sum over atoms, so ranks shard ATOMS: each rank writes the amplitudes of its atom
 * range for NQ |q| values to d_amp ([NQ][NM][NF][2] doubles, device), the caller all-reduces d_amp, and
 * sgpu_mpsphere_dsp_partial turns the summed amplitudes into NQ packed partials (no second reduction needed). */
int sgpu_mpsphere_amplitudes(sgpu_ctx *ctx, const double *qlens, size_t NQ, const long *lm, size_t NM, size_t atom_first,
                             size_t atom_count, double *d_amp);
int sgpu_mpsphere_dsp_partial(sgpu_ctx *ctx, const double *d_amp, size_t NQ, size_t NM, int dsp_type, double *d_partials);

/* ---- multi-GPU: partial sums + finalize --------------------------------------------------------- */
/* The reference reduces atfinal_/afinal_/a2final_ over the partition with three boost::mpi::reduce
 * calls (all_vectors_scatter_device.cpp:335-343, self...:213-221, multipole...:375-383).  Here every
 * rank computes an UNSCALED packed partial into device memory, the caller sums the packed buffers
 * over ranks (one NCCL all-reduce, f64 sum), and sgpu_finalize turns the sum into the outputs.
 * sgpu_partial_len gives the packed length in doubles for the staged NF and the dsp type.
 * A *_partial call with zero subvectors / moments just zeroes the buffer (a rank the decomposition left idle). */
int sgpu_partial_len(sgpu_ctx *ctx, int dsp_type, size_t *n_doubles);
int sgpu_compute_all_vectors_partial(sgpu_ctx *ctx, const double *qvecs, size_t NM_local, int dsp_type,
                                     double *d_partial);
int sgpu_compute_self_vectors_partial(sgpu_ctx *ctx, const double *qvecs, size_t NM, int dsp_type,
                                      double *d_partial);
int sgpu_compute_mpsphere_partial(sgpu_ctx *ctx, double qlen, const long *lm, size_t NM_local, int dsp_type,
                                  double *d_partial);
/* ---- frame-sharded coherent path (multi-GPU) ----------------------------------------------------------------------
 * The reference's AllVectorsScatterDevice gives every rank of a partition a block of FRAMES
 * (DivAssignment(NNPP, rank, NF), src/scatter_devices/all_vectors_scatter_device.cpp:61,248,408), computes the amplitudes
 * of those frames (scatterblock()/scatter(), :363-430) and exchanges them with boost::mpi::all_to_all so that whole
 * timelines can be correlated (exchange(), :170-205; dsp, :293-322).  On the GPU: each rank stages its frame block, declares where the block
 * sits in the timeline, fills its columns of A[NM][NF_total] (all other columns zero), the ranks sum the buffers
 * (one NCCL all-reduce over NVSwitch), and each rank correlates a block of the NM timelines into a packed partial
 * (summed and finalized like the other *_partial results). */
/* after sgpu_stage_frames*(NF_local frames): they are frames [f_first, f_first+NF_local) of NF_total.  sgpu_partial_len,
 * sgpu_finalize and the two calls below then work on NF_total-frame timelines. */
int sgpu_set_frame_window(sgpu_ctx *ctx, size_t NF_total, size_t f_first);
/* d_amp: device, complex [NM][NF_total]; this rank's columns are written, the others zeroed */
int sgpu_all_vectors_amplitudes(sgpu_ctx *ctx, const double *qvecs, size_t NM, double *d_amp);
/* DSP of timelines [m_first, m_first+m_count) of d_amp into a packed partial (m_count == 0 zeroes it) */
int sgpu_all_vectors_dsp_partial(sgpu_ctx *ctx, const double *d_amp, size_t m_first, size_t m_count, int dsp_type,
                                 double *d_partial);
/* ---- |q|-scan coherent path ---------------------------------------------------------------------------------------
 * A scan (scattering.vectors.scans, parameters.cpp:1125-1189) evaluates the same orientation vectors at many |q|:
 * q_{n,m} = s_n v_m, n < NQ (init_subvectors scales unit vectors by |q| for sphere/file vectors and the cylinder
 * construction is linear in |q| for a fixed direction, abstract_vectors_scatter_device.cpp:96-175).  For equally spaced
 * s_n the phases of one (atom, v_m) pair form an arithmetic progression, and a pass of up to 28 |q| costs two sincos
 * evaluations plus a 3-term recurrence instead of one sincos per |q|.  The reference builds scans from FLOAT-rounded
 * fractions (parameters.cpp:1151), so real scans deviate from a progression by ~1e-8 relative: those take a corrected
 * kernel (first-order term in FP64, second-order term in FP32, exact to ~1e-13).  Any other spacing, and |q|-dependent
 * factors whose rows differ, are evaluated one |q| at a time by the general kernel — the call is valid for any s.
 * Results equal NQ sgpu_compute_all_vectors calls with q = s_n v (to ~1e-13 relative).
 * Factors: sgpu_set_factors (same b for every |q|) or sgpu_set_factors_batch(b[NQ][NA]).
 * v: host [NM][3] direction vectors; s: host [NQ].  Outputs are NQ consecutive blocks laid out like the single-|q|
 * calls: atfinal [NQ][NF][2], afinal/a2final [NQ][2], d_partials NQ packed partials, d_amp [NQ][NM][NF_total] complex. */
int sgpu_compute_all_vectors_scan(sgpu_ctx *ctx, const double *v, size_t NM, const double *s, size_t NQ, int dsp_type,
                                  int dsp_method, double *atfinal, double *afinal, double *a2final);
int sgpu_compute_all_vectors_scan_partial(sgpu_ctx *ctx, const double *v, size_t NM_local, const double *s, size_t NQ,
                                          int dsp_type, double *d_partials);
/* frame-window aware (see sgpu_set_frame_window); follow with sgpu_all_vectors_dsp_partial per |q| plane */
int sgpu_all_vectors_scan_amplitudes(sgpu_ctx *ctx, const double *v, size_t NM, const double *s, size_t NQ, double *d_amp);
/* how the last scan call was evaluated: passes of the plain / corrected scan kernel, |q| values sent to the general kernel */
int sgpu_last_scan_plan(const sgpu_ctx *ctx, int *plain, int *corrected, int *single);
/* scale = 1/NM_total (vectors) or 1/(4 pi) (multipole sphere). */
int sgpu_finalize(sgpu_ctx *ctx, const double *d_partial, int dsp_type, int dsp_method, double scale,
                  double *atfinal, double afinal[2], double a2final[2]);
/* The CUDA stream (cudaStream_t) the partial/finalize work is queued on, for callers that order their
 * own collective against it. */
void *sgpu_stream(sgpu_ctx *ctx);

/* ---- introspection used by tests and bench ------------------------------------------------------ */
/* Raw amplitudes A[NM][NF][2] of the last all_vectors / mpsphere compute (pre-DSP), copied to host. */
int sgpu_get_amplitudes(sgpu_ctx *ctx, double *A, size_t NM, size_t NF);
/* Time of the amplitude kernel(s) of the last compute call, in ms (CUDA events on the compute stream). */
int sgpu_last_amplitude_ms(sgpu_ctx *ctx, float *ms);
/* Time of the DSP (correlation/reduction) kernels of the last compute call, in ms. */
int sgpu_last_dsp_ms(sgpu_ctx *ctx, float *ms);
/* CUDA-event stopwatch on the compute stream: start records an event, stop records another, waits for it and
 * returns the elapsed device time in ms (used by bench.py to time whole steps on the device). */
int sgpu_timer_start(sgpu_ctx *ctx);
int sgpu_timer_stop(sgpu_ctx *ctx, float *ms);
/* Dependency-free DFMA microbenchmark over all SMs: measured FP64 peak in TFLOP/s (FMA = 2 flop). */
int sgpu_measure_fp64_peak(sgpu_ctx *ctx, double *tflops);
/* Fill device coordinates with the synthetic random-walk trajectory of tests/bench (counter-based
 * RNG, see sassena_b200/synth.py for the CPU twin).  layout: 0 = [NF][NA][3], 1 = [NA][NF][3]. */
int sgpu_synth_trajectory(sgpu_ctx *ctx, float *d_xyz, size_t NF, size_t NA, size_t atom0, size_t atom_stride,
                          size_t NA_out, float box, float offset, float step_scale, uint64_t seed, int layout);
/* Plain device allocation helpers for callers without their own allocator (the C++ host layer). */
int sgpu_device_alloc(void **d_ptr, size_t bytes);
int sgpu_device_free(void *d_ptr);
int sgpu_memcpy_d2h(sgpu_ctx *ctx, void *dst, const void *d_src, size_t bytes);
int sgpu_memcpy_h2d(sgpu_ctx *ctx, void *d_dst, const void *src, size_t bytes);

/* ---- the partition's communicator inside the library (NCCL over NVLink / NVSwitch) -----------------------------------
 * Replaces the boost::mpi communicator of one partition (scatter_device_factory.cpp:117-120) for device-resident data.
 * One process per GPU: one rank obtains a unique id (sgpu_comm_get_unique_id, 128 bytes), hands it to the others by any
 * means (torch.distributed, a file, MPI), and every rank calls sgpu_comm_init — a collective (ncclCommInitRank).  All
 * communication then runs on the context's streams with no host synchronisation.  libnccl.so.2 is bound at run time. */
int sgpu_comm_get_unique_id(char *id128);
int sgpu_comm_init(sgpu_ctx *ctx, const char *id128, int nranks, int rank);
/* adopt a communicator the caller owns (an ncclComm_t, e.g. the host layer's partition communicator) */
int sgpu_comm_adopt(sgpu_ctx *ctx, void *nccl_comm, int nranks, int rank);
int sgpu_comm_destroy(sgpu_ctx *ctx);
int sgpu_comm_info(sgpu_ctx *ctx, int *nranks, int *rank);
/* in-place sum of n doubles in device memory over the ranks, queued on the compute stream: the three boost::mpi::reduce
 * calls of all_vectors_scatter_device.cpp:335-343 / self_vectors_scatter_device.cpp:213-221 as one all-reduce of the packed
 * partial, and the amplitude sum of the atom-sharded multipole path. */
int sgpu_comm_allreduce(sgpu_ctx *ctx, double *d_buf, size_t n);
/* AllVectorsScatterDevice::compute with the FRAMES sharded over the partition (NNPP > 1: all_vectors_scatter_device.cpp:
 * 61,248,408).  Every rank has staged its DivAssignment block of the frames (sgpu_set_frame_window).  The ranks evaluate
 * their blocks, exchange them so that every rank holds complete timelines of its DivAssignment block of the subvectors
 * (grouped ncclSend/ncclRecv = the reference's all_to_all, :169-184; assembling the pieces = its alignpad, :186-207; each
 * rank sends 1/N of its amplitudes to each peer instead of summing a zero-padded buffer), correlate them and sum the packed
 * partials.  The exchange of a pass overlaps the next pass's amplitudes.  On return (asynchronously) d_partials holds the
 * REDUCED partials on every rank: [NQ][sgpu_partial_len] for the scan form, one partial for the single-|q| form; pass them
 * to sgpu_finalize.  Collective: every rank of the communicator must call with the same arguments. */
int sgpu_compute_all_vectors_scan_sharded(sgpu_ctx *ctx, const double *v, size_t NM, const double *s, size_t NQ, int dsp_type,
                                          double *d_partials);
int sgpu_compute_all_vectors_sharded(sgpu_ctx *ctx, const double *qvecs, size_t NM, int dsp_type, double *d_partial);

#ifdef __cplusplus
}
#endif
#endif /* SASSENA_B200_H */
