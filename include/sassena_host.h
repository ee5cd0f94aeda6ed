/*
 * sassena_host.h — C entry points of the host layer (sassena_b200/csrc/host): the reference's scatter-device
 * interface (ScatterDeviceFactory::create + IScatterDevice::run, src/scatter_devices/scatter_device_factory.cpp:23-210,
 * src/scatter_devices/abstract_scatter_device.cpp:105-175) driven from C / ctypes, plus the pure host logic
 * (assignments, decomposition plan, q-vector / orientation / moment generators) as unit-testable functions.
 * All functions return 0 on success; on failure the message is in sass_last_error().
 */
#ifndef SASSENA_HOST_H
#define SASSENA_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "sassena_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* communicator callbacks: stands in for boost::mpi::communicator (one rank = one GPU). allreduce_sum sums n doubles
 * in DEVICE memory over the ranks (NCCL all-reduce); the callee must leave the result visible to the caller's
 * subsequent CUDA work (synchronise its stream before returning). */
typedef struct sass_comm_vtbl {
    void *user;
    size_t (*rank)(void *user);
    size_t (*size)(void *user);
    int (*allreduce_sum)(void *user, double *d_buf, size_t n);
    int (*barrier)(void *user);
    void *(*split)(void *user, int color); /* boost::mpi::communicator::split; returns the new `user` handle */
    void (*release)(void *user);
    /* optional (may be NULL): the ncclComm_t behind this communicator.  When present, the frame-sharded coherent device hands
     * it to the library (sgpu_comm_adopt) and exchanges amplitudes with ncclSend/ncclRecv on the device streams. */
    void *(*nccl_comm)(void *user);
} sass_comm_vtbl;

/* A communicator implemented inside the library on NCCL (one process per GPU): every callback is native code, so
 * sass_scatter_run / sass_job_run drive a multi-GPU run without any Python on the path.  One rank creates the id
 * (sass_comm_nccl_unique_id) and distributes it — sass_comm_nccl_bootstrap_file does that through a file on a single
 * node: rank 0 writes `path`, the others wait for it.  sass_comm_nccl_create is collective (ncclCommInitRank); split() is
 * ncclCommSplit; all-reduces run on a stream of the communicator's own and are complete on return. */
int sass_comm_nccl_unique_id(char id128[128]);
int sass_comm_nccl_bootstrap_file(const char *path, int nranks, int rank, double timeout_s, char id128[128]);
int sass_comm_nccl_create(const char id128[128], int nranks, int rank, int device, sass_comm_vtbl *out);

/* The C-ABI (sassena_b200.h) as a table.  NULL selects the library's own sgpu_* functions; tests bind the table
 * to the CPU oracle to exercise the multi-rank host logic without a GPU. */
typedef struct sass_backend_vtbl {
    int (*init)(int, sgpu_ctx **);
    void (*destroy)(sgpu_ctx *);
    const char *(*last_error)(const sgpu_ctx *);
    int (*synchronize)(sgpu_ctx *);
    int (*stage_frames)(sgpu_ctx *, const float *, size_t, size_t, int);
    int (*frames_to_spherical)(sgpu_ctx *);
    int (*stage_atoms)(sgpu_ctx *, const float *, size_t, size_t);
    int (*stage_atoms_from_frames)(sgpu_ctx *, const float *, size_t, size_t, size_t, size_t);
    int (*set_factors)(sgpu_ctx *, const double *, size_t);
    int (*partial_len)(sgpu_ctx *, int, size_t *);
    int (*compute_all_vectors_partial)(sgpu_ctx *, const double *, size_t, int, double *);
    int (*compute_self_vectors_partial)(sgpu_ctx *, const double *, size_t, int, double *);
    int (*compute_mpsphere_partial)(sgpu_ctx *, double, const long *, size_t, int, double *);
    int (*finalize)(sgpu_ctx *, const double *, int, int, double, double *, double *, double *);
    int (*device_alloc)(void **, size_t);
    int (*device_free)(void *);
    /* batched / atom-sharded multipole path */
    int (*set_factors_batch)(sgpu_ctx *, const double *, size_t, size_t);
    int (*mpsphere_amplitudes)(sgpu_ctx *, const double *, size_t, const long *, size_t, size_t, size_t, double *);
    int (*mpsphere_dsp_partial)(sgpu_ctx *, const double *, size_t, size_t, int, double *);
    /* frame-sharded coherent path */
    int (*set_frame_window)(sgpu_ctx *, size_t, size_t);
    int (*all_vectors_amplitudes)(sgpu_ctx *, const double *, size_t, double *);
    int (*all_vectors_dsp_partial)(sgpu_ctx *, const double *, size_t, size_t, int, double *);
    /* |q|-scan coherent path */
    int (*compute_all_vectors_scan_partial)(sgpu_ctx *, const double *, size_t, const double *, size_t, int, double *);
    int (*all_vectors_scan_amplitudes)(sgpu_ctx *, const double *, size_t, const double *, size_t, double *);
    /* atom waves of the self path (trajectory share larger than limits.stage.memory.data) */
    int (*stage_atoms_wave)(sgpu_ctx *, const float *, size_t, size_t, size_t, size_t, size_t);
    int (*accumulate)(sgpu_ctx *, double *, const double *, size_t);
    /* multipole cylinder */
    int (*frames_to_cylindrical)(sgpu_ctx *, const double *);
    int (*mpcylinder_amplitudes)(sgpu_ctx *, const double *, const double *, const long *, size_t, size_t, size_t, double *);
    /* double-buffered atom waves of the self path: pinned host blocks [count][NF][3] -> back buffer, then swap */
    int (*stage_atoms_prefetch)(sgpu_ctx *, const float *, size_t, size_t);
    int (*stage_atoms_swap)(sgpu_ctx *);
    int (*host_alloc)(void **, size_t);
    int (*host_free)(void *);
    /* optional (may be NULL): NCCL inside the library for the frame-sharded coherent path */
    int (*comm_adopt)(sgpu_ctx *, void *, int, int);
    int (*compute_all_vectors_scan_sharded)(sgpu_ctx *, const double *, size_t, const double *, size_t, int, double *);
    int (*compute_all_vectors_sharded)(sgpu_ctx *, const double *, size_t, int, double *);
} sass_backend_vtbl;

const char *sass_last_error(void);

/* ---- Params: keys are the scatter.xml paths (src/control/parameters.cpp:372-397,612-636) -------------------- */
typedef struct sass_params sass_params;
sass_params *sass_params_new(void);
void sass_params_free(sass_params *p);
/* keys: scattering.type, scattering.dsp.type, scattering.dsp.method, scattering.average.orientation.type,
 * scattering.average.orientation.axis.{x,y,z}, scattering.average.orientation.vectors.{type,algorithm,resolution,seed},
 * scattering.average.orientation.multipole.type, scattering.average.orientation.multipole.moments.{type,resolution},
 * limits.stage.memory.data, limits.stage.stream (self: stream atom waves when the share exceeds the budget; default
 * true), limits.decomposition.utilization, limits.decomposition.partitions.{automatic,size},
 * limits.decomposition.coherent (auto | frames | vectors: how a partition's ranks share one coherent |q|),
 * limits.computation.scan (largest |q| batch of the coherent scan path; 0 or 1 = one |q| per pass),
 * limits.computation.scan_snap (true: |q| equally spaced to within 1e-6 are moved onto the exact progression; off by
 * default because it changes the q-vectors with respect to the reference's float-rounded scan fractions) */
int sass_params_set(sass_params *p, const char *key, const char *value);
/* vectors.type=file rows / multipole.moments.type=file rows */
int sass_params_set_vectors(sass_params *p, const double *xyz, size_t n);
int sass_params_set_moments(sass_params *p, const long *lm, size_t n);
/* run the generators (Params::init: vectors.create(), moments.create(); parameters.cpp:930-1122) */
int sass_params_create(sass_params *p);
size_t sass_params_num_vectors(const sass_params *p);
int sass_params_get_vectors(const sass_params *p, double *out);
size_t sass_params_num_moments(const sass_params *p);
int sass_params_get_moments(const sass_params *p, long *out);

/* ---- run: factory + device.run() ------------------------------------------------------------------------------ */
/* ScatterFactors::update(q)+get_all(): fill b[NA] for |q| = ql (scatter_factors.cpp:56-78) */
typedef void (*sass_factors_fn)(void *user, double ql, double *b, size_t NA);
/* HDF5WriterClient::write(qvector, fqt, NF, fq, fq2) (file_writer_service.cpp:504-515); fq0 = fqt[0] */
typedef void (*sass_write_fn)(void *user, const double q[3], const double *fqt, size_t NF, const double fq[2],
                              const double fq2[2]);
/* frames: host float [NF][NA][3] cartesian.  Either b_const (NA doubles, |q|-independent factors) or ffn is used.
 * has_device (optional out): 0 on ranks the decomposition left spare.  timers (optional out): "key=sum_s:count;..." */
int sass_scatter_run(const sass_params *p, const sass_comm_vtbl *comm, const sass_backend_vtbl *backend, sgpu_ctx *ctx,
                     size_t NA, size_t NF, const float *frames, const double *b_const, sass_factors_fn ffn,
                     void *fuser, const double *qvectors, size_t NQ, sass_write_fn wfn, void *wuser, int *has_device,
                     char *timers, size_t timers_cap);

/* ---- host logic, unit-testable ---------------------------------------------------------------------------------- */
/* sample.motions walkers (reference src/sample/motion_walker.cpp, coordinate_sets.cpp:120-148): the 4x4 transform (row major,
 * applied to row vectors (x, y, z, 1) * T) of frames 0 .. n-1.  type: "linear" | "fixed" | "oscillation" | "randomwalk" |
 * "brownian" | "localbrownian" | "rotationalbrownian" */
int sass_motion_transforms(const char *type, double displace, double frequency, double radius, unsigned long seed, long sampling,
                           const double dir[3], size_t n, double *out /* [n][16] */);
int sass_div_assignment(size_t NN, size_t rank, size_t NAF, size_t *offset, size_t *size, size_t *max);
int sass_mod_assignment(size_t NN, size_t rank, size_t NAF, size_t *offset, size_t *size, size_t *max);
int sass_decomposition_plan(size_t nn, size_t nq, size_t naf, size_t elbytesize, size_t nmaxbytesize, double utilization,
                            int automatic, size_t manual_size, size_t *partitions, size_t *partitionsize,
                            size_t *penalty);
/* scans: [nscans][7] = base x,y,z, from, to, points, exponent.  Returns the count (out may be NULL to size). */
size_t sass_create_from_scans(const double *scans, size_t nscans, double *out, size_t cap);
/* AbstractVectorsScatterDevice::init_subvectors for q (abstract_vectors_scatter_device.cpp:112-175) */
size_t sass_init_subvectors(const sass_params *p, const double q[3], double *out, size_t cap);

/* ---- DCD trajectories (reader: src/sample/frames.cpp:272-436; writer layout: src/stager/coordinate_writer.cpp:37-144) --- */
typedef struct sass_dcd sass_dcd;
/* opens and indexes the file, then applies first/last/stride like FileFrameset::trim_index (frames.cpp:224-245) */
int sass_dcd_open(const char *path, size_t first, size_t last, int last_set, size_t stride, sass_dcd **out);
int sass_dcd_info(const sass_dcd *d, size_t *nframes, size_t *natoms, int *has_unitcell);
/* frames [first, first+count) of the trimmed index -> out[count][natoms][3] (the stager's frame-major layout) */
int sass_dcd_read(sass_dcd *d, size_t first, size_t count, float *out);
void sass_dcd_close(sass_dcd *d);
/* writes xyz[NF][NA][3] as a CHARMM DCD the reference's reader accepts (stager.dump format) */
int sass_dcd_write(const char *path, const float *xyz, size_t NF, size_t NA);

/* ---- GROMACS XTC / TRR trajectories (reference: src/sample/frames.cpp:592-858 through the xdrfile library) ---------
 * format: "xtc" or "trr".  Coordinates come back in Angstrom, (float)(10.0 * nm), frame-major like sass_dcd_read. */
typedef struct sass_xdr sass_xdr;
int sass_xdr_open(const char *path, const char *format, size_t first, size_t last, int last_set, size_t stride,
                  sass_xdr **out);
int sass_xdr_info(const sass_xdr *d, size_t *nframes, size_t *natoms);
/* out[count][natoms][3]; box (may be NULL) [count][9] doubles in Angstrom */
int sass_xdr_read(sass_xdr *d, size_t first, size_t count, float *out, double *box);
void sass_xdr_close(sass_xdr *d);

/* ---- HDF5 signal file ("next" row; reference src/services/file_writer_service.cpp:44-171,314-484) ------
 * A minimal HDF5 container writer/reader (csrc/host/h5mini.cpp): superblock v0, symbol-table groups, float64 datasets
 * extendible along the q-vector dimension and chunked like the reference's, meta/{rawconfig,config,database}.
 * sass_h5_read walks a file and reports every dataset: path ("fqt", "meta/config"), kind (0 float64, 1 char array,
 * 2 string), rank, dims, maxdims (NULL when absent), chunk dims (NULL when contiguous), the values and their byte count. */
typedef void (*sass_h5_dataset_fn)(void *user, const char *path, int kind, size_t rank, const uint64_t *dims,
                                   const uint64_t *maxdims, const uint64_t *chunk, const void *data, size_t nbytes);
int sass_h5_read(const char *path, sass_h5_dataset_fn fn, void *user);
/* writes a signal file in one call: q [n][3], fqt [n][NF][2], fq / fq2 [n][2] (fq0 = fqt[:,0]); flags bit 0..3 = store
 * fqt, fq0, fq, fq2; resume != 0 keeps the rows of an existing file (HDF5WriterService::init) and appends */
int sass_h5_write_signal(const char *path, size_t NF, size_t chunksize, int flags, int resume, const char *rawconfig,
                         const char *config, const char *database, size_t n, const double *q, const double *fqt,
                         const double *fq, const double *fq2, size_t *rows_total);

/* ---- control plane: scatter.xml + db.xml + PDB + DCD -> hot path ("next" row) ------------------------
 * Replaces, for one process, the flow of the reference executable src/main/sassena.cpp:132-417:
 *   Params::init/read_xml   src/control/parameters.cpp:64-792
 *   Database::read_xml      src/control/database.cpp:31-145, evaluation :293-528
 *   Sample::init            src/sample/sample.cpp:30-101 (PDB atoms, selections, framesets)
 *   ScatterFactors          src/scatter_devices/scatter_factors.cpp:28-135
 * Output: <signal_dir>/{qvectors,fqt,fq0,fq,fq2}.npy with the dataset names and shapes of the reference's signal.h5
 * (src/services/file_writer_service.cpp:44-171); complex values are trailing [2] = (re, im). */
typedef struct sass_job sass_job;
int sass_job_load(const char *config_file, sass_job **out);
/* the same with the reference's command-line overwrite options (Params::options / overwrite_options,
 * src/control/parameters.cpp:795-875): n (key, value) pairs applied after the configuration file has been read.  Keys:
 * sample.structure.file, sample.structure.format, stager.target, stager.dump, stager.file, stager.format,
 * scattering.signal.file, limits.computation.threads; any other key is an error ("unrecognised option"). */
int sass_job_load_overwrite(const char *config_file, const char *const *keys, const char *const *values, size_t n,
                            sass_job **out);
void sass_job_free(sass_job *j);
/* counts: all atoms of the structure, atoms of stager.target, frames, q-vectors */
int sass_job_info(const sass_job *j, size_t *natoms, size_t *ntarget, size_t *nframes, size_t *nqvectors);
int sass_job_qvectors(const sass_job *j, double *q_out /* [nqvectors][3] */);
/* b_j(|q|) for the atoms of stager.target (ScatterFactors::update + get_all) */
int sass_job_factors(const sass_job *j, double ql, double *b_out /* [ntarget] */);
/* coordinates as staged: [nframes][ntarget][3]; the pointer stays valid until sass_job_free */
int sass_job_frames(const sass_job *j, const float **frames);
/* atom indices of a named selection; *n receives the full count, at most cap ids are stored */
int sass_job_selection(const sass_job *j, const char *name, size_t *ids, size_t cap, size_t *n);
/* the Params subset the scatter devices read (borrowed; do not free) */
const sass_params *sass_job_params(const sass_job *j);
/* scattering.signal.file resolved against the configuration file's directory (parameters.cpp:41-54; default signal.h5) */
const char *sass_job_signal_file(const sass_job *j);
/* the value in effect (after the configuration file and the overwrite options; file names resolved) of sample.structure.file,
 * sample.structure.format, stager.target, stager.dump ("true" / "false"), stager.file, stager.format, scattering.signal.file;
 * NULL for any other key.  Borrowed, valid until sass_job_free. */
const char *sass_job_option(const sass_job *j, const char *key);
/* the reference's `s_stage` executable (src/main/s_stage.cpp:205-232): stages the trajectory of stager.target the way
 * stager.mode says ("frames" / "atoms") on the ranks of `comm` and, with stager.dump, writes the staged coordinates to
 * stager.file (data_stager.cpp:131-165, 352-391).  *staged_bytes = what this rank staged. */
int sass_job_stage(sass_job *j, const sass_comm_vtbl *comm, const sass_backend_vtbl *backend, sgpu_ctx *ctx, size_t *staged_bytes,
                   char *report, size_t report_cap);
/* runs every q-vector; comm/backend NULL = single process / the in-library CUDA backend */
int sass_job_run(sass_job *j, const char *signal_dir, const sass_comm_vtbl *comm, const sass_backend_vtbl *backend,
                 sgpu_ctx *ctx, size_t *written, char *report, size_t report_cap);

#ifdef __cplusplus
}
#endif
#endif
