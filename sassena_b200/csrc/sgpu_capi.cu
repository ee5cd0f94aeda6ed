// sgpu_capi.cu — implementation of the C-ABI in include/sassena_b200.h.
//
// One sgpu_ctx owns one GPU: the staged coordinates, the per-|q| scratch (amplitudes, FFT work,
// packed partial) and two streams (compute + copy).  All compute is CUDA; there is no CPU path.
#include "../../include/sassena_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels/kernels.hpp"
#include "nccl_api.hpp"

using namespace sass;

namespace {
std::string g_init_error;
const double kTwoOverPi = 0.63661977236758134308;
}  // namespace

struct sgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t conv_stream = nullptr;  // per-chunk coordinate conversions of frames still being staged (created on demand)
    std::string err;
    uint64_t launches = 0;

    // staged coordinates
    int mode = 0;  // 0 none, 1 frames [NF][NA][3], 2 atoms [NA][NF][3]
    float *d_xyz = nullptr;
    bool own_xyz = false;
    size_t xyz_cap = 0;  // bytes owned
    size_t NF = 0, NA = 0;
    // frame window (sgpu_set_frame_window): the staged frames are [f_first, f_first + NF) of a timeline of NFt frames;
    // NFt == NF unless a window is set.  Amplitude kernels run over NF frames, the DSP over NFt.
    size_t NFt = 0, f_first = 0;
    int repr = SGPU_REPR_CARTESIAN;
    struct Chunk {
        size_t f0, nf;
        cudaEvent_t ready;
    };
    std::vector<Chunk> chunks;  // pending async staging chunks (frames mode)
    float *d_stage_tmp = nullptr;
    size_t stage_tmp_cap = 0;
    // double-buffered wave streaming (sgpu_stage_atoms_prefetch / _swap)
    float *d_wave[2] = {nullptr, nullptr};
    size_t wave_cap = 0;           // floats per wave buffer
    int wave_front = 0;            // buffer compute calls read (when wave_staged)
    bool wave_staged = false;      // d_xyz points at d_wave[wave_front]
    bool wave_pending = false;     // a prefetch is queued into the back buffer
    size_t wave_pending_count = 0, wave_pending_NF = 0;
    cudaEvent_t wave_ready[2] = {nullptr, nullptr};  // copy stream: H2D into the buffer done
    cudaEvent_t wave_free[2] = {nullptr, nullptr};   // compute stream: every reader of the buffer queued so far is done

    double *d_b = nullptr;
    size_t nb = 0, b_cap = 0;
    double *d_qs = nullptr;
    size_t q_cap = 0;
    int *d_lm = nullptr;
    size_t lm_cap = 0;
    double *d_qlens = nullptr;  // |q| batch of the multipole path
    size_t qlens_cap = 0;
    double *d_bq = nullptr;     // per-|q| factors [NQ][NA] (sgpu_set_factors_batch)
    size_t bq_cap = 0, bq_nq = 0, bq_n = 0;
    bool bq_uniform = false;    // every |q| row of the batch holds the same factors
    double rmax = 0.0;          // bound of |r| over the staged coordinates (ensure_rmax)
    bool rmax_valid = false;
    int scan_kinds[3] = {0, 0, 0};
    std::vector<double> h_qs;

    double2 *d_A = nullptr;
    size_t A_cap = 0;  // entries
    size_t A_NM = 0;   // timelines held by the last all/mpsphere compute
    void *d_work = nullptr;
    int atoms_dec_R = 0;  // atoms mode: 0 = frames in natural order, R = decimated [r][m] order of the split self path
    double cyl_axis[3] = {0, 0, 0};  // axis the staged frames were converted to cylindrical coordinates with
    size_t work_cap = 0;
    double *d_partial = nullptr;
    size_t partial_cap = 0;
    double2 *d_out = nullptr;
    size_t out_cap = 0;
    double *h_acc = nullptr;  // pinned, 4 doubles
    char *h_up = nullptr;     // pinned + mapped staging area for small uploads (factors, q-vectors, moments)
    size_t up_cap = 0;

    // NCCL communicator of the partition (sgpu_comm_init): exchange of the frame-sharded coherent path, all-reduces
    ncclComm_t comm = nullptr;
    bool comm_owned = true;
    int comm_size = 1, comm_rank = 0;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t comm_pass = nullptr, comm_done = nullptr;
    double2 *d_xloc = nullptr, *d_xrecv = nullptr, *d_xtl = nullptr;  // local amplitudes, received pieces, assembled timelines
    size_t xloc_cap = 0, xrecv_cap = 0, xtl_cap = 0;
    float comm_last_ms = 0.f;

    CorrPlan plan;
    SelfPlan splan;  // fused self path (atoms mode, dsp=autocorrelate)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;
    cudaEvent_t evd0 = nullptr, evd1 = nullptr;  // DSP of the frame-sharded path (separate call)
    bool dsp_split = false;
    bool have_times = false;

    ~sgpu_ctx() {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        if (conv_stream) cudaStreamSynchronize(conv_stream);
        for (auto &c : chunks) cudaEventDestroy(c.ready);
        if (own_xyz && d_xyz) cudaFree(d_xyz);
        if (comm && comm_owned) {
            if (const NcclApi *api = nccl_api(nullptr)) api->CommDestroy(comm);
        }
        if (comm_stream) cudaStreamDestroy(comm_stream);
        if (comm_pass) cudaEventDestroy(comm_pass);
        if (comm_done) cudaEventDestroy(comm_done);
        if (d_xloc) cudaFree(d_xloc);
        if (d_xrecv) cudaFree(d_xrecv);
        if (d_xtl) cudaFree(d_xtl);
        if (d_stage_tmp) cudaFree(d_stage_tmp);
        for (int i = 0; i < 2; i++) {
            if (d_wave[i]) cudaFree(d_wave[i]);
            if (wave_ready[i]) cudaEventDestroy(wave_ready[i]);
            if (wave_free[i]) cudaEventDestroy(wave_free[i]);
        }
        if (d_b) cudaFree(d_b);
        if (d_qs) cudaFree(d_qs);
        if (d_lm) cudaFree(d_lm);
        if (d_qlens) cudaFree(d_qlens);
        if (d_bq) cudaFree(d_bq);
        if (d_A) cudaFree(d_A);
        if (d_work) cudaFree(d_work);
        if (d_partial) cudaFree(d_partial);
        if (d_out) cudaFree(d_out);
        if (h_acc) cudaFreeHost(h_acc);
        if (h_up) cudaFreeHost(h_up);
        corr_plan_destroy(&plan);
        self_plan_destroy(&splan);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev2) cudaEventDestroy(ev2);
        if (tm0) cudaEventDestroy(tm0);
        if (tm1) cudaEventDestroy(tm1);
        if (evd0) cudaEventDestroy(evd0);
        if (evd1) cudaEventDestroy(evd1);
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (conv_stream) cudaStreamDestroy(conv_stream);
    }
};

namespace {

int fail(sgpu_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    else g_init_error = msg;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            cudaGetLastError();                                                                           \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? SGPU_ENOMEM : SGPU_ECUDA,                 \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                             \
        }                                                                                                 \
    } while (0)

template <typename T>
int ensure(sgpu_ctx *ctx, T **p, size_t *cap, size_t need_elems) {
    if (*cap >= need_elems && *p) return SGPU_OK;
    if (*p) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(*p));
        *p = nullptr;
        *cap = 0;
    }
    size_t n = std::max<size_t>(need_elems, 1);
    CK(cudaMalloc(reinterpret_cast<void **>(p), n * sizeof(T)));
    *cap = n;
    return SGPU_OK;
}

int ensure_work(sgpu_ctx *ctx, size_t bytes) {
    char *p = reinterpret_cast<char *>(ctx->d_work);
    int rc = ensure<char>(ctx, &p, &ctx->work_cap, bytes);
    ctx->d_work = p;
    return rc;
}

// Small host->device upload that does not touch the DMA copy engine: the bytes go through a mapped pinned buffer and a
// copy kernel on the compute stream.  (A cudaMemcpyAsync would queue behind the stager's multi-GB chunk copies on the
// H2D engine and serialise staging with compute.)  Synchronous with respect to the host.
int small_upload(sgpu_ctx *ctx, void *d_dst, const void *src, size_t bytes) {
    const size_t words = (bytes + 3) / 4;
    if (ctx->up_cap < words * 4) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_up) CK(cudaFreeHost(ctx->h_up));
        ctx->h_up = nullptr;
        ctx->up_cap = 0;
        const size_t cap = std::max<size_t>(words * 4, (size_t)1 << 20);
        CK(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_up), cap, cudaHostAllocMapped));
        ctx->up_cap = cap;
    }
    memcpy(ctx->h_up, src, bytes);
    void *d_view = nullptr;
    CK(cudaHostGetDevicePointer(&d_view, ctx->h_up, 0));
    ctx->launches += launch_copy_words(d_dst, d_view, words, ctx->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return SGPU_OK;
}

void drop_chunks(sgpu_ctx *ctx) {
    for (auto &c : ctx->chunks) cudaEventDestroy(c.ready);
    ctx->chunks.clear();
}

int release_xyz(sgpu_ctx *ctx) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->conv_stream) CK(cudaStreamSynchronize(ctx->conv_stream));
    drop_chunks(ctx);
    if (!ctx->own_xyz) {
        ctx->d_xyz = nullptr;
        ctx->xyz_cap = 0;
    }
    ctx->wave_staged = false;
    ctx->mode = 0;
    ctx->rmax_valid = false;
    ctx->atoms_dec_R = 0;
    return SGPU_OK;
}

// Atoms mode, owned buffer: bring the frames of every atom into decimated order for R (R > 0) or back to natural order
// (R == 0).  Out of place through a bounce buffer, a batch of atoms at a time.  Adopted (caller-owned) buffers stay natural.
// atoms mode: sgpu_stage_atoms copies in chunks of atoms on the copy stream (ctx->chunks holds (first atom, count, event)).
// Consumers that need every atom make the compute stream wait for the last chunk (the copy stream is in order); only the
// autocorrelation path of sgpu_compute_self_vectors_partial walks the chunks one by one.
int atoms_landed(sgpu_ctx *ctx) {
    if (ctx->mode != 2 || ctx->chunks.empty()) return SGPU_OK;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->chunks.back().ready, 0));
    drop_chunks(ctx);
    return SGPU_OK;
}

// rows [a0, a0 + na) of an owned atom-major buffer from natural frame order into the decimated order for R (out of place
// through the bounce buffer, on the compute stream)
int decimate_rows(sgpu_ctx *ctx, size_t a0, size_t na, int R) {
    const size_t row = ctx->NF * 3 * sizeof(float);
    const size_t batch = std::max<size_t>(1, std::min(na, ((size_t)256 << 20) / row));
    int rc = ensure<float>(ctx, &ctx->d_stage_tmp, &ctx->stage_tmp_cap, batch * ctx->NF * 3);
    if (rc) return rc;
    for (size_t b0 = 0; b0 < na; b0 += batch) {
        const size_t nb = std::min(batch, na - b0);
        float *rows = ctx->d_xyz + (a0 + b0) * ctx->NF * 3;
        CK(cudaMemcpyAsync(ctx->d_stage_tmp, rows, nb * row, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->launches += self_decimate_layout(ctx->d_stage_tmp, rows, nb, ctx->NF, R, 1, ctx->stream);
    }
    CK(cudaGetLastError());
    return SGPU_OK;
}

int set_atoms_layout(sgpu_ctx *ctx, int R) {
    if (ctx->mode != 2) return SGPU_OK;
    int rc = atoms_landed(ctx);
    if (rc) return rc;
    if (ctx->atoms_dec_R == R) return SGPU_OK;
    if (!ctx->own_xyz && !ctx->wave_staged) return SGPU_OK;
    const size_t row = ctx->NF * 3 * sizeof(float);
    const size_t batch = std::max<size_t>(1, std::min(ctx->NA, ((size_t)256 << 20) / row));
    rc = ensure<float>(ctx, &ctx->d_stage_tmp, &ctx->stage_tmp_cap, batch * ctx->NF * 3);
    if (rc) return rc;
    for (int pass = 0; pass < 2; pass++) {
        const int cur = (pass == 0) ? ctx->atoms_dec_R : 0;
        const int want = (pass == 0) ? 0 : R;  // first back to natural (if decimated for another R), then to R
        if (cur == want || (pass == 0 && cur == 0) || (pass == 1 && R == 0)) continue;
        for (size_t a0 = 0; a0 < ctx->NA; a0 += batch) {
            const size_t na = std::min(batch, ctx->NA - a0);
            float *rows = ctx->d_xyz + a0 * ctx->NF * 3;
            CK(cudaMemcpyAsync(ctx->d_stage_tmp, rows, na * row, cudaMemcpyDeviceToDevice, ctx->stream));
            ctx->launches += self_decimate_layout(ctx->d_stage_tmp, rows, na, ctx->NF, pass == 0 ? cur : R, pass == 0 ? 0 : 1,
                                                  ctx->stream);
        }
    }
    CK(cudaGetLastError());
    ctx->atoms_dec_R = R;
    return SGPU_OK;
}

int own_xyz_buffer(sgpu_ctx *ctx, size_t bytes) {
    if (!ctx->own_xyz || ctx->xyz_cap < bytes || !ctx->d_xyz) {
        if (ctx->own_xyz && ctx->d_xyz) CK(cudaFree(ctx->d_xyz));
        ctx->d_xyz = nullptr;
        ctx->own_xyz = true;
        ctx->xyz_cap = 0;
        CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_xyz), std::max<size_t>(bytes, 4)));
        ctx->xyz_cap = bytes;
    }
    return SGPU_OK;
}

int ensure_self_plan(sgpu_ctx *ctx);

int ensure_plan(sgpu_ctx *ctx) {
    if (ctx->mode == 2) {
        int rc = ensure_self_plan(ctx);
        if (rc) return rc;
    }
    if (ctx->plan.NF == ctx->NFt && ctx->plan.d_tw) return SGPU_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    corr_plan_destroy(&ctx->plan);
    int rc = corr_plan_create(&ctx->plan, ctx->NFt, ctx->stream, &ctx->launches);
    if (rc == 1) return fail(ctx, SGPU_EINVAL, "number of frames not supported by the correlation plan (1 <= NF <= 2^21)");
    if (rc) return fail(ctx, SGPU_ECUDA, std::string("corr_plan_create: ") + cudaGetErrorString(cudaGetLastError()));
    return SGPU_OK;
}

int ensure_self_plan(sgpu_ctx *ctx) {
    if (ctx->splan.NF == ctx->NF && ctx->splan.d_tw) return SGPU_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    self_plan_destroy(&ctx->splan);
    int rc = self_plan_create(&ctx->splan, ctx->NF, ctx->stream, &ctx->launches);
    if (rc == 1) return fail(ctx, SGPU_EINVAL, "number of frames not supported by the self correlation plan");
    if (rc) return fail(ctx, SGPU_ECUDA, std::string("self_plan_create: ") + cudaGetErrorString(cudaGetLastError()));
    return SGPU_OK;
}

// atoms mode + autocorrelate uses the fused self plan (its padded length differs from the four-step plan's)
bool uses_self_plan(const sgpu_ctx *ctx, int dsp_type) { return ctx->mode == 2 && dsp_type == SGPU_DSP_AUTOCORRELATE; }

size_t partial_len(const sgpu_ctx *ctx, int dsp_type) {
    if (uses_self_plan(ctx, dsp_type)) return ctx->splan.L + 4;
    return (dsp_type == SGPU_DSP_AUTOCORRELATE ? ctx->plan.L : 2 * ctx->NFt) + 4;
}

int check_dsp(sgpu_ctx *ctx, int dsp_type, int dsp_method) {
    // reference: all_vectors_scatter_device.cpp:209-229 (Err::write + throw on unknown type/method)
    if (dsp_type != SGPU_DSP_AUTOCORRELATE && dsp_type != SGPU_DSP_SQUARE && dsp_type != SGPU_DSP_PLAIN)
        return fail(ctx, SGPU_EINVAL, "DSP type not understood: scattering.dsp.type == autocorrelate, square, plain");
    if (dsp_type == SGPU_DSP_AUTOCORRELATE && dsp_method != SGPU_METHOD_FFTW && dsp_method != SGPU_METHOD_DIRECT)
        return fail(ctx, SGPU_EINVAL, "Correlation method not understood: scattering.dsp.method == direct, fftw");
    return SGPU_OK;
}

// upload q-vectors pre-scaled to quarter turns, zero padded to `pad` multiples
int upload_q(sgpu_ctx *ctx, const double *qvecs, size_t NM, size_t pad) {
    const size_t NMpad = ((NM + pad - 1) / pad) * pad + 8;  // + slack: exactly-sized tail launches read < 8 past NM
    ctx->h_qs.assign(NMpad * 3, 0.0);
    for (size_t i = 0; i < NM * 3; i++) ctx->h_qs[i] = qvecs[i] * kTwoOverPi;
    int rc = ensure<double>(ctx, &ctx->d_qs, &ctx->q_cap, NMpad * 3);
    if (rc) return rc;
    return small_upload(ctx, ctx->d_qs, ctx->h_qs.data(), NMpad * 3 * sizeof(double));
}

// DSP of nt timelines in d_A (ld = NF) accumulated into the packed partial
int dsp_accumulate(sgpu_ctx *ctx, size_t nt, int dsp_type, double *d_partial, const double2 *d_A = nullptr) {
    if (!d_A) d_A = ctx->d_A;
    if (dsp_type == SGPU_DSP_AUTOCORRELATE) {
        ctx->launches += corr_power_accumulate(&ctx->plan, d_A, ctx->NFt, nt, ctx->d_work, d_partial,
                                               d_partial + ctx->plan.L, ctx->stream);
    } else {
        ctx->launches += dsp_elementwise_accumulate(d_A, ctx->NFt, nt, ctx->NFt, dsp_type == SGPU_DSP_SQUARE,
                                                    reinterpret_cast<double2 *>(d_partial), d_partial + 2 * ctx->NFt,
                                                    ctx->d_work, ctx->stream);
    }
    CK(cudaGetLastError());
    return SGPU_OK;
}

size_t dsp_work_bytes(const sgpu_ctx *ctx, size_t nt, int dsp_type) {
    size_t b = (dsp_type == SGPU_DSP_AUTOCORRELATE) ? corr_work_bytes(&ctx->plan, nt) : dsp_elementwise_work_bytes(nt);
    return std::max(b, corr_work_bytes(&ctx->plan, 1));
}

int zero_partial(sgpu_ctx *ctx, int dsp_type, double *d_partial) {
    if (ctx->mode == 0) return fail(ctx, SGPU_ESTATE, "compute: nothing staged");
    if (!d_partial) return fail(ctx, SGPU_EINVAL, "compute: d_partial is NULL");
    int rc = ensure_plan(ctx);
    if (rc) return rc;
    CK(cudaMemsetAsync(d_partial, 0, partial_len(ctx, dsp_type) * sizeof(double), ctx->stream));
    return SGPU_OK;
}

int ensure_internal_partial(sgpu_ctx *ctx, int dsp_type) {
    return ensure<double>(ctx, &ctx->d_partial, &ctx->partial_cap, partial_len(ctx, dsp_type));
}

}  // namespace

extern "C" {

const char *sgpu_version(void) { return "sassena_b200 0.1 (sm_100a)"; }

int sgpu_init(int device, sgpu_ctx **out) {
    sgpu_ctx *ctx = nullptr;  // for CK/fail: errors go to g_init_error
    if (!out) return fail(nullptr, SGPU_EINVAL, "sgpu_init: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, SGPU_ECUDA,
                    std::string("sgpu_init: no CUDA device available (") + cudaGetErrorString(e) +
                        "); this library has no CPU fallback");
    }
    if (device == -1) {  // the device this process is already bound to
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device < 0 || device >= count) return fail(nullptr, SGPU_EINVAL, "sgpu_init: device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, SGPU_ECUDA,
                    std::string("sgpu_init: device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                        ", this library is built for sm_100a (B200) only");
    sgpu_ctx *c = new sgpu_ctx();
    c->device = device;
    ctx = nullptr;
    cudaError_t e2;
    if ((e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e2 = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e2 = cudaEventCreate(&c->ev0)) != cudaSuccess || (e2 = cudaEventCreate(&c->ev1)) != cudaSuccess ||
        (e2 = cudaEventCreate(&c->ev2)) != cudaSuccess || (e2 = cudaEventCreate(&c->evd0)) != cudaSuccess ||
        (e2 = cudaEventCreate(&c->evd1)) != cudaSuccess ||
        (e2 = cudaHostAlloc(reinterpret_cast<void **>(&c->h_acc), 4 * sizeof(double), cudaHostAllocDefault)) !=
            cudaSuccess) {
        delete c;
        return fail(nullptr, SGPU_ECUDA, std::string("sgpu_init: ") + cudaGetErrorString(e2));
    }
    *out = c;
    return SGPU_OK;
}

void sgpu_destroy(sgpu_ctx *ctx) { delete ctx; }

const char *sgpu_last_error(const sgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

uint64_t sgpu_launch_count(const sgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int sgpu_synchronize(sgpu_ctx *ctx) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->conv_stream) CK(cudaStreamSynchronize(ctx->conv_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SGPU_OK;
}

void *sgpu_stream(sgpu_ctx *ctx) { return ctx ? reinterpret_cast<void *>(ctx->stream) : nullptr; }

int sgpu_host_alloc(void **ptr, size_t bytes) {
    sgpu_ctx *ctx = nullptr;
    if (!ptr) return SGPU_EINVAL;
    CK(cudaHostAlloc(ptr, std::max<size_t>(bytes, 1), cudaHostAllocDefault));
    return SGPU_OK;
}
int sgpu_host_free(void *ptr) {
    sgpu_ctx *ctx = nullptr;
    CK(cudaFreeHost(ptr));
    return SGPU_OK;
}
int sgpu_device_alloc(void **d_ptr, size_t bytes) {
    sgpu_ctx *ctx = nullptr;
    if (!d_ptr) return SGPU_EINVAL;
    CK(cudaMalloc(d_ptr, std::max<size_t>(bytes, 1)));
    return SGPU_OK;
}
int sgpu_device_free(void *d_ptr) {
    sgpu_ctx *ctx = nullptr;
    CK(cudaFree(d_ptr));
    return SGPU_OK;
}
int sgpu_memcpy_d2h(sgpu_ctx *ctx, void *dst, const void *d_src, size_t bytes) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SGPU_OK;
}
int sgpu_memcpy_h2d(sgpu_ctx *ctx, void *d_dst, const void *src, size_t bytes) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SGPU_OK;
}

/* ---- staging ---------------------------------------------------------------------------------- */

int sgpu_stage_frames(sgpu_ctx *ctx, const float *xyz, size_t NF, size_t NA, int repr) {
    if (!ctx) return SGPU_EINVAL;
    if (!xyz || NF < 1 || NA < 1)
        return fail(ctx, SGPU_EINVAL, "sgpu_stage_frames: No frames / atoms available");  // factory.cpp:39-47
    if (repr != SGPU_REPR_CARTESIAN && repr != SGPU_REPR_SPHERICAL)
        return fail(ctx, SGPU_EINVAL, "sgpu_stage_frames: unknown representation");
    CK(cudaSetDevice(ctx->device));
    int rc = release_xyz(ctx);
    if (rc) return rc;
    const size_t frame_bytes = NA * 3 * sizeof(float);
    rc = own_xyz_buffer(ctx, NF * frame_bytes);
    if (rc) return rc;
    // chunk by frames.  Every chunk becomes one amplitude launch that waits for its copy, and every launch ends in a
    // partially filled last wave, so few large chunks are better for the kernel while a small FIRST chunk lets it start
    // early: sizes grow geometrically (32 MB, x2 per chunk, up to 2 GB).  SASSENA_STAGE_CHUNK_MB fixes the size.
    size_t chunk_mb = 32, chunk_mb_max = 2048;
    if (const char *e = getenv("SASSENA_STAGE_CHUNK_MB")) chunk_mb = chunk_mb_max = std::max<size_t>(1, strtoull(e, nullptr, 10));
    size_t nfc = std::max<size_t>(1, (chunk_mb << 20) / frame_bytes);
    const size_t nfc_max = std::max<size_t>(1, (chunk_mb_max << 20) / frame_bytes);
    for (size_t f0 = 0, nf = 0; f0 < NF; f0 += nf, nfc = std::min(2 * nfc, nfc_max)) {
        nf = std::min(nfc, NF - f0);
        sgpu_ctx::Chunk c{f0, nf, nullptr};
        CK(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
        cudaError_t e = cudaMemcpyAsync(ctx->d_xyz + f0 * NA * 3, xyz + f0 * NA * 3, nf * frame_bytes,
                                        cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(c.ready, ctx->copy_stream);
        ctx->chunks.push_back(c);
        if (e != cudaSuccess) {
            drop_chunks(ctx);
            return fail(ctx, SGPU_ECUDA, std::string("sgpu_stage_frames: ") + cudaGetErrorString(e));
        }
    }
    ctx->mode = 1;
    ctx->NF = NF;
    ctx->NFt = NF;
    ctx->f_first = 0;
    ctx->NA = NA;
    ctx->repr = repr;
    return SGPU_OK;
}

int sgpu_stage_frames_device(sgpu_ctx *ctx, const float *d_xyz, size_t NF, size_t NA, int repr) {
    if (!ctx) return SGPU_EINVAL;
    if (!d_xyz || NF < 1 || NA < 1) return fail(ctx, SGPU_EINVAL, "sgpu_stage_frames_device: No frames / atoms available");
    CK(cudaSetDevice(ctx->device));
    int rc = release_xyz(ctx);
    if (rc) return rc;
    if (ctx->own_xyz && ctx->d_xyz) CK(cudaFree(ctx->d_xyz));
    ctx->own_xyz = false;
    ctx->xyz_cap = 0;
    ctx->d_xyz = const_cast<float *>(d_xyz);
    ctx->mode = 1;
    ctx->NF = NF;
    ctx->NFt = NF;
    ctx->f_first = 0;
    ctx->NA = NA;
    ctx->repr = repr;
    return SGPU_OK;
}

int sgpu_frames_to_spherical(sgpu_ctx *ctx) {
    if (!ctx) return SGPU_EINVAL;
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_frames_to_spherical: no frames staged");
    if (ctx->repr == SGPU_REPR_SPHERICAL) return SGPU_OK;
    if (ctx->repr != SGPU_REPR_CARTESIAN) return fail(ctx, SGPU_ESTATE, "sgpu_frames_to_spherical: staged frames are not cartesian");
    if (!ctx->own_xyz) return fail(ctx, SGPU_ESTATE, "sgpu_frames_to_spherical: adopted device buffers are read-only");
    CK(cudaSetDevice(ctx->device));
    if (ctx->chunks.empty()) {
        ctx->launches += launch_cart_to_spherical(ctx->d_xyz, ctx->NF * ctx->NA, ctx->stream);
    } else {
        // staging chunks still in flight: convert every chunk as it lands, on a stream of its own (on the compute stream the
        // conversions would queue IN FRONT of the multipole launches and hold all of them back until the last chunk has
        // crossed PCIe), and let the chunk's event from now on mean "converted": the multipole kernel starts on the first
        // chunks while the last ones are still being copied
        if (!ctx->conv_stream) CK(cudaStreamCreateWithFlags(&ctx->conv_stream, cudaStreamNonBlocking));
        for (auto &c : ctx->chunks) {
            CK(cudaStreamWaitEvent(ctx->conv_stream, c.ready, 0));
            ctx->launches += launch_cart_to_spherical(ctx->d_xyz + c.f0 * ctx->NA * 3, c.nf * ctx->NA, ctx->conv_stream);
            CK(cudaEventRecord(c.ready, ctx->conv_stream));
        }
    }
    CK(cudaGetLastError());
    ctx->repr = SGPU_REPR_SPHERICAL;
    return SGPU_OK;
}

// CartesianVectorBase(axis) (reference src/math/coor3d.cpp:278-298): rows e_r, e_phi, e_z
static bool vector_base(const double axis[3], double base[9]) {
    auto len = [](const double *v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    auto cross = [](const double *a, const double *b, double *o) {
        o[0] = a[1] * b[2] - a[2] * b[1];
        o[1] = a[2] * b[0] - a[0] * b[2];
        o[2] = a[0] * b[1] - a[1] * b[0];
    };
    const double al = len(axis);
    if (!(al > 0.0)) return false;
    double ez[3] = {axis[0] / al, axis[1] / al, axis[2] / al};
    const double ek[3] = {0, 0, 1}, ej[3] = {0, 1, 0};
    double t[3];
    cross(ek, ez, t);
    if (len(t) == 0) cross(ej, ez, t);
    const double tl = len(t);
    double er[3] = {t[0] / tl, t[1] / tl, t[2] / tl};
    double u[3];
    cross(ez, er, u);
    const double ul = len(u);
    for (int c = 0; c < 3; c++) {
        base[c] = er[c];
        base[3 + c] = u[c] / ul;
        base[6 + c] = ez[c];
    }
    return true;
}

int sgpu_frames_to_cylindrical(sgpu_ctx *ctx, const double axis[3]) {
    if (!ctx) return SGPU_EINVAL;
    if (!axis) return fail(ctx, SGPU_EINVAL, "sgpu_frames_to_cylindrical: NULL axis");
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_frames_to_cylindrical: no frames staged");
    if (ctx->repr != SGPU_REPR_CARTESIAN)
        return fail(ctx, SGPU_ESTATE, "sgpu_frames_to_cylindrical: staged frames are not cartesian");
    if (!ctx->own_xyz) return fail(ctx, SGPU_ESTATE, "sgpu_frames_to_cylindrical: adopted device buffers are read-only");
    double base[9];
    if (!vector_base(axis, base)) return fail(ctx, SGPU_EINVAL, "sgpu_frames_to_cylindrical: axis has zero length");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->conv_stream) CK(cudaStreamSynchronize(ctx->conv_stream));
    drop_chunks(ctx);
    ctx->launches += launch_cart_to_cylindrical(ctx->d_xyz, ctx->NF * ctx->NA, base, ctx->stream);
    CK(cudaGetLastError());
    ctx->repr = SGPU_REPR_CYLINDRICAL;
    for (int i = 0; i < 3; i++) ctx->cyl_axis[i] = axis[i];
    return SGPU_OK;
}

int sgpu_stage_atoms(sgpu_ctx *ctx, const float *xyz, size_t NA_local, size_t NF) {
    if (!ctx) return SGPU_EINVAL;
    if (!xyz || NF < 1 || NA_local < 1) return fail(ctx, SGPU_EINVAL, "sgpu_stage_atoms: No frames / atoms available");
    CK(cudaSetDevice(ctx->device));
    int rc = release_xyz(ctx);
    if (rc) return rc;
    const size_t bytes = NA_local * NF * 3 * sizeof(float);
    rc = own_xyz_buffer(ctx, bytes);
    if (rc) return rc;
    // chunks of atoms on the copy stream, growing 64 MB ... 1 GB: the autocorrelation path starts on the first chunk while
    // the rest is still crossing PCIe (release_xyz above has synchronised both streams, so nothing reads the buffer any more)
    const size_t row = NF * 3 * sizeof(float);
    size_t nac = std::max<size_t>(1, ((size_t)64 << 20) / row);
    const size_t nac_max = std::max<size_t>(1, ((size_t)1024 << 20) / row);
    for (size_t a0 = 0, na = 0; a0 < NA_local; a0 += na, nac = std::min(2 * nac, nac_max)) {
        na = std::min(nac, NA_local - a0);
        sgpu_ctx::Chunk c{a0, na, nullptr};
        CK(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
        cudaError_t e = cudaMemcpyAsync(ctx->d_xyz + a0 * NF * 3, xyz + a0 * NF * 3, na * row, cudaMemcpyHostToDevice,
                                        ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(c.ready, ctx->copy_stream);
        ctx->chunks.push_back(c);
        if (e != cudaSuccess) {
            drop_chunks(ctx);
            return fail(ctx, SGPU_ECUDA, std::string("sgpu_stage_atoms: ") + cudaGetErrorString(e));
        }
    }
    ctx->mode = 2;
    ctx->NF = NF;
    ctx->NFt = NF;
    ctx->f_first = 0;
    ctx->NA = NA_local;
    ctx->repr = SGPU_REPR_CARTESIAN;
    return SGPU_OK;
}

int sgpu_stage_atoms_device(sgpu_ctx *ctx, const float *d_xyz, size_t NA_local, size_t NF) {
    if (!ctx) return SGPU_EINVAL;
    if (!d_xyz || NF < 1 || NA_local < 1) return fail(ctx, SGPU_EINVAL, "sgpu_stage_atoms_device: No frames / atoms available");
    CK(cudaSetDevice(ctx->device));
    int rc = release_xyz(ctx);
    if (rc) return rc;
    if (ctx->own_xyz && ctx->d_xyz) CK(cudaFree(ctx->d_xyz));
    ctx->own_xyz = false;
    ctx->xyz_cap = 0;
    ctx->d_xyz = const_cast<float *>(d_xyz);
    ctx->mode = 2;
    ctx->NF = NF;
    ctx->NFt = NF;
    ctx->f_first = 0;
    ctx->NA = NA_local;
    ctx->repr = SGPU_REPR_CARTESIAN;
    return SGPU_OK;
}

// atoms atom_first + i*stride, i < count, of frame-major host input -> atom-major device layout (transpose on the GPU)
static int stage_atoms_strided(sgpu_ctx *ctx, const float *xyz, size_t NF, size_t NA, size_t atom_first, size_t stride,
                               size_t count) {
    CK(cudaSetDevice(ctx->device));
    int rc = release_xyz(ctx);
    if (rc) return rc;
    rc = own_xyz_buffer(ctx, count * NF * 3 * sizeof(float));
    if (rc) return rc;
    const size_t frame_bytes = NA * 3 * sizeof(float);
    size_t nfc = std::max<size_t>(32, ((size_t)256 << 20) / frame_bytes);
    nfc = std::min(nfc, NF);
    // two bounce buffers so the H2D copy of chunk i+1 overlaps the transpose of chunk i
    rc = ensure<float>(ctx, &ctx->d_stage_tmp, &ctx->stage_tmp_cap, 2 * nfc * NA * 3);
    if (rc) return rc;
    cudaEvent_t copied[2], consumed[2];
    for (int i = 0; i < 2; i++) {
        CK(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
    }
    int slot = 0;
    size_t iter = 0;
    for (size_t f0 = 0; f0 < NF; f0 += nfc, slot ^= 1, iter++) {
        const size_t nf = std::min(nfc, NF - f0);
        float *tmp = ctx->d_stage_tmp + (size_t)slot * nfc * NA * 3;
        if (iter >= 2) CK(cudaStreamWaitEvent(ctx->copy_stream, consumed[slot], 0));
        CK(cudaMemcpyAsync(tmp, xyz + f0 * NA * 3, nf * frame_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(copied[slot], ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->stream, copied[slot], 0));
        ctx->launches += launch_frames_to_atoms(tmp, ctx->d_xyz, NF, nf, f0, NA, atom_first, stride, count, ctx->stream);
        CK(cudaEventRecord(consumed[slot], ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 2; i++) {
        cudaEventDestroy(copied[i]);
        cudaEventDestroy(consumed[i]);
    }
    CK(cudaGetLastError());
    ctx->mode = 2;
    ctx->NF = NF;
    ctx->NFt = NF;
    ctx->f_first = 0;
    ctx->NA = count;
    ctx->repr = SGPU_REPR_CARTESIAN;
    return SGPU_OK;
}

int sgpu_stage_atoms_from_frames(sgpu_ctx *ctx, const float *xyz, size_t NF, size_t NA, size_t nranks, size_t rank) {
    if (!ctx) return SGPU_EINVAL;
    if (!xyz || NF < 1 || NA < 1) return fail(ctx, SGPU_EINVAL, "sgpu_stage_atoms_from_frames: No frames / atoms available");
    if (nranks < 1 || rank >= nranks) return fail(ctx, SGPU_EINVAL, "sgpu_stage_atoms_from_frames: bad rank / nranks");
    // ModAssignment::size (assignment.cpp:82-90)
    size_t NA_local = NA / nranks;
    if ((NA % nranks) != 0 && rank < (NA - nranks * (NA / nranks))) NA_local += 1;
    if (NA_local == 0) return fail(ctx, SGPU_EINVAL, "sgpu_stage_atoms_from_frames: this rank owns no atoms");
    return stage_atoms_strided(ctx, xyz, NF, NA, rank, nranks, NA_local);
}

int sgpu_stage_atoms_wave(sgpu_ctx *ctx, const float *xyz, size_t NF, size_t NA, size_t atom_first, size_t atom_stride,
                          size_t count) {
    if (!ctx) return SGPU_EINVAL;
    if (!xyz || NF < 1 || NA < 1) return fail(ctx, SGPU_EINVAL, "sgpu_stage_atoms_wave: No frames / atoms available");
    if (atom_stride < 1 || count < 1 || atom_first + (count - 1) * atom_stride >= NA)
        return fail(ctx, SGPU_EINVAL, "sgpu_stage_atoms_wave: atom range outside the trajectory");
    return stage_atoms_strided(ctx, xyz, NF, NA, atom_first, atom_stride, count);
}

int sgpu_stage_atoms_prefetch(sgpu_ctx *ctx, const float *xyz, size_t count, size_t NF) {
    if (!ctx) return SGPU_EINVAL;
    if (!xyz || NF < 1 || count < 1) return fail(ctx, SGPU_EINVAL, "sgpu_stage_atoms_prefetch: No frames / atoms available");
    if (ctx->wave_pending) return fail(ctx, SGPU_ESTATE, "sgpu_stage_atoms_prefetch: a prefetched wave is waiting for sgpu_stage_atoms_swap");
    CK(cudaSetDevice(ctx->device));
    const size_t need = count * NF * 3;
    if (need > ctx->wave_cap) {
        // grow both buffers: nothing may be in flight on either
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->copy_stream));
        if (ctx->conv_stream) CK(cudaStreamSynchronize(ctx->conv_stream));
        if (ctx->wave_staged) {
            ctx->wave_staged = false;
            ctx->d_xyz = nullptr;
            ctx->mode = 0;
        }
        for (int i = 0; i < 2; i++) {
            if (ctx->d_wave[i]) CK(cudaFree(ctx->d_wave[i]));
            ctx->d_wave[i] = nullptr;
        }
        ctx->wave_cap = 0;
        for (int i = 0; i < 2; i++) CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_wave[i]), need * sizeof(float)));
        ctx->wave_cap = need;
    }
    for (int i = 0; i < 2; i++) {
        if (!ctx->wave_ready[i]) CK(cudaEventCreateWithFlags(&ctx->wave_ready[i], cudaEventDisableTiming));
        if (!ctx->wave_free[i]) CK(cudaEventCreateWithFlags(&ctx->wave_free[i], cudaEventDisableTiming));
    }
    const int back = ctx->wave_staged ? (ctx->wave_front ^ 1) : ctx->wave_front;
    // readers of the back buffer (the wave before the staged one) were all queued before the swap that retired it
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->wave_free[back], 0));
    CK(cudaMemcpyAsync(ctx->d_wave[back], xyz, need * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->wave_ready[back], ctx->copy_stream));
    ctx->wave_pending = true;
    ctx->wave_pending_count = count;
    ctx->wave_pending_NF = NF;
    return SGPU_OK;
}

int sgpu_stage_atoms_swap(sgpu_ctx *ctx) {
    if (!ctx) return SGPU_EINVAL;
    if (!ctx->wave_pending) return fail(ctx, SGPU_ESTATE, "sgpu_stage_atoms_swap: no prefetched wave (sgpu_stage_atoms_prefetch first)");
    CK(cudaSetDevice(ctx->device));
    const int back = ctx->wave_staged ? (ctx->wave_front ^ 1) : ctx->wave_front;
    if (ctx->wave_staged) {
        // everything queued so far on the compute stream may read the front buffer: it is free once that has run
        CK(cudaEventRecord(ctx->wave_free[ctx->wave_front], ctx->stream));
    } else {
        // leaving another staging mode: queued work may still read the old coordinates (no host sync needed for the swap
        // itself, but release_xyz drops chunk events and the frames-mode state)
        int rc = release_xyz(ctx);
        if (rc) return rc;
        if (ctx->own_xyz && ctx->d_xyz) {
            CK(cudaFree(ctx->d_xyz));
            ctx->d_xyz = nullptr;
            ctx->xyz_cap = 0;
        }
        ctx->own_xyz = false;
    }
    CK(cudaStreamWaitEvent(ctx->stream, ctx->wave_ready[back], 0));
    ctx->wave_front = back;
    ctx->wave_staged = true;
    ctx->wave_pending = false;
    ctx->d_xyz = ctx->d_wave[back];
    ctx->own_xyz = false;
    ctx->xyz_cap = 0;
    ctx->mode = 2;
    ctx->NF = ctx->wave_pending_NF;
    ctx->NFt = ctx->NF;
    ctx->f_first = 0;
    ctx->NA = ctx->wave_pending_count;
    ctx->repr = SGPU_REPR_CARTESIAN;
    ctx->atoms_dec_R = 0;  // a fresh wave arrives in natural frame order
    ctx->rmax_valid = false;
    ctx->nb = 0;           // factors belong to the previous wave's atoms
    return SGPU_OK;
}

int sgpu_staged_shape(const sgpu_ctx *ctx, size_t *NF, size_t *NA, size_t *NF_total) {
    if (!ctx) return SGPU_EINVAL;
    if (NF) *NF = ctx->mode ? ctx->NF : 0;
    if (NA) *NA = ctx->mode ? ctx->NA : 0;
    if (NF_total) *NF_total = ctx->mode ? ctx->NFt : 0;
    return SGPU_OK;
}

int sgpu_device_bytes(sgpu_ctx *ctx, size_t *bytes) {
    if (!ctx || !bytes) return SGPU_EINVAL;
    size_t n = 0;
    if (ctx->own_xyz) n += ctx->xyz_cap;
    n += ctx->stage_tmp_cap * sizeof(float) + 2 * ctx->wave_cap * sizeof(float);
    n += ctx->b_cap * sizeof(double) + ctx->q_cap * sizeof(double) + ctx->bq_cap * sizeof(double);
    n += ctx->A_cap * sizeof(double2) + ctx->work_cap + ctx->partial_cap * sizeof(double) + ctx->out_cap * sizeof(double2);
    *bytes = n;
    return SGPU_OK;
}

int sgpu_accumulate(sgpu_ctx *ctx, double *d_dst, const double *d_src, size_t n) {
    if (!ctx) return SGPU_EINVAL;
    if (!d_dst || !d_src) return fail(ctx, SGPU_EINVAL, "sgpu_accumulate: NULL buffer");
    CK(cudaSetDevice(ctx->device));
    ctx->launches += launch_accumulate(d_dst, d_src, n, ctx->stream);
    CK(cudaGetLastError());
    return SGPU_OK;
}

int sgpu_set_factors(sgpu_ctx *ctx, const double *b, size_t n) {
    if (!ctx) return SGPU_EINVAL;
    if (!b || n == 0) return fail(ctx, SGPU_EINVAL, "sgpu_set_factors: empty factors");
    CK(cudaSetDevice(ctx->device));
    int rc = ensure<double>(ctx, &ctx->d_b, &ctx->b_cap, n);
    if (rc) return rc;
    rc = small_upload(ctx, ctx->d_b, b, n * sizeof(double));
    if (rc) return rc;
    ctx->nb = n;
    return SGPU_OK;
}

/* ---- compute ---------------------------------------------------------------------------------- */

int sgpu_partial_len(sgpu_ctx *ctx, int dsp_type, size_t *n_doubles) {
    if (!ctx || !n_doubles) return SGPU_EINVAL;
    if (ctx->mode == 0) return fail(ctx, SGPU_ESTATE, "sgpu_partial_len: nothing staged");
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    rc = ensure_plan(ctx);
    if (rc) return rc;
    *n_doubles = partial_len(ctx, dsp_type);
    return SGPU_OK;
}

static int frames_amplitude_prologue(sgpu_ctx *ctx, const char *who, size_t NM, int dsp_type, double *d_partial) {
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, std::string(who) + ": frames are not staged (stage_frames first)");
    if (ctx->NFt != ctx->NF)
        return fail(ctx, SGPU_ESTATE, std::string(who) + ": a frame window is set; use sgpu_all_vectors_amplitudes / _dsp_partial");
    if (ctx->nb != ctx->NA) return fail(ctx, SGPU_ESTATE, std::string(who) + ": scattering factors not set for the staged atoms");
    if (NM == 0) return fail(ctx, SGPU_EINVAL, std::string(who) + ": No qvectors left to compute");
    if (!d_partial) return fail(ctx, SGPU_EINVAL, std::string(who) + ": d_partial is NULL");
    int rc = ensure_plan(ctx);
    if (rc) return rc;
    rc = ensure<double2>(ctx, &ctx->d_A, &ctx->A_cap, NM * ctx->NF);
    if (rc) return rc;
    rc = ensure_work(ctx, dsp_work_bytes(ctx, NM, dsp_type));
    if (rc) return rc;
    CK(cudaMemsetAsync(d_partial, 0, partial_len(ctx, dsp_type) * sizeof(double), ctx->stream));
    return SGPU_OK;
}

int sgpu_compute_all_vectors_partial(sgpu_ctx *ctx, const double *qvecs, size_t NM, int dsp_type, double *d_partial) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    if (NM == 0) return zero_partial(ctx, dsp_type, d_partial);  // a rank without subvectors contributes zeros
    if (!qvecs) return fail(ctx, SGPU_EINVAL, "sgpu_compute_all_vectors: qvecs is NULL");
    if (ctx->mode == 1 && ctx->repr != SGPU_REPR_CARTESIAN)
        return fail(ctx, SGPU_ESTATE, "sgpu_compute_all_vectors: staged frames are not cartesian");
    rc = frames_amplitude_prologue(ctx, "sgpu_compute_all_vectors", NM, dsp_type, d_partial);
    if (rc) return rc;
    rc = upload_q(ctx, qvecs, NM, (size_t)amplitude_all_qpad());
    if (rc) return rc;

    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    // if every staging chunk has landed, run one launch over all frames; else one launch per chunk
    bool all_ready = true;
    for (auto &c : ctx->chunks)
        if (cudaEventQuery(c.ready) != cudaSuccess) all_ready = false;
    cudaGetLastError();
    if (getenv("SASSENA_FORCE_CHUNKED")) all_ready = ctx->chunks.empty();
    if (all_ready) {
        drop_chunks(ctx);
        ctx->launches += launch_amplitude_all(ctx->d_xyz, ctx->d_b, ctx->d_qs, ctx->d_A, ctx->NF, ctx->NA, NM, 0,
                                              ctx->NF, ctx->stream);
    } else {
        for (auto &c : ctx->chunks) {
            CK(cudaStreamWaitEvent(ctx->stream, c.ready, 0));
            ctx->launches += launch_amplitude_all(ctx->d_xyz, ctx->d_b, ctx->d_qs, ctx->d_A, ctx->NF, ctx->NA, NM, c.f0,
                                                  c.nf, ctx->stream);
        }
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->A_NM = NM;
    rc = dsp_accumulate(ctx, NM, dsp_type, d_partial);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    return SGPU_OK;
}

/* ---- frame-sharded coherent path ----------------------------------------------------------------------------- */

int sgpu_set_frame_window(sgpu_ctx *ctx, size_t NF_total, size_t f_first) {
    if (!ctx) return SGPU_EINVAL;
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_set_frame_window: frames are not staged (stage_frames first)");
    if (f_first + ctx->NF > NF_total || NF_total < 1)
        return fail(ctx, SGPU_EINVAL, "sgpu_set_frame_window: the staged frames do not fit the window");
    ctx->NFt = NF_total;
    ctx->f_first = f_first;
    return SGPU_OK;
}

int sgpu_all_vectors_amplitudes(sgpu_ctx *ctx, const double *qvecs, size_t NM, double *d_amp) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (!qvecs || !d_amp) return fail(ctx, SGPU_EINVAL, "sgpu_all_vectors_amplitudes: NULL argument");
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_all_vectors_amplitudes: frames are not staged (stage_frames first)");
    if (ctx->repr != SGPU_REPR_CARTESIAN)
        return fail(ctx, SGPU_ESTATE, "sgpu_all_vectors_amplitudes: staged frames are not cartesian");
    if (ctx->nb != ctx->NA) return fail(ctx, SGPU_ESTATE, "sgpu_all_vectors_amplitudes: scattering factors not set for the staged atoms");
    if (NM == 0) return fail(ctx, SGPU_EINVAL, "sgpu_all_vectors_amplitudes: No qvectors left to compute");
    int rc = upload_q(ctx, qvecs, NM, (size_t)amplitude_all_qpad());
    if (rc) return rc;
    double2 *A = reinterpret_cast<double2 *>(d_amp);
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    // columns outside this rank's window stay zero so that a sum over the ranks assembles the timelines
    if (ctx->NFt != ctx->NF) CK(cudaMemsetAsync(A, 0, NM * ctx->NFt * sizeof(double2), ctx->stream));
    // the kernels index coordinates and amplitudes by the same frame number: shift the coordinate base so that frame
    // f_first of the timeline is the first staged frame (only frames inside the window are ever addressed)
    const float *xyz = ctx->d_xyz - ctx->f_first * ctx->NA * 3;
    bool all_ready = true;
    for (auto &c : ctx->chunks)
        if (cudaEventQuery(c.ready) != cudaSuccess) all_ready = false;
    cudaGetLastError();
    if (getenv("SASSENA_FORCE_CHUNKED")) all_ready = ctx->chunks.empty();
    if (all_ready) {
        drop_chunks(ctx);
        ctx->launches += launch_amplitude_all(xyz, ctx->d_b, ctx->d_qs, A, ctx->NFt, ctx->NA, NM, ctx->f_first, ctx->NF, ctx->stream);
    } else {
        for (auto &c : ctx->chunks) {
            CK(cudaStreamWaitEvent(ctx->stream, c.ready, 0));
            ctx->launches += launch_amplitude_all(xyz, ctx->d_b, ctx->d_qs, A, ctx->NFt, ctx->NA, NM, ctx->f_first + c.f0, c.nf,
                                                  ctx->stream);
        }
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    ctx->A_NM = 0;
    return SGPU_OK;
}

int sgpu_all_vectors_dsp_partial(sgpu_ctx *ctx, const double *d_amp, size_t m_first, size_t m_count, int dsp_type,
                                 double *d_partial) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    if (!d_amp || !d_partial) return fail(ctx, SGPU_EINVAL, "sgpu_all_vectors_dsp_partial: NULL argument");
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_all_vectors_dsp_partial: frames are not staged");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    CK(cudaMemsetAsync(d_partial, 0, partial_len(ctx, dsp_type) * sizeof(double), ctx->stream));
    if (m_count == 0) return SGPU_OK;  // a rank without timelines contributes zeros
    rc = ensure_work(ctx, dsp_work_bytes(ctx, m_count, dsp_type));
    if (rc) return rc;
    CK(cudaEventRecord(ctx->evd0, ctx->stream));
    rc = dsp_accumulate(ctx, m_count, dsp_type, d_partial, reinterpret_cast<const double2 *>(d_amp) + m_first * ctx->NFt);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->evd1, ctx->stream));
    ctx->dsp_split = true;
    return SGPU_OK;
}

/* ---- |q|-scan coherent path ------------------------------------------------------------------------------------ */

namespace {
// bound of |r| over the staged coordinates (sqrt(3) * max |component|), computed once per staging
int ensure_rmax(sgpu_ctx *ctx) {
    if (ctx->rmax_valid) return SGPU_OK;
    CK(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->conv_stream) CK(cudaStreamSynchronize(ctx->conv_stream));  // every staging chunk has to be resident
    drop_chunks(ctx);
    float *d_m = nullptr;
    CK(cudaMalloc(reinterpret_cast<void **>(&d_m), sizeof(float)));
    ctx->launches += launch_max_abs(ctx->d_xyz, ctx->NF * ctx->NA * 3, d_m, ctx->stream);
    float m = 0.f;
    cudaError_t e = cudaMemcpyAsync(&m, d_m, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_m);
    CK(e);
    ctx->rmax = 1.7320508075688772 * (double)m;
    ctx->rmax_valid = true;
    return SGPU_OK;
}

struct ScanPass {
    size_t n0;
    int nq;
    double s0, ds;
    int kind;  // 0 plain scan kernel, 1 corrected scan kernel, 2 general kernel per |q|, 3 plain scan kernel with per-|q| factors
    std::vector<double> kappa;
};

// Splits NQ |q| values into kernel passes.  Exactly equally spaced values take the plain scan kernel; values within a
// small deviation of a progression (the reference builds scans from float-rounded fractions, parameters.cpp:1151) take
// the corrected kernel as long as the neglected third-order phase term (kappa sigma)^3/6 stays below ~5e-12; anything
// else is evaluated one |q| at a time by the general kernel.
int plan_scan(sgpu_ctx *ctx, const double *s, size_t NQ, double vmax, bool uniform_b, std::vector<ScanPass> &plan) {
    plan.clear();
    auto fit = [&](size_t n0, size_t L, double &s0, double &ds, double &epsmax) {
        s0 = s[n0];
        ds = L > 1 ? (s[n0 + L - 1] - s[n0]) / (double)(L - 1) : 0.0;
        epsmax = 0.0;
        for (size_t n = 0; n < L; n++) epsmax = std::max(epsmax, std::fabs(s[n0 + n] - (s0 + (double)n * ds)));
    };
    double smax = 0.0;
    for (size_t n = 0; n < NQ; n++) smax = std::max(smax, std::fabs(s[n]));
    double s0, ds, eps;
    fit(0, NQ, s0, ds, eps);
    const bool exact = eps <= 4.5e-16 * smax;
    if (getenv("SASSENA_SCAN_DISABLE") || NQ < 3) {
        for (size_t n = 0; n < NQ; n++) plan.push_back(ScanPass{n, 1, s[n], 0.0, 2, {}});
        return SGPU_OK;
    }
    if (!uniform_b && !exact) {
        // |q|-dependent factors AND a float-rounded scan: the correction terms would need the factors too; one |q| at a time
        for (size_t n = 0; n < NQ; n++) plan.push_back(ScanPass{n, 1, s[n], 0.0, 2, {}});
        return SGPU_OK;
    }
    const bool force_corr = uniform_b && getenv("SASSENA_SCAN_FORCE_CORR") != nullptr;
    const size_t maxB = (size_t)amplitude_scan_max_pass(!uniform_b ? 2 : (exact && !force_corr) ? 0 : 1);
    if (!exact || force_corr) {
        int rc = ensure_rmax(ctx);
        if (rc) return rc;
    }
    // lays the |q| values out in even passes of at most maxB; fp32d: only if every pass is a near-progression whose
    // first-order correction may be summed in FP32 (scan_sym.cu, CORR = 2): the error of that path is bounded by
    // theta_max * (K^2 ulp32 of the recurrence + FP32 accumulation over NA / 32 terms per lane), accepted below 5e-10
    auto lay_out = [&](size_t maxB_, bool fp32d, std::vector<ScanPass> &out) -> bool {
        out.clear();
        const size_t npass = (NQ + maxB_ - 1) / maxB_;
        size_t n0 = 0;
        for (size_t p = 0; p < npass; p++) {
            const size_t want = (NQ - n0 + (npass - p) - 1) / (npass - p);  // even share of what is left
            // the symmetric scan kernel takes any pass length (2K+1 slots; an even length masks one); the kernel for
            // |q|-dependent factors is instantiated for multiples of 4
            size_t L = std::min<size_t>(std::min<size_t>(uniform_b ? want : ((want + 3) / 4) * 4, maxB_), NQ - n0);
            if (exact && !force_corr) {
                out.push_back(ScanPass{n0, (int)L, s[0] + (double)n0 * ds, ds, uniform_b ? 0 : 3, {}});
            } else {
                double ps0, pds, peps;
                fit(n0, L, ps0, pds, peps);
                const double theta = peps * vmax * ctx->rmax;  // largest phase deviation in radians
                if (fp32d) {
                    const double K = (double)(L / 2);
                    const double bound = theta * (K * K * 7e-8 + 4e-8 * std::sqrt((double)ctx->NA / 32.0));
                    if (L < 3 || theta > 3e-4 || bound > 5e-10) return false;
                }
                if (L >= 3 && theta <= 3e-4) {
                    ScanPass sp{n0, (int)L, ps0, pds, 1, {}};
                    for (size_t n = 0; n < L; n++) sp.kappa.push_back(1.5707963267948966 * (s[n0 + n] - (ps0 + (double)n * pds)));
                    out.push_back(sp);
                } else {
                    for (size_t n = 0; n < L; n++) out.push_back(ScanPass{n0 + n, 1, s[n0 + n], 0.0, 2, {}});
                }
            }
            n0 += L;
        }
        return true;
    };
    // float-rounded scans longer than one FP64-D pass: try the longer FP32-D passes first (fewer passes, 3 instead of 5 FP64
    // instructions per evaluation); SASSENA_SCAN_FP64_CORR keeps the first-order sums in FP64
    if ((!exact || force_corr) && uniform_b && NQ > maxB && !getenv("SASSENA_SCAN_FP64_CORR") &&
        lay_out((size_t)amplitude_scan_sym_max_pass(2), true, plan))
        return SGPU_OK;
    lay_out(maxB, false, plan);
    return SGPU_OK;
}

// amplitudes of NQ |q| values s[n] along fixed directions into A[NQ][NM][NFt] (this rank's frame columns)
extern "C++" {
// compact: A is this rank's block only, [NQ][NM][NF] (row stride = the staged frames), for the exchange of the sharded path;
// otherwise A is [NQ][NM][NFt] with the columns outside the frame window zeroed.  after_pass(n0, nq) runs after the launches
// of every pass have been queued (the sharded path starts that pass's exchange there).
template <class AfterPass>
int scan_amplitudes_into(sgpu_ctx *ctx, const char *who, const double *v, size_t NM, const double *s, size_t NQ, double2 *A,
                         bool compact, AfterPass after_pass) {
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, std::string(who) + ": frames are not staged (stage_frames first)");
    if (ctx->repr != SGPU_REPR_CARTESIAN) return fail(ctx, SGPU_ESTATE, std::string(who) + ": staged frames are not cartesian");
    if (!v || !s || NM == 0 || NQ == 0) return fail(ctx, SGPU_EINVAL, std::string(who) + ": No qvectors left to compute");
    // factors: a per-|q| batch set for exactly this NQ, else the single set
    const bool batch = ctx->bq_nq == NQ && ctx->bq_n == ctx->NA && ctx->d_bq;
    if (!batch && ctx->nb != ctx->NA)
        return fail(ctx, SGPU_ESTATE, std::string(who) + ": scattering factors not set for the staged atoms");
    const bool uniform = !batch || ctx->bq_uniform;
    const double *d_b = batch ? ctx->d_bq : ctx->d_b;
    // the kernels index coordinates and amplitudes by the same (timeline) frame number: compact output shifts the base
    const size_t NFt = compact ? ctx->NF : ctx->NFt, strideQ = NM * NFt;
    if (compact) A -= ctx->f_first;
    else if (ctx->NFt != ctx->NF) CK(cudaMemsetAsync(A, 0, NQ * strideQ * sizeof(double2), ctx->stream));
    const float *xyz = ctx->d_xyz - ctx->f_first * ctx->NA * 3;  // see sgpu_all_vectors_amplitudes
    int rc;
    std::vector<ScanPass> plan;
    double vmax = 0.0;
    for (size_t m = 0; m < NM; m++)
        vmax = std::max(vmax, std::sqrt(v[3 * m] * v[3 * m] + v[3 * m + 1] * v[3 * m + 1] + v[3 * m + 2] * v[3 * m + 2]));
    // factors that differ between the |q| values (X-ray form factors, background) ride along as one row per |q|
    // (bulk copies need NA % 4 == 0, which also aligns the rows; other NA take the cp.async staging path)
    rc = plan_scan(ctx, s, NQ, vmax, uniform, plan);
    if (rc) return rc;
    bool all_ready = true;
    for (auto &c : ctx->chunks)
        if (cudaEventQuery(c.ready) != cudaSuccess) all_ready = false;
    cudaGetLastError();
    if (getenv("SASSENA_FORCE_CHUNKED")) all_ready = ctx->chunks.empty();
    std::vector<sgpu_ctx::Chunk> spans;
    if (all_ready) {
        drop_chunks(ctx);
        spans.push_back(sgpu_ctx::Chunk{0, ctx->NF, nullptr});
    } else {
        spans = ctx->chunks;
    }
    bool dirs_uploaded = false;
    std::vector<double> q(3 * NM);
    ctx->scan_kinds[0] = ctx->scan_kinds[1] = ctx->scan_kinds[2] = 0;
    for (auto &ps : plan) {
        ctx->scan_kinds[ps.kind == 3 ? 0 : ps.kind]++;
        if (ps.kind == 2) {
            for (size_t i = 0; i < 3 * NM; i++) q[i] = ps.s0 * v[i];
            CK(cudaStreamSynchronize(ctx->stream));  // d_qs may still be read by the previous pass
            rc = upload_q(ctx, q.data(), NM, (size_t)amplitude_all_qpad());
            if (rc) return rc;
            dirs_uploaded = false;
            for (auto &c : spans) {
                if (c.ready) CK(cudaStreamWaitEvent(ctx->stream, c.ready, 0));
                ctx->launches += launch_amplitude_all(xyz, d_b + (uniform ? 0 : ps.n0 * ctx->NA), ctx->d_qs, A + ps.n0 * strideQ,
                                                      NFt, ctx->NA, NM, ctx->f_first + c.f0, c.nf, ctx->stream);
            }
            rc = after_pass(ps.n0, (size_t)ps.nq);
            if (rc) return rc;
            continue;
        }
        if (!dirs_uploaded) {
            CK(cudaStreamSynchronize(ctx->stream));
            rc = upload_q(ctx, v, NM, (size_t)amplitude_scan_qpad());  // directions, in quarter turns per unit |q|
            if (rc) return rc;
            dirs_uploaded = true;
        }
        for (auto &c : spans) {
            if (c.ready) CK(cudaStreamWaitEvent(ctx->stream, c.ready, 0));
            const int l = launch_amplitude_scan_pass(xyz, d_b + (ps.kind == 3 ? ps.n0 * ctx->NA : 0), ctx->d_qs, ps.s0, ps.ds, ps.nq,
                                                     ps.kind == 1 ? ps.kappa.data() : nullptr, A + ps.n0 * strideQ, NFt, strideQ,
                                                     ctx->NA, NM, ctx->f_first + c.f0, c.nf, ctx->stream,
                                                     ps.kind == 3 ? ctx->NA : 0);
            if (l < 0) return fail(ctx, SGPU_EINVAL, std::string(who) + ": internal error, scan pass too long");
            ctx->launches += l;
        }
        rc = after_pass(ps.n0, (size_t)ps.nq);
        if (rc) return rc;
    }
    CK(cudaGetLastError());
    return SGPU_OK;
}

}  // extern "C++"

int scan_amplitudes_into(sgpu_ctx *ctx, const char *who, const double *v, size_t NM, const double *s, size_t NQ, double2 *A) {
    return scan_amplitudes_into(ctx, who, v, NM, s, NQ, A, false, [](size_t, size_t) { return 0; });
}

// DivAssignment (reference src/decomposition/assignment.cpp:27-35)
void div_block(size_t NN, size_t rank, size_t N, size_t *off, size_t *cnt) {
    *off = (rank * N) / NN;
    *cnt = ((rank + 1) * N) / NN - *off;
}

#define NCK(call)                                                                                              \
    do {                                                                                                       \
        ncclResult_t r__ = (call);                                                                             \
        if (r__ != ncclSuccess) return fail(ctx, SGPU_ECUDA, std::string(#call) + ": " + api->GetErrorString(r__)); \
    } while (0)

// pieces received from rank s, [s][n][m_local][f of s] -> timelines T[n][m_local][NFt]: the reference's alignpad
// (all_vectors_scatter_device.cpp:186-207) after its all_to_all (:169-184); rows beyond NF need no padding here because the
// column FFT treats them as zeros
__global__ void assemble_timelines_kernel(const double2 *__restrict__ R, double2 *__restrict__ T, size_t NQ, size_t mc, size_t NFt,
                                          int NN) {
    // grid: (frames / 256, NQ * mc); every thread moves one (n, m, f) entry
    const size_t f = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (f >= NFt) return;
    const size_t row = blockIdx.y;  // n * mc + m
    const size_t n = row / mc, m = row - n * mc;
    // rank owning frame f: blocks are contiguous, off_s = s NFt / NN
    int s = (int)(((f + 1) * (size_t)NN - 1) / NFt);
    while ((size_t)s * NFt / NN > f) s--;
    while (((size_t)(s + 1) * NFt) / NN <= f) s++;
    const size_t off = (size_t)s * NFt / NN, cnt = ((size_t)(s + 1) * NFt) / NN - off;
    T[row * NFt + f] = R[NQ * mc * off + (n * mc + m) * cnt + (f - off)];
}

// exchange of the |q| planes [n0, n0 + nq) of the local amplitudes A_loc[NQ][NM][nf] on the communication stream: to rank s
// go the rows of ITS timeline block, from rank s come this rank's rows of ITS frame block (grouped ncclSend / ncclRecv = the
// all_to_all of all_vectors_scatter_device.cpp:181).  The own block is a device copy.
int exchange_planes(sgpu_ctx *ctx, const NcclApi *api, size_t NQ, size_t NM, size_t n0, size_t nq) {
    const size_t NN = (size_t)ctx->comm_size, me = (size_t)ctx->comm_rank, nf = ctx->NF, NFt = ctx->NFt;
    size_t m_off, mc;
    div_block(NN, me, NM, &m_off, &mc);
    CK(cudaEventRecord(ctx->comm_pass, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_pass, 0));
    NCK(api->GroupStart());
    for (size_t s = 0; s < NN; s++) {
        size_t ms_off, ms_cnt, fs_off, fs_cnt;
        div_block(NN, s, NM, &ms_off, &ms_cnt);
        div_block(NN, s, NFt, &fs_off, &fs_cnt);
        for (size_t n = n0; n < n0 + nq; n++) {
            const double2 *src = ctx->d_xloc + (n * NM + ms_off) * nf;           // rows of rank s's timelines, my frames
            double2 *dst = ctx->d_xrecv + NQ * mc * fs_off + n * mc * fs_cnt;    // my rows, rank s's frames
            if (s == me) {
                if (mc > 0 && nf > 0) CK(cudaMemcpyAsync(dst, src, mc * nf * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->comm_stream));
                continue;
            }
            if (ms_cnt > 0 && nf > 0) NCK(api->Send(src, 2 * ms_cnt * nf, ncclFloat64, (int)s, ctx->comm, ctx->comm_stream));
            if (mc > 0 && fs_cnt > 0) NCK(api->Recv(dst, 2 * mc * fs_cnt, ncclFloat64, (int)s, ctx->comm, ctx->comm_stream));
        }
    }
    NCK(api->GroupEnd());
    return SGPU_OK;
}

// after the amplitudes of all NQ planes have been exchanged: assemble this rank's timelines, correlate them, sum the packed
// partials over the ranks (the three boost::mpi::reduce calls of all_vectors_scatter_device.cpp:335-343 as one all-reduce)
int sharded_dsp_and_reduce(sgpu_ctx *ctx, const NcclApi *api, size_t NQ, size_t NM, int dsp_type, double *d_partials) {
    const size_t NN = (size_t)ctx->comm_size, me = (size_t)ctx->comm_rank, NFt = ctx->NFt;
    size_t m_off, mc;
    div_block(NN, me, NM, &m_off, &mc);
    const size_t plen = partial_len(ctx, dsp_type);
    CK(cudaEventRecord(ctx->comm_done, ctx->comm_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->comm_done, 0));
    CK(cudaMemsetAsync(d_partials, 0, NQ * plen * sizeof(double), ctx->stream));
    if (mc > 0) {
        dim3 grid((unsigned)((NFt + 255) / 256), (unsigned)(NQ * mc));
        assemble_timelines_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_xrecv, ctx->d_xtl, NQ, mc, NFt, (int)NN);
        ctx->launches++;
        CK(cudaGetLastError());
        for (size_t n = 0; n < NQ; n++) {
            int rc = dsp_accumulate(ctx, mc, dsp_type, d_partials + n * plen, ctx->d_xtl + n * mc * NFt);
            if (rc) return rc;
        }
    }
    NCK(api->AllReduce(d_partials, d_partials, NQ * plen, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
    return SGPU_OK;
}

int sharded_prologue(sgpu_ctx *ctx, const char *who, const NcclApi **api_out, size_t NQ, size_t NM, int dsp_type, double *d_partials) {
    std::string err;
    const NcclApi *api = nccl_api(&err);
    if (!api) return fail(ctx, SGPU_ESTATE, std::string(who) + ": " + err);
    if (!ctx->comm) return fail(ctx, SGPU_ESTATE, std::string(who) + ": no communicator (sgpu_comm_init first)");
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, std::string(who) + ": frames are not staged (stage_frames first)");
    if (!d_partials) return fail(ctx, SGPU_EINVAL, std::string(who) + ": d_partials is NULL");
    if (NM == 0 || NQ == 0) return fail(ctx, SGPU_EINVAL, std::string(who) + ": No qvectors left to compute");
    size_t f_off, f_cnt;
    div_block((size_t)ctx->comm_size, (size_t)ctx->comm_rank, ctx->NFt, &f_off, &f_cnt);
    if (f_off != ctx->f_first || f_cnt != ctx->NF)
        return fail(ctx, SGPU_ESTATE, std::string(who) + ": the staged frames are not this rank's DivAssignment block of the "
                                                         "timeline (sgpu_set_frame_window)");
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    rc = ensure_plan(ctx);
    if (rc) return rc;
    size_t m_off, mc;
    div_block((size_t)ctx->comm_size, (size_t)ctx->comm_rank, NM, &m_off, &mc);
    rc = ensure<double2>(ctx, &ctx->d_xloc, &ctx->xloc_cap, NQ * NM * ctx->NF);
    if (rc) return rc;
    rc = ensure<double2>(ctx, &ctx->d_xrecv, &ctx->xrecv_cap, NQ * mc * ctx->NFt);
    if (rc) return rc;
    rc = ensure<double2>(ctx, &ctx->d_xtl, &ctx->xtl_cap, NQ * mc * ctx->NFt);
    if (rc) return rc;
    rc = ensure_work(ctx, dsp_work_bytes(ctx, std::max<size_t>(mc, 1), dsp_type));
    if (rc) return rc;
    *api_out = api;
    return SGPU_OK;
}
}  // namespace

int sgpu_all_vectors_scan_amplitudes(sgpu_ctx *ctx, const double *v, size_t NM, const double *s, size_t NQ, double *d_amp) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (!d_amp) return fail(ctx, SGPU_EINVAL, "sgpu_all_vectors_scan_amplitudes: d_amp is NULL");
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = scan_amplitudes_into(ctx, "sgpu_all_vectors_scan_amplitudes", v, NM, s, NQ, reinterpret_cast<double2 *>(d_amp));
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    ctx->A_NM = 0;
    return SGPU_OK;
}

/* ---- NCCL inside the library: the partition's communicator ----------------------------------------------------- */
int sgpu_comm_get_unique_id(char *id128) {
    sgpu_ctx *ctx = nullptr;
    if (!id128) return fail(nullptr, SGPU_EINVAL, "sgpu_comm_get_unique_id: NULL");
    std::string err;
    const NcclApi *api = nccl_api(&err);
    if (!api) return fail(nullptr, SGPU_ESTATE, err);
    ncclUniqueId id;
    NCK(api->GetUniqueId(&id));
    memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return SGPU_OK;
}

int sgpu_comm_init(sgpu_ctx *ctx, const char *id128, int nranks, int rank) {
    if (!ctx) return SGPU_EINVAL;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, SGPU_EINVAL, "sgpu_comm_init: bad arguments");
    std::string err;
    const NcclApi *api = nccl_api(&err);
    if (!api) return fail(ctx, SGPU_ESTATE, "sgpu_comm_init: " + err);
    CK(cudaSetDevice(ctx->device));
    if (ctx->comm) {
        int rc = sgpu_comm_destroy(ctx);
        if (rc) return rc;
    }
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    NCK(api->CommInitRank(&ctx->comm, nranks, id, rank));
    ctx->comm_owned = true;
    ctx->comm_size = nranks;
    ctx->comm_rank = rank;
    if (!ctx->comm_stream) CK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    if (!ctx->comm_pass) CK(cudaEventCreateWithFlags(&ctx->comm_pass, cudaEventDisableTiming));
    if (!ctx->comm_done) CK(cudaEventCreateWithFlags(&ctx->comm_done, cudaEventDisableTiming));
    return SGPU_OK;
}

int sgpu_comm_adopt(sgpu_ctx *ctx, void *nccl_comm, int nranks, int rank) {
    if (!ctx) return SGPU_EINVAL;
    if (!nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, SGPU_EINVAL, "sgpu_comm_adopt: bad arguments");
    std::string err;
    const NcclApi *api = nccl_api(&err);
    if (!api) return fail(ctx, SGPU_ESTATE, "sgpu_comm_adopt: " + err);
    CK(cudaSetDevice(ctx->device));
    int rc = sgpu_comm_destroy(ctx);
    if (rc) return rc;
    ctx->comm = static_cast<ncclComm_t>(nccl_comm);
    ctx->comm_owned = false;
    ctx->comm_size = nranks;
    ctx->comm_rank = rank;
    if (!ctx->comm_stream) CK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    if (!ctx->comm_pass) CK(cudaEventCreateWithFlags(&ctx->comm_pass, cudaEventDisableTiming));
    if (!ctx->comm_done) CK(cudaEventCreateWithFlags(&ctx->comm_done, cudaEventDisableTiming));
    return SGPU_OK;
}

int sgpu_comm_destroy(sgpu_ctx *ctx) {
    if (!ctx) return SGPU_EINVAL;
    if (!ctx->comm) return SGPU_OK;
    const NcclApi *api = nccl_api(nullptr);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->comm_stream) CK(cudaStreamSynchronize(ctx->comm_stream));
    if (api && ctx->comm_owned) NCK(api->CommDestroy(ctx->comm));
    ctx->comm_owned = true;
    ctx->comm = nullptr;
    ctx->comm_size = 1;
    ctx->comm_rank = 0;
    return SGPU_OK;
}

int sgpu_comm_info(sgpu_ctx *ctx, int *nranks, int *rank) {
    if (!ctx) return SGPU_EINVAL;
    if (nranks) *nranks = ctx->comm ? ctx->comm_size : 1;
    if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
    return SGPU_OK;
}

int sgpu_comm_allreduce(sgpu_ctx *ctx, double *d_buf, size_t n) {
    if (!ctx) return SGPU_EINVAL;
    if (!d_buf) return fail(ctx, SGPU_EINVAL, "sgpu_comm_allreduce: NULL buffer");
    if (!ctx->comm) return fail(ctx, SGPU_ESTATE, "sgpu_comm_allreduce: no communicator (sgpu_comm_init first)");
    const NcclApi *api = nccl_api(nullptr);
    CK(cudaSetDevice(ctx->device));
    if (n) NCK(api->AllReduce(d_buf, d_buf, n, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
    return SGPU_OK;
}

int sgpu_compute_all_vectors_scan_sharded(sgpu_ctx *ctx, const double *v, size_t NM, const double *s, size_t NQ, int dsp_type,
                                          double *d_partials) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const NcclApi *api = nullptr;
    int rc = sharded_prologue(ctx, "sgpu_compute_all_vectors_scan_sharded", &api, NQ, NM, dsp_type, d_partials);
    if (rc) return rc;
    if (!v || !s) return fail(ctx, SGPU_EINVAL, "sgpu_compute_all_vectors_scan_sharded: NULL argument");
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    // every pass's planes leave for their owners while the next pass is evaluated
    rc = scan_amplitudes_into(ctx, "sgpu_compute_all_vectors_scan_sharded", v, NM, s, NQ, ctx->d_xloc, true,
                              [&](size_t n0, size_t nq) { return exchange_planes(ctx, api, NQ, NM, n0, nq); });
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    rc = sharded_dsp_and_reduce(ctx, api, NQ, NM, dsp_type, d_partials);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    ctx->A_NM = 0;
    return SGPU_OK;
}

int sgpu_compute_all_vectors_sharded(sgpu_ctx *ctx, const double *qvecs, size_t NM, int dsp_type, double *d_partial) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const NcclApi *api = nullptr;
    int rc = sharded_prologue(ctx, "sgpu_compute_all_vectors_sharded", &api, 1, NM, dsp_type, d_partial);
    if (rc) return rc;
    if (!qvecs) return fail(ctx, SGPU_EINVAL, "sgpu_compute_all_vectors_sharded: qvecs is NULL");
    if (ctx->repr != SGPU_REPR_CARTESIAN) return fail(ctx, SGPU_ESTATE, "sgpu_compute_all_vectors_sharded: staged frames are not cartesian");
    if (ctx->nb != ctx->NA) return fail(ctx, SGPU_ESTATE, "sgpu_compute_all_vectors_sharded: scattering factors not set for the staged atoms");
    rc = upload_q(ctx, qvecs, NM, (size_t)amplitude_all_qpad());
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const float *xyz = ctx->d_xyz - ctx->f_first * ctx->NA * 3;
    double2 *A = ctx->d_xloc - ctx->f_first;
    for (auto &c : ctx->chunks) CK(cudaStreamWaitEvent(ctx->stream, c.ready, 0));
    ctx->launches += launch_amplitude_all(xyz, ctx->d_b, ctx->d_qs, A, ctx->NF, ctx->NA, NM, ctx->f_first, ctx->NF, ctx->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    rc = exchange_planes(ctx, api, 1, NM, 0, 1);
    if (rc) return rc;
    rc = sharded_dsp_and_reduce(ctx, api, 1, NM, dsp_type, d_partial);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    ctx->A_NM = 0;
    return SGPU_OK;
}

int sgpu_compute_all_vectors_scan_partial(sgpu_ctx *ctx, const double *v, size_t NM, const double *s, size_t NQ, int dsp_type,
                                          double *d_partials) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    if (!d_partials) return fail(ctx, SGPU_EINVAL, "sgpu_compute_all_vectors_scan: d_partials is NULL");
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_all_vectors_scan: frames are not staged (stage_frames first)");
    if (ctx->NFt != ctx->NF)
        return fail(ctx, SGPU_ESTATE, "sgpu_compute_all_vectors_scan: a frame window is set; use sgpu_all_vectors_scan_amplitudes");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    const size_t plen = partial_len(ctx, dsp_type);
    if (NM == 0) {  // a rank without subvectors contributes zeros
        CK(cudaMemsetAsync(d_partials, 0, NQ * plen * sizeof(double), ctx->stream));
        return SGPU_OK;
    }
    rc = ensure<double2>(ctx, &ctx->d_A, &ctx->A_cap, NQ * NM * ctx->NF);
    if (rc) return rc;
    rc = ensure_work(ctx, dsp_work_bytes(ctx, NM, dsp_type));
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    rc = scan_amplitudes_into(ctx, "sgpu_compute_all_vectors_scan", v, NM, s, NQ, ctx->d_A);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->A_NM = (NQ == 1) ? NM : 0;
    CK(cudaMemsetAsync(d_partials, 0, NQ * plen * sizeof(double), ctx->stream));
    for (size_t n = 0; n < NQ; n++) {
        rc = dsp_accumulate(ctx, NM, dsp_type, d_partials + n * plen, ctx->d_A + n * NM * ctx->NF);
        if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    return SGPU_OK;
}

int sgpu_compute_all_vectors_scan(sgpu_ctx *ctx, const double *v, size_t NM, const double *s, size_t NQ, int dsp_type,
                                  int dsp_method, double *atfinal, double *afinal, double *a2final) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, dsp_method);
    if (rc) return rc;
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_all_vectors_scan: frames are not staged (stage_frames first)");
    if (NM == 0 || NQ == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_all_vectors_scan: No qvectors left to compute");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    const size_t plen = partial_len(ctx, dsp_type);
    rc = ensure<double>(ctx, &ctx->d_partial, &ctx->partial_cap, NQ * plen);
    if (rc) return rc;
    rc = sgpu_compute_all_vectors_scan_partial(ctx, v, NM, s, NQ, dsp_type, ctx->d_partial);
    if (rc) return rc;
    for (size_t n = 0; n < NQ; n++) {
        rc = sgpu_finalize(ctx, ctx->d_partial + n * plen, dsp_type, dsp_method, 1.0 / (double)NM, atfinal + n * 2 * ctx->NFt,
                           afinal + 2 * n, a2final + 2 * n);
        if (rc) return rc;
    }
    return SGPU_OK;
}

/* how the last scan call was evaluated: passes of the plain scan kernel, of the corrected scan kernel, and |q| values
 * that went through the general kernel one at a time */
int sgpu_last_scan_plan(const sgpu_ctx *ctx, int *plain, int *corrected, int *single) {
    if (!ctx) return SGPU_EINVAL;
    if (plain) *plain = ctx->scan_kinds[0];
    if (corrected) *corrected = ctx->scan_kinds[1];
    if (single) *single = ctx->scan_kinds[2];
    return SGPU_OK;
}

int sgpu_compute_self_vectors_partial(sgpu_ctx *ctx, const double *qvecs, size_t NM, int dsp_type, double *d_partial) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    if (NM == 0) return zero_partial(ctx, dsp_type, d_partial);  // a rank without atoms contributes zeros
    if (!qvecs) return fail(ctx, SGPU_EINVAL, "sgpu_compute_self_vectors: qvecs is NULL");
    if (ctx->mode != 2) return fail(ctx, SGPU_ESTATE, "sgpu_compute_self_vectors: atoms are not staged (stage_atoms first)");
    if (ctx->nb != ctx->NA) return fail(ctx, SGPU_ESTATE, "sgpu_compute_self_vectors: scattering factors not set for the staged atoms");
    if (NM == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_self_vectors: No qvectors left to compute");
    if (!d_partial) return fail(ctx, SGPU_EINVAL, "sgpu_compute_self_vectors: d_partial is NULL");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    rc = upload_q(ctx, qvecs, NM, 1);
    if (rc) return rc;

    if (dsp_type == SGPU_DSP_AUTOCORRELATE) {
        // fused path: timelines are generated, transformed and reduced inside the SM (selffused.cu)
        const SelfPlan &sp = ctx->splan;
        // freshly staged atoms still arriving in chunks (sgpu_stage_atoms): every chunk is brought into the layout the path
        // wants and evaluated as soon as it has landed, the copy of the next ones runs underneath
        const bool chunked = !ctx->chunks.empty() && ctx->own_xyz && ctx->atoms_dec_R == 0;
        if (!chunked) {
            rc = set_atoms_layout(ctx, sp.split ? sp.R : 0);  // the split path reads sub-sequences of frames: keep them contiguous
            if (rc) return rc;
        }
        size_t atoms_per_batch = std::max<size_t>(1, ((size_t)256 << 20) / (NM * (size_t)sp.R * sizeof(double2)));
        atoms_per_batch = std::min(atoms_per_batch, ctx->NA);
        rc = ensure_work(ctx, std::max(self_work_bytes(&sp, atoms_per_batch * NM), corr_work_bytes(&ctx->plan, 1)));
        if (rc) return rc;
        CK(cudaMemsetAsync(d_partial, 0, partial_len(ctx, dsp_type) * sizeof(double), ctx->stream));
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        std::vector<sgpu_ctx::Chunk> spans;
        if (chunked) spans = ctx->chunks;
        else spans.push_back(sgpu_ctx::Chunk{0, ctx->NA, nullptr});
        const int dec = chunked ? (sp.split ? 1 : 0) : (ctx->atoms_dec_R == sp.R && sp.split ? 1 : 0);
        for (auto &c : spans) {
            if (c.ready) {
                CK(cudaStreamWaitEvent(ctx->stream, c.ready, 0));
                if (sp.split) {
                    rc = decimate_rows(ctx, c.f0, c.nf, sp.R);
                    if (rc) return rc;
                }
            }
            for (size_t n0 = c.f0; n0 < c.f0 + c.nf; n0 += atoms_per_batch) {
                const size_t nn = std::min(atoms_per_batch, c.f0 + c.nf - n0);
                ctx->launches += self_power_accumulate(&sp, ctx->d_xyz, ctx->d_b, ctx->d_qs, NM, n0, nn, ctx->d_work, d_partial,
                                                       d_partial + sp.L, dec, ctx->stream);
                CK(cudaGetLastError());
            }
        }
        if (chunked) {
            drop_chunks(ctx);
            ctx->atoms_dec_R = sp.split ? sp.R : 0;
        }
    } else {
    rc = set_atoms_layout(ctx, 0);
    if (rc) return rc;
    // batch atoms so that amplitudes + scratch fit a memory budget
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = std::min<size_t>((free_b + ctx->work_cap + ctx->A_cap * sizeof(double2)) / 3, (size_t)24 << 30);
    const size_t per_tl = ctx->NF * sizeof(double2) + 64;
    size_t nt_max = std::max<size_t>(budget / per_tl, NM);
    size_t atoms_per_batch = std::max<size_t>(1, nt_max / NM);
    atoms_per_batch = std::min(atoms_per_batch, ctx->NA);
    const size_t nt_batch = atoms_per_batch * NM;
    rc = ensure<double2>(ctx, &ctx->d_A, &ctx->A_cap, nt_batch * ctx->NF);
    if (rc) return rc;
    rc = ensure_work(ctx, dsp_work_bytes(ctx, nt_batch, dsp_type));
    if (rc) return rc;
    CK(cudaMemsetAsync(d_partial, 0, partial_len(ctx, dsp_type) * sizeof(double), ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (size_t n0 = 0; n0 < ctx->NA; n0 += atoms_per_batch) {
        const size_t nn = std::min(atoms_per_batch, ctx->NA - n0);
        ctx->launches += launch_amplitude_self(ctx->d_xyz, ctx->d_b, ctx->d_qs, ctx->d_A, ctx->NF, ctx->NF, NM, n0, nn,
                                               ctx->stream);
        CK(cudaGetLastError());
        rc = dsp_accumulate(ctx, nn * NM, dsp_type, d_partial);
        if (rc) return rc;
    }
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    ctx->A_NM = 0;
    return SGPU_OK;
}

// moments -> device (l, m) ints; returns lmax through *lmax_out.  Mirrors the reference's validity check
// (multipole_scatter_device.cpp:459-465, parameters.cpp:1101-1113).
static int upload_moments(sgpu_ctx *ctx, const long *lm, size_t NM, int *lmax_out) {
    int lmax = 0;
    std::vector<int> h_lm(NM * 2);
    for (size_t i = 0; i < NM; i++) {
        const long l = lm[2 * i], m = lm[2 * i + 1];
        if (l < 0 || std::labs(m) > l)
            return fail(ctx, SGPU_EINVAL,
                        "Combination of Major and minor moment not allowed: l=" + std::to_string(l) + ", m" +
                            std::to_string(m));
        if (l > 50) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpsphere: major moment > 50 not supported");
        lmax = std::max<int>(lmax, (int)l);
        h_lm[2 * i] = (int)l;
        h_lm[2 * i + 1] = (int)m;
    }
    int rc = ensure<int>(ctx, &ctx->d_lm, &ctx->lm_cap, NM * 2);
    if (rc) return rc;
    rc = small_upload(ctx, ctx->d_lm, h_lm.data(), NM * 2 * sizeof(int));
    if (rc) return rc;
    *lmax_out = lmax;
    return SGPU_OK;
}

int sgpu_set_factors_batch(sgpu_ctx *ctx, const double *b, size_t NQ, size_t n) {
    if (!ctx) return SGPU_EINVAL;
    if (!b || n == 0 || NQ == 0) return fail(ctx, SGPU_EINVAL, "sgpu_set_factors_batch: empty factors");
    CK(cudaSetDevice(ctx->device));
    int rc = ensure<double>(ctx, &ctx->d_bq, &ctx->bq_cap, NQ * n);
    if (rc) return rc;
    rc = small_upload(ctx, ctx->d_bq, b, NQ * n * sizeof(double));
    if (rc) return rc;
    ctx->bq_nq = NQ;
    ctx->bq_n = n;
    ctx->bq_uniform = true;
    for (size_t q = 1; q < NQ && ctx->bq_uniform; q++)
        if (memcmp(b, b + q * n, n * sizeof(double)) != 0) ctx->bq_uniform = false;
    return SGPU_OK;
}

// amplitudes A[q][mom][f] of NQ |q| values over the atoms [atom_first, atom_first+atom_count) into d_amp
int sgpu_mpsphere_amplitudes(sgpu_ctx *ctx, const double *qlens, size_t NQ, const long *lm, size_t NM, size_t atom_first,
                             size_t atom_count, double *d_amp) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (!qlens || !lm || !d_amp) return fail(ctx, SGPU_EINVAL, "sgpu_mpsphere_amplitudes: NULL argument");
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpsphere: frames are not staged (stage_frames first)");
    if (ctx->NFt != ctx->NF) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpsphere: frame windows are not supported on this path");
    if (ctx->repr != SGPU_REPR_SPHERICAL)
        return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpsphere: staged frames are not in spherical representation");
    if (NM == 0 || NQ == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpsphere: No moments / qvectors to compute");
    if (atom_first + atom_count > ctx->NA) return fail(ctx, SGPU_EINVAL, "sgpu_mpsphere_amplitudes: atom range out of bounds");
    // factors: a per-|q| batch if one was set for exactly this NQ, else the single set
    const double *d_b = ctx->d_b;
    size_t b_stride = 0;
    if (ctx->bq_nq == NQ && ctx->bq_n == ctx->NA && ctx->d_bq) {
        d_b = ctx->d_bq;
        b_stride = ctx->NA;
    } else if (ctx->nb != ctx->NA) {
        return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpsphere: scattering factors not set for the staged atoms");
    }
    int lmax = 0;
    int rc = upload_moments(ctx, lm, NM, &lmax);
    if (rc) return rc;
    rc = ensure<double>(ctx, &ctx->d_qlens, &ctx->qlens_cap, NQ);
    if (rc) return rc;
    rc = small_upload(ctx, ctx->d_qlens, qlens, NQ * sizeof(double));
    if (rc) return rc;
    double2 *A = reinterpret_cast<double2 *>(d_amp);
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (lmax <= 21) {
        // the atom splits (and with them the scratch) depend on the frames of a launch: size for the whole trajectory and
        // for every staging chunk that may be launched on its own below
        size_t work = multipole_batch_work_doubles(ctx->NF, lmax, std::max<size_t>(atom_count, 1), (int)NQ);
        for (auto &c : ctx->chunks)
            work = std::max(work, multipole_batch_work_doubles(c.nf, lmax, std::max<size_t>(atom_count, 1), (int)NQ));
        rc = ensure_work(ctx, work * sizeof(double));
        if (rc) return rc;
        // one launch over all frames if every staging chunk has landed (and been converted), else one launch per chunk as
        // the chunks arrive: staging overlaps the kernel (as the coherent amplitude kernels do)
        bool all_ready = true;
        for (auto &c : ctx->chunks)
            if (cudaEventQuery(c.ready) != cudaSuccess) all_ready = false;
        cudaGetLastError();
        if (all_ready) {
            drop_chunks(ctx);
            ctx->launches += launch_multipole_sphere_batch(ctx->d_xyz, d_b, b_stride, ctx->d_qlens, (int)NQ, ctx->d_lm, NM, lmax, A,
                                                           ctx->NF, ctx->NA, atom_first, atom_first + atom_count, 0, ctx->NF,
                                                           reinterpret_cast<double *>(ctx->d_work), ctx->stream);
        } else {
            for (auto &c : ctx->chunks) {
                CK(cudaStreamWaitEvent(ctx->stream, c.ready, 0));
                ctx->launches += launch_multipole_sphere_batch(ctx->d_xyz, d_b, b_stride, ctx->d_qlens, (int)NQ, ctx->d_lm, NM, lmax,
                                                               A, ctx->NF, ctx->NA, atom_first, atom_first + atom_count, c.f0, c.nf,
                                                               reinterpret_cast<double *>(ctx->d_work), ctx->stream);
            }
        }
    } else {
        CK(cudaStreamSynchronize(ctx->copy_stream));
        if (ctx->conv_stream) CK(cudaStreamSynchronize(ctx->conv_stream));
        drop_chunks(ctx);
        // more (l,m) pairs than threads of the batched kernel: one pass per |q| with the shuffle-reduction kernel
        if (atom_first != 0 || atom_count != ctx->NA)
            return fail(ctx, SGPU_EINVAL, "sgpu_mpsphere_amplitudes: atom sharding needs moments with l <= 21");
        int nsplit = 1;
        rc = ensure_work(ctx, multipole_work_doubles(ctx->NF, lmax, &nsplit, ctx->NA) * sizeof(double));
        if (rc) return rc;
        for (size_t q = 0; q < NQ; q++)
            ctx->launches += launch_multipole_sphere(ctx->d_xyz, d_b + q * b_stride, qlens[q], ctx->d_lm, NM, lmax,
                                                     A + q * NM * ctx->NF, ctx->NF, ctx->NA, 0, ctx->NF,
                                                     reinterpret_cast<double *>(ctx->d_work), ctx->stream);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));  // sgpu_last_amplitude_ms: the amplitude kernels of this call
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    return SGPU_OK;
}

// DSP of the (summed) amplitudes of NQ |q| values: d_partials receives NQ consecutive packed partials
int sgpu_mpsphere_dsp_partial(sgpu_ctx *ctx, const double *d_amp, size_t NQ, size_t NM, int dsp_type, double *d_partials) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    if (!d_amp || !d_partials) return fail(ctx, SGPU_EINVAL, "sgpu_mpsphere_dsp_partial: NULL argument");
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_mpsphere_dsp_partial: frames are not staged");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    rc = ensure_work(ctx, dsp_work_bytes(ctx, NM, dsp_type));
    if (rc) return rc;
    const size_t plen = partial_len(ctx, dsp_type);
    CK(cudaMemsetAsync(d_partials, 0, NQ * plen * sizeof(double), ctx->stream));
    const double2 *A = reinterpret_cast<const double2 *>(d_amp);
    for (size_t q = 0; q < NQ; q++) {
        rc = dsp_accumulate(ctx, NM, dsp_type, d_partials + q * plen, A + q * NM * ctx->NF);
        if (rc) return rc;
    }
    return SGPU_OK;
}

int sgpu_compute_mpsphere_batch_partial(sgpu_ctx *ctx, const double *qlens, size_t NQ, const long *lm, size_t NM,
                                        int dsp_type, double *d_partials) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpsphere: frames are not staged (stage_frames first)");
    if (NM == 0 || NQ == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpsphere: No moments / qvectors to compute");
    rc = ensure<double2>(ctx, &ctx->d_A, &ctx->A_cap, NQ * NM * ctx->NF);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    rc = sgpu_mpsphere_amplitudes(ctx, qlens, NQ, lm, NM, 0, ctx->NA, reinterpret_cast<double *>(ctx->d_A));
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->A_NM = (NQ == 1) ? NM : 0;
    rc = sgpu_mpsphere_dsp_partial(ctx, reinterpret_cast<const double *>(ctx->d_A), NQ, NM, dsp_type, d_partials);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    return SGPU_OK;
}

int sgpu_compute_mpsphere_partial(sgpu_ctx *ctx, double qlen, const long *lm, size_t NM, int dsp_type,
                                  double *d_partial) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    if (NM == 0) return zero_partial(ctx, dsp_type, d_partial);  // a rank without moments contributes zeros
    if (!lm) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpsphere: lm is NULL");
    // a per-|q| factor batch does not apply to the single-|q| entry point
    ctx->bq_nq = 0;
    return sgpu_compute_mpsphere_batch_partial(ctx, &qlen, 1, lm, NM, dsp_type, d_partial);
}

int sgpu_compute_mpsphere_batch(sgpu_ctx *ctx, const double *qlens, size_t NQ, const long *lm, size_t NM, int dsp_type,
                                int dsp_method, double *atfinal, double *afinal, double *a2final) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, dsp_method);
    if (rc) return rc;
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpsphere: frames are not staged (stage_frames first)");
    if (NM == 0 || NQ == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpsphere: No moments / qvectors to compute");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    const size_t plen = partial_len(ctx, dsp_type);
    rc = ensure<double>(ctx, &ctx->d_partial, &ctx->partial_cap, NQ * plen);
    if (rc) return rc;
    rc = sgpu_compute_mpsphere_batch_partial(ctx, qlens, NQ, lm, NM, dsp_type, ctx->d_partial);
    if (rc) return rc;
    const double scale = 1.0 / (4.0 * 3.14159265358979323846);
    for (size_t q = 0; q < NQ; q++) {
        rc = sgpu_finalize(ctx, ctx->d_partial + q * plen, dsp_type, dsp_method, scale, atfinal + q * 2 * ctx->NF,
                           afinal + 2 * q, a2final + 2 * q);
        if (rc) return rc;
    }
    return SGPU_OK;
}

int sgpu_finalize(sgpu_ctx *ctx, const double *d_partial, int dsp_type, int dsp_method, double scale, double *atfinal,
                  double afinal[2], double a2final[2]) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, dsp_method);
    if (rc) return rc;
    if (ctx->mode == 0) return fail(ctx, SGPU_ESTATE, "sgpu_finalize: nothing staged");
    if (!d_partial || !atfinal || !afinal || !a2final) return fail(ctx, SGPU_EINVAL, "sgpu_finalize: NULL argument");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    rc = ensure<double2>(ctx, &ctx->d_out, &ctx->out_cap, ctx->NFt);
    if (rc) return rc;
    rc = ensure_work(ctx, corr_work_bytes(&ctx->plan, 1));
    if (rc) return rc;
    const bool conj = (dsp_type == SGPU_DSP_AUTOCORRELATE && dsp_method == SGPU_METHOD_DIRECT);
    const double *acc;
    if (uses_self_plan(ctx, dsp_type)) {
        rc = ensure_work(ctx, self_work_bytes(&ctx->splan, 1));
        if (rc) return rc;
        ctx->launches += self_finalize(&ctx->splan, d_partial, ctx->d_work, ctx->d_out, scale, conj ? 1 : 0, ctx->stream);
        acc = d_partial + ctx->splan.L;
    } else if (dsp_type == SGPU_DSP_AUTOCORRELATE) {
        ctx->launches += corr_finalize(&ctx->plan, d_partial, ctx->d_work, ctx->d_out, scale, conj ? 1 : 0, ctx->stream);
        acc = d_partial + ctx->plan.L;
    } else {
        ctx->launches += launch_scale_complex(reinterpret_cast<const double2 *>(d_partial), ctx->d_out, ctx->NFt, scale,
                                              ctx->stream);
        acc = d_partial + 2 * ctx->NFt;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_acc, acc, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(atfinal, ctx->d_out, ctx->NFt * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    afinal[0] = ctx->h_acc[0] * scale;
    afinal[1] = (conj ? -ctx->h_acc[1] : ctx->h_acc[1]) * scale;
    a2final[0] = ctx->h_acc[2] * scale;
    a2final[1] = 0.0;
    return SGPU_OK;
}

int sgpu_compute_all_vectors(sgpu_ctx *ctx, const double *qvecs, size_t NM, int dsp_type, int dsp_method,
                             double *atfinal, double afinal[2], double a2final[2]) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, dsp_method);
    if (rc) return rc;
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_all_vectors: frames are not staged (stage_frames first)");
    if (NM == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_all_vectors: No qvectors left to compute");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    rc = ensure_internal_partial(ctx, dsp_type);
    if (rc) return rc;
    rc = sgpu_compute_all_vectors_partial(ctx, qvecs, NM, dsp_type, ctx->d_partial);
    if (rc) return rc;
    return sgpu_finalize(ctx, ctx->d_partial, dsp_type, dsp_method, 1.0 / (double)NM, atfinal, afinal, a2final);
}

int sgpu_compute_self_vectors(sgpu_ctx *ctx, const double *qvecs, size_t NM, int dsp_type, int dsp_method,
                              double *atfinal, double afinal[2], double a2final[2]) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, dsp_method);
    if (rc) return rc;
    if (ctx->mode != 2) return fail(ctx, SGPU_ESTATE, "sgpu_compute_self_vectors: atoms are not staged (stage_atoms first)");
    if (NM == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_self_vectors: No qvectors left to compute");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    rc = ensure_internal_partial(ctx, dsp_type);
    if (rc) return rc;
    rc = sgpu_compute_self_vectors_partial(ctx, qvecs, NM, dsp_type, ctx->d_partial);
    if (rc) return rc;
    return sgpu_finalize(ctx, ctx->d_partial, dsp_type, dsp_method, 1.0 / (double)NM, atfinal, afinal, a2final);
}

int sgpu_compute_mpsphere(sgpu_ctx *ctx, double qlen, const long *lm, size_t NM, int dsp_type, int dsp_method,
                          double *atfinal, double afinal[2], double a2final[2]) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, dsp_method);
    if (rc) return rc;
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpsphere: frames are not staged (stage_frames first)");
    if (NM == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpsphere: No moments to compute");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    rc = ensure_internal_partial(ctx, dsp_type);
    if (rc) return rc;
    rc = sgpu_compute_mpsphere_partial(ctx, qlen, lm, NM, dsp_type, ctx->d_partial);
    if (rc) return rc;
    const double scale = 1.0 / (4.0 * 3.14159265358979323846);  // multipole_scatter_device.cpp:395
    return sgpu_finalize(ctx, ctx->d_partial, dsp_type, dsp_method, scale, atfinal, afinal, a2final);
}

/* ---- multipole cylinder ------------------------------------------------------------------------ */

int sgpu_mpcylinder_amplitudes(sgpu_ctx *ctx, const double q[3], const double axis[3], const long *lm, size_t NM,
                               size_t atom_first, size_t atom_count, double *d_amp) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (!q || !axis || !lm || !d_amp) return fail(ctx, SGPU_EINVAL, "sgpu_mpcylinder_amplitudes: NULL argument");
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpcylinder: frames are not staged (stage_frames first)");
    if (ctx->NFt != ctx->NF) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpcylinder: frame windows are not supported on this path");
    if (ctx->repr != SGPU_REPR_CYLINDRICAL)
        return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpcylinder: staged frames are not in cylindrical representation");
    if (axis[0] != ctx->cyl_axis[0] || axis[1] != ctx->cyl_axis[1] || axis[2] != ctx->cyl_axis[2])
        return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpcylinder: the frames were converted with a different axis");
    if (NM == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpcylinder: No moments to compute");
    if (atom_first + atom_count > ctx->NA) return fail(ctx, SGPU_EINVAL, "sgpu_mpcylinder_amplitudes: atom range out of bounds");
    if (ctx->nb != ctx->NA) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpcylinder: scattering factors not set for the staged atoms");
    // moments: parameters.cpp:1082-1102
    int nmax = 0;
    std::vector<int> h_lm(NM * 2);
    for (size_t i = 0; i < NM; i++) {
        const long l = lm[2 * i], m = lm[2 * i + 1];
        if (l < 0) return fail(ctx, SGPU_EINVAL, "Major multipole moment must be >= 0!");
        if (m < 0 || m > 3) return fail(ctx, SGPU_EINVAL, "Minor multipole moment must be between 0 and 3!");
        if (l == 0 && m != 0) return fail(ctx, SGPU_EINVAL, "Minor multipole moment must be 0 for Major 0!");
        const long order = (l == 0) ? 0 : ((m < 2) ? 2 * l : 2 * l - 1);
        if (order > mpcylinder_max_order()) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpcylinder: Bessel order > 199 not supported");
        nmax = std::max<int>(nmax, (int)order);
        h_lm[2 * i] = (int)l;
        h_lm[2 * i + 1] = (int)m;
    }
    int rc = ensure<int>(ctx, &ctx->d_lm, &ctx->lm_cap, NM * 2);
    if (rc) return rc;
    rc = small_upload(ctx, ctx->d_lm, h_lm.data(), NM * 2 * sizeof(int));
    if (rc) return rc;
    // q in the cylinder basis: CylinderCoor3D(base.project(q)) (multipole_scatter_device.cpp:925-929, coor3d.cpp:113-138)
    double base[9];
    if (!vector_base(axis, base)) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpcylinder: axis has zero length");
    const double px = q[0] * base[0] + q[1] * base[1] + q[2] * base[2];
    const double py = q[0] * base[3] + q[1] * base[4] + q[2] * base[5];
    const double pz = q[0] * base[6] + q[1] * base[7] + q[2] * base[8];
    const double qr = std::sqrt(px * px + py * py);
    const double PIf = (double)3.14159274101257324f, PI2f = (double)1.57079637050628662f;  // sign() returns float there
    double qphi = 0.0;
    if (px != 0.0) {
        qphi = std::atan(py / px);
        if (px < 0.0) qphi = ((py < 0.0) ? -PIf : PIf) + qphi;
    } else if (py != 0.0) {
        qphi = (py < 0.0) ? -PI2f : PI2f;
    }
    if (qphi < 0) qphi = 2 * 3.14159265358979323846 + qphi;
    CK(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->conv_stream) CK(cudaStreamSynchronize(ctx->conv_stream));
    drop_chunks(ctx);
    rc = ensure_work(ctx, mpcylinder_work_doubles(ctx->NF, nmax, std::max<size_t>(atom_count, 1)) * sizeof(double));
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    ctx->launches += launch_mpcylinder(ctx->d_xyz, ctx->d_b, qr, qphi, pz, ctx->d_lm, NM, nmax, reinterpret_cast<double2 *>(d_amp),
                                       ctx->NF, ctx->NA, atom_first, atom_first + atom_count,
                                       reinterpret_cast<double *>(ctx->d_work), ctx->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));  // sgpu_last_amplitude_ms: the amplitude kernels of this call
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    return SGPU_OK;
}

int sgpu_compute_mpcylinder_partial(sgpu_ctx *ctx, const double q[3], const double axis[3], const long *lm, size_t NM,
                                    int dsp_type, double *d_partial) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, SGPU_METHOD_FFTW);
    if (rc) return rc;
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpcylinder: frames are not staged (stage_frames first)");
    if (NM == 0) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpcylinder: No moments to compute");
    if (!d_partial) return fail(ctx, SGPU_EINVAL, "sgpu_compute_mpcylinder: d_partial is NULL");
    rc = ensure<double2>(ctx, &ctx->d_A, &ctx->A_cap, NM * ctx->NF);
    if (rc) return rc;
    rc = sgpu_mpcylinder_amplitudes(ctx, q, axis, lm, NM, 0, ctx->NA, reinterpret_cast<double *>(ctx->d_A));  // records ev0, ev1
    if (rc) return rc;
    ctx->A_NM = NM;
    rc = sgpu_mpsphere_dsp_partial(ctx, reinterpret_cast<const double *>(ctx->d_A), 1, NM, dsp_type, d_partial);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->have_times = true;
    ctx->dsp_split = false;
    return SGPU_OK;
}

int sgpu_compute_mpcylinder(sgpu_ctx *ctx, const double q[3], const double axis[3], const long *lm, size_t NM, int dsp_type,
                            int dsp_method, double *atfinal, double afinal[2], double a2final[2]) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = check_dsp(ctx, dsp_type, dsp_method);
    if (rc) return rc;
    if (ctx->mode != 1) return fail(ctx, SGPU_ESTATE, "sgpu_compute_mpcylinder: frames are not staged (stage_frames first)");
    rc = ensure_plan(ctx);
    if (rc) return rc;
    rc = ensure_internal_partial(ctx, dsp_type);
    if (rc) return rc;
    rc = sgpu_compute_mpcylinder_partial(ctx, q, axis, lm, NM, dsp_type, ctx->d_partial);
    if (rc) return rc;
    const double scale = 1.0 / (2.0 * 3.14159265358979323846);  // multipole_scatter_device.cpp:866
    return sgpu_finalize(ctx, ctx->d_partial, dsp_type, dsp_method, scale, atfinal, afinal, a2final);
}

/* ---- introspection ---------------------------------------------------------------------------- */

int sgpu_get_amplitudes(sgpu_ctx *ctx, double *A, size_t NM, size_t NF) {
    if (!ctx || !A) return SGPU_EINVAL;
    if (ctx->A_NM == 0 || NM != ctx->A_NM || NF != ctx->NFt)
        return fail(ctx, SGPU_ESTATE, "sgpu_get_amplitudes: no matching amplitudes from the last compute");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(A, ctx->d_A, NM * NF * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SGPU_OK;
}

int sgpu_last_amplitude_ms(sgpu_ctx *ctx, float *ms) {
    if (!ctx || !ms) return SGPU_EINVAL;
    if (!ctx->have_times) return fail(ctx, SGPU_ESTATE, "no compute has run yet");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev2));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return SGPU_OK;
}

int sgpu_last_dsp_ms(sgpu_ctx *ctx, float *ms) {
    if (!ctx || !ms) return SGPU_EINVAL;
    if (!ctx->have_times) return fail(ctx, SGPU_ESTATE, "no compute has run yet");
    CK(cudaSetDevice(ctx->device));
    if (ctx->dsp_split) {
        CK(cudaEventSynchronize(ctx->evd1));
        CK(cudaEventElapsedTime(ms, ctx->evd0, ctx->evd1));
        return SGPU_OK;
    }
    CK(cudaEventSynchronize(ctx->ev2));
    CK(cudaEventElapsedTime(ms, ctx->ev1, ctx->ev2));
    return SGPU_OK;
}

int sgpu_timer_start(sgpu_ctx *ctx) {
    if (!ctx) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->tm0) CK(cudaEventCreate(&ctx->tm0));
    if (!ctx->tm1) CK(cudaEventCreate(&ctx->tm1));
    CK(cudaEventRecord(ctx->tm0, ctx->stream));
    return SGPU_OK;
}

int sgpu_timer_stop(sgpu_ctx *ctx, float *ms) {
    if (!ctx || !ms) return SGPU_EINVAL;
    if (!ctx->tm0 || !ctx->tm1) return fail(ctx, SGPU_ESTATE, "sgpu_timer_stop: timer not started");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->tm1, ctx->stream));
    CK(cudaEventSynchronize(ctx->tm1));
    CK(cudaEventElapsedTime(ms, ctx->tm0, ctx->tm1));
    return SGPU_OK;
}

int sgpu_measure_fp64_peak(sgpu_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return SGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    double *sink = nullptr;
    CK(cudaMalloc(reinterpret_cast<void **>(&sink), 64));
    const int blocks = prop.multiProcessorCount * 8;
    double best = 0.0;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(a, ctx->stream));
        const double flops = launch_fp64_peak(sink, 20000, blocks, ctx->stream);
        ctx->launches++;
        CK(cudaEventRecord(b, ctx->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    *tflops = best;
    return SGPU_OK;
}

int sgpu_synth_trajectory(sgpu_ctx *ctx, float *d_xyz, size_t NF, size_t NA, size_t atom0, size_t atom_stride,
                          size_t NA_out, float box, float offset, float step_scale, uint64_t seed, int layout) {
    if (!ctx || !d_xyz) return SGPU_EINVAL;
    if (layout != 0 && layout != 1) return fail(ctx, SGPU_EINVAL, "sgpu_synth_trajectory: layout must be 0 or 1");
    CK(cudaSetDevice(ctx->device));
    ctx->launches += launch_synth_trajectory(d_xyz, NF, NA, atom0, atom_stride, NA_out, box, offset, step_scale, seed,
                                             layout, ctx->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return SGPU_OK;
}

}  // extern "C"
