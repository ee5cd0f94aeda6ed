// nccl_comm.cpp — the host layer's communicator on NCCL, native code only (include/sassena_host.h, sass_comm_nccl_*).
// Stands in for the boost::mpi::communicator the reference's devices take (abstract_scatter_device.hpp:173-181): rank / size,
// the sum over the partition (all_vectors_scatter_device.cpp:335-343), barrier, and split (scatter_device_factory.cpp:116-120).
#include <cuda_runtime.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/sassena_host.h"
#include "../nccl_api.hpp"

namespace {

thread_local std::string g_err;

struct NcclComm {
    ncclComm_t comm = nullptr;
    int size = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;
    double *d_one = nullptr;  // barrier scratch
};

const sass::NcclApi *api() {
    std::string e;
    const sass::NcclApi *a = sass::nccl_api(&e);
    if (!a) g_err = e;
    return a;
}

size_t cb_rank(void *u) { return (size_t) static_cast<NcclComm *>(u)->rank; }
size_t cb_size(void *u) { return (size_t) static_cast<NcclComm *>(u)->size; }

int cb_allreduce(void *u, double *d_buf, size_t n) {
    NcclComm *c = static_cast<NcclComm *>(u);
    const sass::NcclApi *a = api();
    if (!a || !c->comm) return 1;
    if (cudaSetDevice(c->device) != cudaSuccess) return 1;
    if (n == 0) return 0;
    // the caller has synchronised the stream that produced d_buf (sass_comm_vtbl contract); the result is complete on return
    if (a->AllReduce(d_buf, d_buf, n, ncclFloat64, ncclSum, c->comm, c->stream) != ncclSuccess) return 1;
    return cudaStreamSynchronize(c->stream) == cudaSuccess ? 0 : 1;
}

int cb_barrier(void *u) {
    NcclComm *c = static_cast<NcclComm *>(u);
    if (c->size == 1) return 0;
    if (cudaSetDevice(c->device) != cudaSuccess) return 1;
    if (!c->d_one && cudaMalloc(reinterpret_cast<void **>(&c->d_one), sizeof(double)) != cudaSuccess) return 1;
    if (cudaMemsetAsync(c->d_one, 0, sizeof(double), c->stream) != cudaSuccess) return 1;
    return cb_allreduce(u, c->d_one, 1);
}

void cb_release(void *u) {
    NcclComm *c = static_cast<NcclComm *>(u);
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) {
        if (const sass::NcclApi *a = api()) a->CommDestroy(c->comm);
    }
    if (c->d_one) cudaFree(c->d_one);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void *cb_split(void *u, int color) {
    NcclComm *c = static_cast<NcclComm *>(u);
    const sass::NcclApi *a = api();
    if (!a) return nullptr;
    if (cudaSetDevice(c->device) != cudaSuccess) return nullptr;
    NcclComm *n = new NcclComm();
    n->device = c->device;
    if (cudaStreamCreateWithFlags(&n->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete n;
        return nullptr;
    }
    // boost::mpi::communicator::split(color): ranks of equal colour, ordered by their rank in the parent (key = rank)
    if (a->CommSplit(c->comm, color, c->rank, &n->comm, nullptr) != ncclSuccess || !n->comm) {
        cb_release(n);
        return nullptr;
    }
    // rank / size in the child: colours are known only locally, so count through the child itself
    double *d = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&d), (size_t)c->size * sizeof(double)) != cudaSuccess) {
        cb_release(n);
        return nullptr;
    }
    // child ranks keep the parent order: my child rank = members with a smaller parent rank.  Every member marks its parent
    // rank in a vector of the parent's size and the child sums it.
    std::vector<double> h((size_t)c->size, 0.0);
    h[(size_t)c->rank] = 1.0;
    bool ok = cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, n->stream) == cudaSuccess &&
              a->AllReduce(d, d, h.size(), ncclFloat64, ncclSum, n->comm, n->stream) == ncclSuccess &&
              cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, n->stream) == cudaSuccess &&
              cudaStreamSynchronize(n->stream) == cudaSuccess;
    cudaFree(d);
    if (!ok) {
        cb_release(n);
        return nullptr;
    }
    n->size = 0;
    n->rank = 0;
    for (int r = 0; r < c->size; r++) {
        if (h[(size_t)r] > 0.5) {
            if (r < c->rank) n->rank++;
            n->size++;
        }
    }
    return n;
}

void *cb_nccl(void *u) { return static_cast<NcclComm *>(u)->comm; }

}  // namespace

extern "C" {

int sass_comm_nccl_unique_id(char id128[128]) {
    const sass::NcclApi *a = api();
    if (!a || !id128) return 1;
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != ncclSuccess) return 1;
    memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return 0;
}

int sass_comm_nccl_bootstrap_file(const char *path, int nranks, int rank, double timeout_s, char id128[128]) {
    if (!path || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return 1;
    if (rank == 0) {
        if (sass_comm_nccl_unique_id(id128)) return 1;
        const std::string tmp = std::string(path) + ".tmp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f) return 1;
        const bool ok = fwrite(id128, 1, 128, f) == 128;
        fclose(f);
        return (ok && rename(tmp.c_str(), path) == 0) ? 0 : 1;  // rename: readers never see a partial id
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        if (FILE *f = fopen(path, "rb")) {
            const size_t n = fread(id128, 1, 128, f);
            fclose(f);
            if (n == 128) return 0;
        }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) return 2;
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
}

int sass_comm_nccl_create(const char id128[128], int nranks, int rank, int device, sass_comm_vtbl *out) {
    const sass::NcclApi *a = api();
    if (!a || !id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return 1;
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    NcclComm *c = new NcclComm();
    c->device = device;
    c->size = nranks;
    c->rank = rank;
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        a->CommInitRank(&c->comm, nranks, id, rank) != ncclSuccess) {
        cb_release(c);
        return 1;
    }
    out->user = c;
    out->rank = cb_rank;
    out->size = cb_size;
    out->allreduce_sum = cb_allreduce;
    out->barrier = cb_barrier;
    out->split = cb_split;
    out->release = cb_release;
    out->nccl_comm = cb_nccl;
    return 0;
}

}  // extern "C"
