// fp64_micro.cu — B200 FP64 pipe microbenchmarks used to size the amplitude kernel:
// DFMA/DADD/DMUL issue rate with register vs constant operands, dependent-issue latency, and co-issue of
// integer-pipe work alongside a saturated FP64 pipe.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_latency(double *out, long long *cyc, double b, double c) {
    double a = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) a = fma(a, b, c);
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// N independent chains, operands all registers and all distinct: a_i = fma(x_i, y_i, a_i)
template <int N>
__global__ void k_rrr(double *out, long long *cyc, const double *in) {
    double a[N], x[N], y[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        a[j] = in[j] + threadIdx.x;
        x[j] = in[N + j];
        y[j] = in[2 * N + j];
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[j]) : "d"(x[j]), "d"(y[j]));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < N; j++) s += a[j];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// N chains, one operand shared (reuse cache friendly): a_i = fma(a_i, m, c)
template <int N>
__global__ void k_shared_ops(double *out, long long *cyc, const double *in) {
    double a[N];
    double m = in[0], c = in[1];
#pragma unroll
    for (int j = 0; j < N; j++) a[j] = in[j] + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j]) : "d"(m), "d"(c));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < N; j++) s += a[j];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// DADD / DMUL throughput
template <int N, int OP>
__global__ void k_addmul(double *out, long long *cyc, const double *in) {
    double a[N], x[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        a[j] = in[j] + threadIdx.x;
        x[j] = in[N + j];
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            if (OP == 0) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[j]) : "d"(x[j]));
            else asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[j]) : "d"(x[j]));
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < N; j++) s += a[j];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FP64 + integer co-issue: per DFMA, K integer ops (lop3/sel) on independent registers
template <int N, int K>
__global__ void k_mixed(double *out, long long *cyc, const double *in) {
    double a[N], x[N], y[N];
    unsigned u[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        a[j] = in[j] + threadIdx.x;
        x[j] = in[N + j];
        y[j] = in[2 * N + j];
        u[j] = threadIdx.x + j;
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[j]) : "d"(x[j]), "d"(y[j]));
#pragma unroll
            for (int k = 0; k < K; k++) asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(u[j]) : "r"(u[(j + 1) % N]));
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < N; j++) s += a[j] + u[j];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
void run(const char *name, F launch, int ops_per_iter) {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1024);
    for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
        int threads = 128 * warps_per_smsp;
        if (threads > 1024) break;
        launch(1, threads, out, cyc);
        launch(1, threads, out, cyc);
        cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double per_warp_instr = (double)h / ((double)ITERS * ops_per_iter);
        printf("%-28s warps/SMSP=%d  cycles/instr/warp=%.2f  -> SMSP issue interval=%.2f cyc/instr\n", name, warps_per_smsp,
               per_warp_instr, per_warp_instr / warps_per_smsp);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    double h_in[64];
    for (int i = 0; i < 64; i++) h_in[i] = 1.0 + 1e-9 * i;
    double *in;
    cudaMalloc(&in, sizeof(h_in));
    cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    run("latency (1 chain, 16/iter)", [&](int g, int t, double *o, long long *c) { k_latency<<<g, t>>>(o, c, 1.0000001, 1e-9); }, 16);
    run("dfma rrr distinct N=8", [&](int g, int t, double *o, long long *c) { k_rrr<8><<<g, t>>>(o, c, in); }, 8);
    run("dfma rrr distinct N=4", [&](int g, int t, double *o, long long *c) { k_rrr<4><<<g, t>>>(o, c, in); }, 4);
    run("dfma rrr distinct N=2", [&](int g, int t, double *o, long long *c) { k_rrr<2><<<g, t>>>(o, c, in); }, 2);
    run("dfma shared ops N=8", [&](int g, int t, double *o, long long *c) { k_shared_ops<8><<<g, t>>>(o, c, in); }, 8);
    run("dadd N=8", [&](int g, int t, double *o, long long *c) { k_addmul<8, 0><<<g, t>>>(o, c, in); }, 8);
    run("dmul N=8", [&](int g, int t, double *o, long long *c) { k_addmul<8, 1><<<g, t>>>(o, c, in); }, 8);
    run("mixed dfma+1 lop3 N=8", [&](int g, int t, double *o, long long *c) { k_mixed<8, 1><<<g, t>>>(o, c, in); }, 8);
    run("mixed dfma+2 lop3 N=8", [&](int g, int t, double *o, long long *c) { k_mixed<8, 2><<<g, t>>>(o, c, in); }, 8);
    run("mixed dfma+3 lop3 N=8", [&](int g, int t, double *o, long long *c) { k_mixed<8, 3><<<g, t>>>(o, c, in); }, 8);
    return 0;
}
