// dmma_micro.cu — does the FP64 tensor path (mma.sync.m8n8k4.f64, SASS DMMA) run beside the FP64 vector pipe on B200?
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096

template <int NMMA, int NFMA>
__global__ void kern(double *out, long long *cyc, const double *in) {
    double a = in[threadIdx.x & 7], b = in[8 + (threadIdx.x & 7)];
    double c[NMMA > 0 ? NMMA : 1][2];
    double f[NFMA > 0 ? NFMA : 1];
    double m = in[20];
    for (int i = 0; i < NMMA; i++) c[i][0] = c[i][1] = 0.0;
    for (int i = 0; i < NFMA; i++) f[i] = in[i] + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < (NMMA > NFMA ? NMMA : NFMA); i++) {
            if (i < NMMA)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            if (i < NFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(m), "d"(a));
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < NMMA; i++) s += c[i][0] + c[i][1];
    for (int i = 0; i < NFMA; i++) s += f[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NMMA, int NFMA>
void run(const double *in) {
    double *out; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    kern<NMMA, NFMA><<<1, 1024>>>(out, cyc, in);
    kern<NMMA, NFMA><<<1, 1024>>>(out, cyc, in);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("per iteration: %d DMMA + %d DFMA : %.2f SMSP cycles  (err=%s)\n", NMMA, NFMA, (double)h / ITERS / 8.0,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i;
    double *in; cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<8, 0>(in); run<0, 8>(in); run<8, 8>(in); run<4, 8>(in); run<2, 8>(in); run<1, 8>(in);
    return 0;
}
