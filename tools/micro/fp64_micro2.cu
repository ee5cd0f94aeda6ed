// fp64_micro2.cu — marginal cost of extra instructions next to a saturated FP64 pipe on B200.
// Base: N independent chains of DFMA with two register operands + one constant-bank operand (the Horner
// form a = fma(z, a, c[k])).  Variants add K extra instructions per DFMA of a given class.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__constant__ double kC[16] = {1.000000001, 1e-9, 1.000000002, 2e-9, 1.000000003, 3e-9, 1.000000004, 4e-9,
                              1.000000005, 5e-9, 1.000000006, 6e-9, 1.000000007, 7e-9, 1.000000008, 8e-9};

enum { NONE = 0, LOP_1R, LOP_2R, SEL_2R, IMAD_1R, FMUL_2R, DFMA_RRR_BASE, MOV_1R, IADD_2R, FSEL64 };

template <int N, int CLS, int K>
__global__ void kern(double *out, long long *cyc, const double *in) {
    double a[N], z[N], y[N];
    unsigned u[N], v[N];
    float f[N], g[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        a[j] = in[j] + threadIdx.x;
        z[j] = in[N + j];
        y[j] = in[2 * N + j] + threadIdx.x;
        u[j] = threadIdx.x * 7 + j;
        v[j] = threadIdx.x * 13 + j + 5;
        f[j] = 1.0f + 1e-7f * threadIdx.x;
        g[j] = 1.0f + 1e-7f * j;
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            if (CLS == DFMA_RRR_BASE) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[j]) : "d"(z[j]), "d"(y[j]));
            else a[j] = fma(z[j], a[j], kC[j]);
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int jj = (j + k) % N;
                if (CLS == LOP_1R) asm volatile("lop3.b32 %0, %0, 0x0f0f0f0f, 0x5a5a5a5a, 0x96;" : "+r"(u[jj]));
                if (CLS == LOP_2R) asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(u[jj]) : "r"(v[jj]));
                if (CLS == SEL_2R) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(u[jj]) : "r"(v[jj]), "r"(i));
                if (CLS == IMAD_1R) asm volatile("mad.lo.u32 %0, %0, 5, 7;" : "+r"(u[jj]));
                if (CLS == FMUL_2R) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[jj]) : "f"(g[jj]));
                if (CLS == MOV_1R) asm volatile("shl.b32 %0, %1, 30;" : "=r"(u[jj]) : "r"(v[jj]));
                if (CLS == IADD_2R) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[jj]) : "r"(v[jj]));
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < N; j++) s += a[j] + u[j] + f[j] + v[j];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CLS, int K>
void run(const char *name, const double *in) {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1024);
    const int threads = 1024;  // 8 warps per SMSP on one SM
    kern<8, CLS, K><<<1, threads>>>(out, cyc, in);
    kern<8, CLS, K><<<1, threads>>>(out, cyc, in);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s K=%d  SMSP cycles per (DFMA + K extra) = %.2f\n", name, K, (double)h / ((double)ITERS * 8) / 8.0);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    double h_in[64];
    for (int i = 0; i < 64; i++) h_in[i] = 1.0 + 1e-9 * i;
    double *in;
    cudaMalloc(&in, sizeof(h_in));
    cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    run<NONE, 0>("dfma rr+const", in);
    run<DFMA_RRR_BASE, 0>("dfma rrr", in);
#define ALLK(C, NAME) run<C, 1>(NAME, in); run<C, 2>(NAME, in); run<C, 4>(NAME, in);
    ALLK(LOP_1R, "dfma rrc + lop3 (1 reg)")
    ALLK(LOP_2R, "dfma rrc + lop3 (2 reg)")
    ALLK(SEL_2R, "dfma rrc + setp+selp (2 reg)")
    ALLK(IMAD_1R, "dfma rrc + imad (1 reg)")
    ALLK(FMUL_2R, "dfma rrc + fmul f32 (2 reg)")
    ALLK(MOV_1R, "dfma rrc + shl (1 reg)")
    ALLK(IADD_2R, "dfma rrc + iadd (2 reg)")
    return 0;
}
