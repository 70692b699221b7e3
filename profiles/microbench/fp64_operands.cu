// FP64 pipe throughput vs operand pattern on B200: does a DFMA with three distinct register operands
// (the normal case in the flux kernels) issue at the same rate as one with constant / repeated operands?
// 8 warps per SM (2 per sub-partition), 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
#define N 8
template <int MODE>
__global__ void __launch_bounds__(64) k(double* out, int iters, double a, double b) {
    double x[N], y[N], z[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (i + threadIdx.x); z[i] = 1e-9 * (i + 1); }
    int w = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (MODE == 0) x[i] = fma(x[i], 1.0000001, 1e-9);          // immediates / constant bank
                if (MODE == 1) x[i] = fma(x[i], a, b);                     // two shared registers
                if (MODE == 2) x[i] = fma(x[i], y[i], z[i]);               // three distinct registers
                if (MODE == 3) x[i] = fma(y[i], z[(i + 1) % N], x[i]);     // three distinct, accumulate form
                if (MODE == 4) x[i] = x[i] * y[i];                         // DMUL two registers
                if (MODE == 5) x[i] = x[i] + y[i];                         // DADD two registers
                if (MODE == 6) { x[i] = fma(x[i], y[i], z[i]); w = w * 3 + i; }   // + one integer op per DFMA
                if (MODE == 7) { x[i] = fma(x[i], y[i], z[i]); y[i] = x[i] > 0.5 ? y[i] : z[i]; }   // + FP64 select
            }
        }
    }
    double s = w;
#pragma unroll
    for (int i = 0; i < N; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double* out, int ctas_per_sm) {
    const int iters = 4000, ctas = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, 64>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<MODE><<<ctas, 64>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)ctas * 2 * iters * 4 * N;               // FP64 warp-instructions
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("%-44s %d warps/SMSP: %.3f FP64 warp-inst/cycle/SMSP\n", name, ctas_per_sm * 2 / 4, winst / (cyc * 148 * 4));
}
int main() {
    double* out;
    cudaMalloc(&out, 148 * 16 * 64 * 8);
    for (int c : {4, 8}) {
        run<0>("DFMA x*imm+imm", out, c);
        run<1>("DFMA x*a+b (shared regs)", out, c);
        run<2>("DFMA x*y[i]+z[i] (3 distinct regs)", out, c);
        run<3>("DFMA y[i]*z[j]+x (3 distinct regs)", out, c);
        run<4>("DMUL x*y[i]", out, c);
        run<5>("DADD x+y[i]", out, c);
        run<6>("DFMA 3 regs + 1 IMAD each", out, c);
        run<7>("DFMA 3 regs + 1 FP64 select each", out, c);
    }
    return 0;
}
