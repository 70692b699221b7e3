// FP64 pipe micro-benchmark for B200 (sm_100a): dependent-issue latency of DFMA / DMUL / DADD,
// MUFU.RCP64H and the throughput reached by W warps per SM sub-partition with ILP independent chains.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_chain(double* out, int iters, double a, double b, long long* cyc) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_rcp_chain(double* out, int iters, long long* cyc) {
    double x = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            double y;
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
            x = y;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_lds_chain(int* out, int iters, long long* cyc) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7 + 1) & 1023;
    __syncthreads();
    int j = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) j = s[j];
    }
    long long t1 = clock64();
    out[threadIdx.x] = j;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, double* out, long long* cyc) {
    const int iters = 2000;
    k_chain<ILP><<<148, 32 * warps_per_sm>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 16.0 * ILP);
    // warp-instructions per cycle per sub-partition: warps_per_sm/4 warps each issuing ILP*16*iters
    const double wpc = (warps_per_sm / 4.0) * iters * 16.0 * ILP / c;
    printf("warps/SM %2d (%.2f per SMSP) ILP %d: %.2f cycles per DFMA per warp, %.3f warp-DFMA/cycle/SMSP (peak 0.5)\n",
           warps_per_sm, warps_per_sm / 4.0, ILP, per, wpc);
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8);
    cudaMalloc(&cyc, 8);
    for (int w : {1, 4, 8, 10, 12, 16, 24, 32}) {
        run<1>(w, out, cyc);
        run<2>(w, out, cyc);
        run<4>(w, out, cyc);
        run<8>(w, out, cyc);
    }
    long long c;
    k_rcp_chain<<<1, 32>>>(out, 2000, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.RCP64H dependent chain: %.2f cycles\n", c / 32000.0);
    k_lds_chain<<<1, 32>>>((int*)out, 2000, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS dependent chain: %.2f cycles\n", c / 32000.0);
    return 0;
}
