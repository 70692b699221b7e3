// FP64 (non-tensor DFMA) peak micro-benchmark: the denominator of the FP64 side of the roofline.
// MEASURED_PEAKS.json has no FP64 number, so bench.py measures it on the box it runs on.
#include <cuda_runtime.h>

#include "../../include/esdg_b200.h"

namespace {
// 4 independent chains, 8x unrolled, multiplier/addend as constant operands: the configuration that reached
// 0.495 warp-DFMA per cycle per SM sub-partition in profiles/microbench (register operands top out at 0.46)
__global__ void __launch_bounds__(128) k_dfma_peak(double* out, int iters) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = fma(x[i], 1.0000001, 1e-9);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (x[0] + x[1]) + (x[2] + x[3]);
}
}  // namespace

extern "C" int esdg_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return ESDG_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return ESDG_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ESDG_ECUDA;
    const int blocks = prop.multiProcessorCount * 16, threads = 128, iters = 1 << 12;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return ESDG_ENOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 32.0 * iters * (double)blocks * threads;
        if (rep > 0 && ms > 0) best = fl / (ms * 1e-3) / 1e12 > best ? fl / (ms * 1e-3) / 1e12 : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? ESDG_OK : ESDG_ECUDA;
}
