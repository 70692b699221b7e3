// Small dense operator application batched over an element batch: Y[c][m] = sum_k A[m][k] X[c][k],
// A = 16x16 (the size of Vq / Pq / Dr / Ds / LIFT for N=3 quads, N=4 triangles), c = 40 columns
// (10 elements x 4 fields) in shared memory -- the shape of the modal / CNS kernels' operator phases.
// Variant 0: table-driven row loop as in kernels_cns.cuh (ell_row: col + val from global, 4 LDS + 4 DFMA per entry)
// Variant 1: dense row loop, A row in registers (no column indices), 4 fields per thread
// Variant 2: FP64 tensor cores, mma.sync.m8n8k4.f64, A fragments from global (L1), X fragments from shared
#include <cstdio>
#include <cuda_runtime.h>
constexpr int M = 16, K = 16, NC = 40, LD = 20;   // LD: padded leading dimension (conflict-free fragment loads)
__global__ void k_ell(const int* __restrict__ col, const double* __restrict__ val, double* out, int reps) {
    __shared__ double X[NC * LD], Y[NC * LD];
    for (int i = threadIdx.x; i < NC * LD; i += blockDim.x) X[i] = 1e-3 * (i % 97);
    __syncthreads();
    for (int r = 0; r < reps; ++r) {
        for (int i = threadIdx.x; i < 10 * M; i += blockDim.x) {
            const int eb = i / M, row = i % M;
            double y[4] = {0, 0, 0, 0};
            for (int k = 0; k < K; ++k) {
                const int c = col[row * K + k];
                const double a = val[row * K + k];
#pragma unroll
                for (int f = 0; f < 4; ++f) y[f] = fma(a, X[(eb * 4 + f) * LD + c], y[f]);
            }
#pragma unroll
            for (int f = 0; f < 4; ++f) Y[(eb * 4 + f) * LD + row] = y[f];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < NC * LD; i += blockDim.x) X[i] = Y[i] * 0.01 + 1e-3;
        __syncthreads();
    }
    if (threadIdx.x < NC) out[blockIdx.x * NC + threadIdx.x] = X[threadIdx.x * LD];
}
__global__ void k_dense(const double* __restrict__ A, double* out, int reps) {
    __shared__ double X[NC * LD], Y[NC * LD];
    for (int i = threadIdx.x; i < NC * LD; i += blockDim.x) X[i] = 1e-3 * (i % 97);
    __syncthreads();
    for (int r = 0; r < reps; ++r) {
        for (int i = threadIdx.x; i < 10 * M; i += blockDim.x) {
            const int eb = i / M, row = i % M;
            double a[K];
#pragma unroll
            for (int k = 0; k < K; ++k) a[k] = A[row * K + k];
            double y[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int f = 0; f < 4; ++f) y[f] = fma(a[k], X[(eb * 4 + f) * LD + k], y[f]);
#pragma unroll
            for (int f = 0; f < 4; ++f) Y[(eb * 4 + f) * LD + row] = y[f];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < NC * LD; i += blockDim.x) X[i] = Y[i] * 0.01 + 1e-3;
        __syncthreads();
    }
    if (threadIdx.x < NC) out[blockIdx.x * NC + threadIdx.x] = X[threadIdx.x * LD];
}
__global__ void k_mma(const double* __restrict__ A, double* out, int reps) {
    __shared__ double X[NC * LD], Y[NC * LD];
    for (int i = threadIdx.x; i < NC * LD; i += blockDim.x) X[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = 0; r < reps; ++r) {
        for (int t = warp; t < (M / 8) * (NC / 8); t += nw) {      // 2 x 5 output tiles
            const int mt = t % (M / 8), nt = t / (M / 8);
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < K; k0 += 4) {
                const double a = A[(mt * 8 + lane / 4) * K + k0 + lane % 4];
                const double b = X[(nt * 8 + lane / 4) * LD + k0 + lane % 4];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
            }
            Y[(nt * 8 + 2 * (lane % 4)) * LD + mt * 8 + lane / 4] = c0;
            Y[(nt * 8 + 2 * (lane % 4) + 1) * LD + mt * 8 + lane / 4] = c1;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < NC * LD; i += blockDim.x) X[i] = Y[i] * 0.01 + 1e-3;
        __syncthreads();
    }
    if (threadIdx.x < NC) out[blockIdx.x * NC + threadIdx.x] = X[threadIdx.x * LD];
}
int main() {
    double hA[M * K]; int hc[M * K];
    for (int i = 0; i < M * K; ++i) { hA[i] = 0.01 * ((i * 7) % 13 - 6); hc[i] = i % K; }
    double *A, *out; int* col;
    cudaMalloc(&A, sizeof hA); cudaMalloc(&col, sizeof hc); cudaMalloc(&out, 148 * 8 * NC * 8);
    cudaMemcpy(A, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(col, hc, sizeof hc, cudaMemcpyHostToDevice);
    const int reps = 2000, ctas = 148 * 4, threads = 160;
    double ref[3][NC];
    for (int v = 0; v < 3; ++v) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int pass = 0; pass < 2; ++pass) {
            cudaEventRecord(e0);
            if (v == 0) k_ell<<<ctas, threads>>>(col, A, out, reps);
            if (v == 1) k_dense<<<ctas, threads>>>(A, out, reps);
            if (v == 2) k_mma<<<ctas, threads>>>(A, out, reps);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(ref[v], out, NC * 8, cudaMemcpyDeviceToHost);
        const double applies = (double)ctas * reps;
        printf("%-28s %.3f ms  %.1f ns per 16x16x40 apply per CTA-slot, %.2f TFLOP/s  check %.12e\n",
               v == 0 ? "ELL row loop (current)" : v == 1 ? "dense rows, A in registers" : "DMMA m8n8k4", ms,
               ms * 1e6 / reps, applies * 2.0 * M * K * NC / (ms * 1e-3) * 1e-12, ref[v][3]);
    }
    return 0;
}
