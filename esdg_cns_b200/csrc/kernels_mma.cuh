// Small dense element-local operators (Vq, Pq, VhP, Ph, LIFT, Vf, Dr, Ds of the modal / CNS paths) applied
// to a whole element batch as ONE GEMM per operator on the FP64 tensor cores:
//     Y[c][m] (+)= sum_k A[m][k] X[c][k],   c = (element of the batch, field),  m < M,  k < K
// with mma.sync.m8n8k4.f64 (DMMA).  On B200 the DMMA peak equals the DFMA peak, so this is not about
// flops: one DMMA replaces 8 warp-wide DFMAs AND the 2 x 8 shared-memory loads that feed them, which
// is where the table-driven row loops spent their instructions (profiles/microbench/dense_apply.cu:
// 16x16 operator on 40 columns, 8.2 us -> 1.3 us per application, identical results).
// A lives in global memory, row-major, zero-padded to (Mp, Kp) = (ceil8(M), ceil4(K)) so that fragment
// loads need no guards (L1-resident: every warp of every CTA reads the same few KB); X and Y are the
// kernels' shared-memory arrays in their natural [column][row] layout.
#pragma once
#include <cuda_runtime.h>

namespace esdg {

struct Dense {
    const double* a;   // [Mp][Kp] row-major, zero padded
    int M, K, Mp, Kp;
};

template <bool ACC>
__device__ __forceinline__ void mma_apply(const Dense& A, const double* X, int ldx, double* Y, int ldy, int NC) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int mt_n = A.Mp >> 3, nt_n = (NC + 7) >> 3;
    const int lr = lane >> 2, lc = lane & 3;
    for (int t = warp; t < mt_n * nt_n; t += nw) {
        const int mt = t % mt_n, nt = t / mt_n;
        const int m = mt * 8 + lr;        // A row and C row of this lane
        const int cb = nt * 8 + lr;       // B column of this lane
        const int c0 = nt * 8 + 2 * lc;   // C columns c0, c0 + 1 of this lane
        double acc0 = 0.0, acc1 = 0.0;
        if (ACC && m < A.M) {
            if (c0 < NC) acc0 = Y[c0 * ldy + m];
            if (c0 + 1 < NC) acc1 = Y[(c0 + 1) * ldy + m];
        }
        const double* arow = A.a + m * A.Kp + lc;
        const double* xcol = X + cb * ldx + lc;
        const bool cok = cb < NC;
        // fragments of four k-steps are fetched together (A from global / L1, X from shared) before their DMMAs
        for (int k0 = 0; k0 < A.Kp; k0 += 16) {
            double a[4], b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + 4 * j;
                a[j] = k < A.Kp ? arow[k] : 0.0;
                b[j] = (cok && k + lc < A.K) ? xcol[k] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + 4 * j < A.Kp)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc0), "+d"(acc1)
                                 : "d"(a[j]), "d"(b[j]));
            }
        }
        if (m < A.M) {
            if (c0 < NC) Y[c0 * ldy + m] = acc0;
            if (c0 + 1 < NC) Y[(c0 + 1) * ldy + m] = acc1;
        }
    }
}

}  // namespace esdg
