// Tensor-product hex stage, one-shot kernel: one CTA = one batch of EB elements, every thread runs the phases of
// hex_common.cuh one after another (A -> C -> E -> F, block barriers in between).  Used for small element ranges
// and as the cross-check of the warp-specialised persistent kernel (kernels_hex_ws.cuh), which is bitwise
// identical to it (same per-item device functions, same summation order).  Also the traces-only kernel.
#pragma once
#include "hex_common.cuh"

namespace esdg {

template <int n>
struct LineDims {
    using G = HexGeom<n>;
    static constexpr int EB = n == 5 ? 2 : (n == 4 ? 2 : (n == 3 ? 4 : 8));
    static constexpr int THREADS = ((EB * G::NL + 31) / 32) * 32;
    static constexpr int MINB = n == 5 ? 2 : 4;   // CTAs per SM the register budget is sized for
    static constexpr int EN = EB * G::Nq, EF = EB * G::Nfq;
    // WV[7][EN] WF[8][EF] P[8][EF] A3[15][EN] (V[5][EN] aliases A3 block 0) geo[EB][40] lm lp w
    static constexpr int WV_DOUBLES = 7 * wv_stride<n, EB>();
    static constexpr int SMEM_DOUBLES = WV_DOUBLES + 16 * EF + 15 * EN + EB * HEX_SGEO + 3 * n;
    static constexpr int SMEM_DOUBLES_CURVED = SMEM_DOUBLES + 1 + 3 * (n + 2) * THREADS;   // + one column of geometric factors per line thread
};

// own traces (bulk copies when every run is 16-byte aligned, else 8-byte cp.async) and exterior traces (gather
// through mapP) of the batch go in flight; `bar` must have been initialised for ONE arrival (the expect_tx)
// plus the transaction bytes
template <int n, int EB, int NT, bool GATHER = true>
__device__ __forceinline__ void issue_trace_loads(int t, int nb, long long e0, const HexParams& p, double* sWF, double* sP,
                                                  unsigned long long* bar, bool bulk, const int (&ipv)[(EB * HexGeom<n>::Nfq + NT - 1) / NT]) {
    using G = HexGeom<n>;
    constexpr int Nfq = G::Nfq, EF = EB * Nfq, MAXF = (EF + NT - 1) / NT;
    const long long gf0 = e0 * Nfq;
    if (bulk) {
        const unsigned bf = (unsigned)(nb * Nfq * 8);
        if (t == 0) mbar_expect_tx(bar, HEX_TP * bf);
        if ((t & 31) == 0) {
            for (int c = t >> 5; c < HEX_TP; c += NT / 32) bulk_g2s(sWF + c * EF, p.Tcur + c * p.t_stride + gf0, bf, bar);
        }
    } else {
        if (t == 0) mbar_arrive(bar);
        for (int it = t; it < nb * Nfq; it += NT) {
#pragma unroll
            for (int q = 0; q < HEX_TP; ++q) cp_async8(sWF + q * EF + it, p.Tcur + q * p.t_stride + gf0 + it);
        }
    }
    if (GATHER) {
#pragma unroll
        for (int r = 0; r < MAXF; ++r) {
            const int it = t + r * NT;
            if (it < nb * Nfq) {
#pragma unroll
                for (int q = 0; q < HEX_TP; ++q) cp_async8(sP + q * EF + it, p.Tcur + q * p.t_stride + ipv[r]);
            }
        }
    }
}

__device__ __forceinline__ bool trace_runs_aligned(const HexParams& p) {
    return ((p.t_stride & 1) == 0) && ((((size_t)p.Tcur) & 15) == 0);
}

// CURVED: per-pair averaged geometric factors, per-point normals and Jacobians (line_compute_curved, hex_common.cuh)
template <int n, bool DIAG, bool CURVED>
__device__ __forceinline__ void hex_stage_line_body(const HexParams& p) {
    using D = LineDims<n>;
    using G = HexGeom<n>;
    constexpr int Nq = G::Nq, Nfq = G::Nfq, NL = G::NL, EB = D::EB, NT = D::THREADS, EN = D::EN, EF = D::EF;
    extern __shared__ __align__(16) double sm[];
    __shared__ unsigned long long mbar;
    __shared__ double red[32];
    double* sWV = sm;                   // [7][EN]  rho, u/2, v/2, w/2, beta, log rho, log beta at volume nodes
    double* sWF = sWV + D::WV_DOUBLES;  // [8][EF]  the same + lambda at the face points (own traces)
    double* sP = sWF + HEX_TP * EF;     // [8][EF]  exterior traces
    double* sA = sP + HEX_TP * EF;      // [3][5][EN] per-direction line results; block 0 later holds V[5][EN]
    double* sGeo = sA + 15 * EN;        // [EB][40]
    double* s_lm = sGeo + EB * HEX_SGEO;
    double* s_lp = s_lm + n;
    double* s_w = s_lp + n;
    double* sGL = s_w + n + ((3 * n) & 1);   // CURVED: [3 (n + 2)][NT]
    const int tid = threadIdx.x;
    long long e0;
    int nb;
    range_batch(p, EB, blockIdx.x, e0, nb);
    constexpr int MAXF = (EF + NT - 1) / NT;
    int ipv[MAXF];
#pragma unroll
    for (int r = 0; r < MAXF; ++r) {
        const int it = tid + r * NT;
        ipv[r] = it < nb * Nfq ? p.mapP[e0 * Nfq + it] : 0;
    }
    if (tid == 0) {
        mbar_init(&mbar, 1);
        mbar_fence_init();
    }
    if (tid < n) {
        s_lm[tid] = p.lm[tid];
        s_lp[tid] = p.lp[tid];
        s_w[tid] = p.w[tid];
    }
    __syncthreads();   // mbarrier initialised before any warp issues a bulk copy on it
    const double c_igm = 1.0 / (2.0 * (p.gamma - 1.0));
    bool bad = false;
    const bool bulk = trace_runs_aligned(p);
    issue_trace_loads<n, EB, NT>(tid, nb, e0, p, sWF, sP, &mbar, bulk, ipv);
    if constexpr (CURVED) {
        if (tid < nb * NL) line_geometry_load<n, EB>(tid, e0, p, sGL + tid, NT);
    }
    cp_async_commit();
    stage_geo<EB>(tid, NT, nb, e0, p, sGeo);
    // ---- A ----
    for (int it = tid; it < nb * Nq; it += NT) node_prims<n, EB>(it, e0 * Nq + it, p, sWV, bad);
    cp_async_wait_all();
    __syncthreads();
    mbar_wait(&mbar, 0);
    // ---- C ----
    if (tid < nb * NL) {
        double acc[n][5];
        if constexpr (CURVED) line_compute_curved<n, EB>(tid, e0, sWV, sWF, sP, sGL + tid, NT, p, c_igm, acc);
        else line_compute<n, EB, DIAG>(tid, sWV, sWF, sP, sGeo, p, c_igm, acc);
        line_store<n, EB>(tid, acc, sA);
    }
    __syncthreads();
    // ---- E ----
    double tsum = 0.0;
    for (int it = tid; it < nb * Nq; it += NT) {
        NodeIn in;
        node_load(e0 * Nq + it, p, in);
        node_update<n, EB, CURVED>(it, e0 * Nq + it, p, in, sA, sA, sGeo, s_w, tsum, bad);
    }
    if (p.mode == 1) {
        __syncthreads();
        // ---- F ----
        for (int l = tid; l < nb * NL; l += NT) line_traces<n, EB, CURVED>(l, e0, p, sA, sGeo, s_lm, s_lp);
    }
    if (p.want_rhstest) {
        for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(0xffffffffu, tsum, o);
        if ((tid & 31) == 0) red[tid >> 5] = tsum;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int wv = 0; wv < NT / 32; ++wv) t += red[wv];
            p.partial[e0] = t;
        }
    }
    if (bad) atomicOr(p.bad, 1);
}

template <int n, bool DIAG>
__global__ void __launch_bounds__(LineDims<n>::THREADS, LineDims<n>::MINB) k_hex_stage_line(const __grid_constant__ HexParams p) {
    hex_stage_line_body<n, DIAG, false>(p);
}
// curved elements: 168 registers = three warps per SM sub-partition, i.e. two 160-thread CTAs per SM at N = 4 (at 200 registers
// only 8 warps fit an SM and the second CTA does not: measured 42.9 -> see profiles/r2_curved_hex.txt)
template <int n>
__global__ void __launch_bounds__(LineDims<n>::THREADS, n == 5 ? 2 : 3) k_hex_stage_line_curved(const __grid_constant__ HexParams p) {
    hex_stage_line_body<n, false, true>(p);
}

// Traces (+ log stash) of a given state: runs once after an upload and in the rhs-only entry point.
template <int n>
struct TraceDims {
    using G = HexGeom<n>;
    static constexpr int EB = (128 / G::Nq) > 0 ? (128 / G::Nq) : 1;
    static constexpr int THREADS = ((EB * G::Nq + 31) / 32) * 32;
    static constexpr int SMEM_DOUBLES = EB * 5 * G::Nq + EB * HEX_SGEO + 2 * n;
};

template <int n, bool CURVED = false>
__global__ void __launch_bounds__(TraceDims<n>::THREADS) k_hex_traces(const __grid_constant__ HexParams p) {
    using D = TraceDims<n>;
    using G = HexGeom<n>;
    constexpr int EB = D::EB, NT = D::THREADS;
    extern __shared__ __align__(16) double sm[];
    double* sV = sm;                               // [5][EB*Nq]
    double* sGeo = sV + EB * 5 * G::Nq;
    double* s_lm = sGeo + EB * HEX_SGEO;
    double* s_lp = s_lm + n;
    const int tid = threadIdx.x;
    if (tid < n) { s_lm[tid] = p.lm[tid]; s_lp[tid] = p.lp[tid]; }
    long long e0;
    int nb;
    range_batch(p, EB, blockIdx.x, e0, nb);
    bool bad = false;
    stage_geo<EB>(tid, NT, nb, e0, p, sGeo);
    for (int it = tid; it < nb * G::Nq; it += NT) node_entropy<n, EB>(it, e0 * G::Nq + it, p, sV, bad);
    __syncthreads();
    for (int l = tid; l < nb * G::NL; l += NT) line_traces<n, EB, CURVED>(l, e0, p, sV, sGeo, s_lm, s_lp);
    if (bad) atomicOr(p.bad, 1);
}

}  // namespace esdg
