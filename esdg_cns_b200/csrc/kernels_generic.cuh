// Generic (operator-table driven) kernels: any element type whose operators
// arrive through esdg_set_operator -- quads, hexes in any node ordering, and
// modal triangles.  One CTA owns a batch of EB consecutive elements; every
// phase is a strided loop over (element, row) items followed by a barrier, all
// element data staged in shared memory.
//
// Path covered (reference call stack):
//   phase 1  VU = v_ufun(Q) [tri: v_ufun(Vq*Q)]              dg2D_euler_quad.jl:149 / tri :138
//   phase 2  Uf = u_vfun(Ef*VU) [tri: Uh = u_vfun(VhP*VU)]    :150 / tri :139-140
//            Qh = (rho,u,v[,w],beta), logs                     :154-155, :111
//   phase 3  QP = QM[mapP]; LF; flux*                          :158-168
//   phase 4  sparse/dense Hadamard sum                         :102-138 / tri :88-126
//   phase 5  rhs = -(Lf*flux + vs*Ph*QF)/J                     :170,180,184
//   phase 6  rhstest / fused LSRK stage                        :186-191 / :205-206
#pragma once
#include "device_tables.h"
#include "physics.cuh"

#define ESDG_GENERIC_MAX_THREADS 320

namespace esdg {

template <int DIM>
struct GenericSmem {
    static constexpr int NF = DIM + 2;
    double *U, *V, *W, *Uf, *Fl, *QF, *VP, *R;
    double* Vface = nullptr;              // optional [EB][NF][Nfq]: projected entropy variables at face points (CNS)
    __device__ GenericSmem(double* base, const GenericParams& p) {
        const int EB = p.EB;
        U = base;                         // [EB][NF][Ns]
        V = U + EB * NF * p.Ns;           // [EB][NF][Nq]
        W = V + EB * NF * p.Nq;           // [EB][NF+2][Nh]
        Uf = W + EB * (NF + 2) * p.Nh;    // [EB][NF][Nfq]
        Fl = Uf + EB * NF * p.Nfq;        // [EB][NF][Nfq]
        QF = Fl + EB * NF * p.Nfq;        // [EB][NF][Nh]
        VP = QF + EB * NF * p.Nh;         // [EB][NF][Nq] (modal only: projected entropy vars, volume rows)
        R = VP + EB * NF * p.Nq;          // [EB][NF][Ns] (modal only: rhs for the rhstest quadrature)
    }
};

inline size_t generic_smem_bytes(int DIM, int EB, int Ns, int Nq, int Nfq, int Nh, int collocated) {
    const int NF = DIM + 2;
    size_t n = (size_t)EB * (NF * Ns + NF * Nq + (NF + 2) * Nh + 2 * NF * Nfq + NF * Nh + (collocated ? 0 : NF * (Ns + Nq)));
    return n * sizeof(double);
}

// ---- operator tables -> shared memory ------------------------------------------------------------
// The ELL tables are read once per (row, entry) by every element of the batch; from global memory
// (L1-cached) those reads were the top stall of the table-driven kernels (lg_throttle / long_scoreboard).
__host__ __device__ inline size_t ell_stage_bytes(int rows, int width) {
    const size_t n = (size_t)rows * width;
    return n * 8 + ((n + 1) / 2) * 8 + (((size_t)rows + 1) / 2) * 8;
}
__host__ __device__ inline size_t pair_stage_bytes(int rows, int width) {
    const size_t n = (size_t)rows * width;
    return n * 24 + ((n + 1) / 2) * 8 + (((size_t)rows + 1) / 2) * 8;
}

__device__ __forceinline__ Ell stage_ell(const Ell& g, char*& sp) {
    const int n = g.rows * g.width;
    double* v = (double*)sp;
    int* c = (int*)(sp + (size_t)n * 8);
    int* cn = (int*)(sp + (size_t)n * 8 + (size_t)((n + 1) / 2) * 8);
    sp += ell_stage_bytes(g.rows, g.width);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { v[i] = g.val[i]; c[i] = g.col[i]; }
    for (int i = threadIdx.x; i < g.rows; i += blockDim.x) cn[i] = g.cnt[i];
    return Ell{g.rows, g.width, cn, c, v};
}

__device__ __forceinline__ PairEll stage_pairs(const PairEll& g, char*& sp) {
    const int n = g.rows * g.width;
    double* v = (double*)sp;
    int* c = (int*)(sp + (size_t)n * 24);
    int* cn = (int*)(sp + (size_t)n * 24 + (size_t)((n + 1) / 2) * 8);
    sp += pair_stage_bytes(g.rows, g.width);
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) v[i] = g.c[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) c[i] = g.col[i];
    for (int i = threadIdx.x; i < g.rows; i += blockDim.x) cn[i] = g.cnt[i];
    return PairEll{g.rows, g.width, cn, c, v};
}

enum { ST_VQ = 1, ST_PROJ = 2, ST_PH = 4, ST_LF = 8, ST_QP = 16 };

// returns the running pointer so that callers (CNS) can append their own tables; the caller syncs
__device__ __forceinline__ char* stage_generic_tables(GenericParams& p, double* smem, int mask) {
    char* sp = (char*)smem + p.table_off;
    if (!p.stage_tables) return sp;
    if (mask & ST_VQ) p.Vq = stage_ell(p.Vq, sp);
    if (mask & ST_PROJ) p.Proj = stage_ell(p.Proj, sp);
    if (mask & ST_PH) p.Ph = stage_ell(p.Ph, sp);
    if (mask & ST_LF) p.Lf = stage_ell(p.Lf, sp);
    if (mask & ST_QP) p.Qp = stage_pairs(p.Qp, sp);
    return sp;
}

// Phases 0-2: stage Q, entropy-project, build Qh (+logs) at all hybridized points.
template <int DIM>
__device__ __forceinline__ void project_batch(const GenericParams& p, GenericSmem<DIM>& s, long long e0, int nb,
                                              bool& bad) {
    constexpr int NF = DIM + 2;
    const int tid = threadIdx.x, nt = blockDim.x;
    // phase 0: Q -> smem (contiguous per field across the batch: coalesced)
    for (int f = 0; f < NF; ++f) {
        const double* src = p.Q + f * p.q_stride + e0 * p.Ns;
        for (int i = tid; i < nb * p.Ns; i += nt) {
            const int eb = i / p.Ns, k = i - eb * p.Ns;
            s.U[(eb * NF + f) * p.Ns + k] = src[i];
        }
    }
    __syncthreads();
    if (!p.collocated) {
        // ---- modal path (dg2D_euler_tri.jl:138-140): the two projections are batch GEMMs on the tensor cores
        // phase 1: Uq = Vq*Q, then v_ufun pointwise (in place in s.V)
        mma_apply<false>(p.dVq, s.U, p.Ns, s.V, p.Nq, nb * NF);
        __syncthreads();
        for (int i = tid; i < nb * p.Nq; i += nt) {
            const int eb = i / p.Nq, q = i - eb * p.Nq;
            double U[NF], V[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) U[f] = s.V[(eb * NF + f) * p.Nq + q];
            v_of_u<DIM>(U, p.gamma, p.lg1, V, bad);
#pragma unroll
            for (int f = 0; f < NF; ++f) s.V[(eb * NF + f) * p.Nq + q] = V[f];
        }
        __syncthreads();
        // phase 2: projected entropy variables at all hybridized points (s.QF as scratch), then u_vfun etc.
        mma_apply<false>(p.dProj, s.V, p.Nq, s.QF, p.Nh, nb * NF);
        __syncthreads();
        for (int i = tid; i < nb * p.Nh; i += nt) {
            const int eb = i / p.Nh, r = i - eb * p.Nh;
            double U[NF], V[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) V[f] = s.QF[(eb * NF + f) * p.Nh + r];
            u_of_v<DIM>(V, p.gamma, p.lg1, U);
            if (r < p.Nq) {
                // projected entropy variables at volume points are the rhstest weights (tri :180)
#pragma unroll
                for (int f = 0; f < NF; ++f) s.VP[(eb * NF + f) * p.Nq + r] = V[f];
            } else if (s.Vface) {
#pragma unroll
                for (int f = 0; f < NF; ++f) s.Vface[(eb * NF + f) * p.Nfq + (r - p.Nq)] = V[f];
            }
            Prim<DIM> P;
            prim_of_u<DIM>(U, p.gamma, P);
            double* w = s.W + (size_t)eb * (NF + 2) * p.Nh + r;
            w[0] = P.rho;
#pragma unroll
            for (int d = 0; d < DIM; ++d) w[(1 + d) * p.Nh] = P.vel[d];
            w[(1 + DIM) * p.Nh] = P.beta;
            w[(2 + DIM) * p.Nh] = P.lrho;
            w[(3 + DIM) * p.Nh] = P.lbeta;
            if (r >= p.Nq) {
#pragma unroll
                for (int f = 0; f < NF; ++f) s.Uf[(eb * NF + f) * p.Nfq + (r - p.Nq)] = U[f];
            }
        }
        __syncthreads();
        return;
    }
    // ---- collocated path (dg2D_euler_quad.jl:149-155): state at the quadrature nodes, Ef is N+1 sparse
    // phase 1: entropy variables at the volume quadrature points
    for (int i = tid; i < nb * p.Nq; i += nt) {
        const int eb = i / p.Nq, q = i - eb * p.Nq;
        double U[NF], V[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) U[f] = s.U[(eb * NF + f) * p.Ns + q];
        v_of_u<DIM>(U, p.gamma, p.lg1, V, bad);
#pragma unroll
        for (int f = 0; f < NF; ++f) s.V[(eb * NF + f) * p.Nq + q] = V[f];
    }
    __syncthreads();
    // phase 2: volume rows are the state itself, face rows u_vfun(Ef*VU) -> primitive + logs
    for (int i = tid; i < nb * p.Nh; i += nt) {
        const int eb = i / p.Nh, r = i - eb * p.Nh;
        double U[NF];
        if (r < p.Nq) {
#pragma unroll
            for (int f = 0; f < NF; ++f) U[f] = s.U[(eb * NF + f) * p.Ns + r];
        } else {
            const int pr = r - p.Nq;
            double V[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) V[f] = 0.0;
            const int n = p.Proj.cnt[pr];
            for (int k = 0; k < n; ++k) {
                const int c = p.Proj.col[pr * p.Proj.width + k];
                const double a = p.Proj.val[pr * p.Proj.width + k];
#pragma unroll
                for (int f = 0; f < NF; ++f) V[f] = fma(a, s.V[(eb * NF + f) * p.Nq + c], V[f]);
            }
            u_of_v<DIM>(V, p.gamma, p.lg1, U);
        }
        Prim<DIM> P;
        prim_of_u<DIM>(U, p.gamma, P);
        double* w = s.W + (size_t)eb * (NF + 2) * p.Nh + r;
        w[0] = P.rho;
#pragma unroll
        for (int d = 0; d < DIM; ++d) w[(1 + d) * p.Nh] = P.vel[d];
        w[(1 + DIM) * p.Nh] = P.beta;
        w[(2 + DIM) * p.Nh] = P.lrho;
        w[(3 + DIM) * p.Nh] = P.lbeta;
        if (r >= p.Nq) {
#pragma unroll
            for (int f = 0; f < NF; ++f) s.Uf[(eb * NF + f) * p.Nfq + (r - p.Nq)] = U[f];
        }
    }
    __syncthreads();
}

// Phase 4, symmetric form (the loop of CompressibleNS/dg2D_CNS_cavity_optimized.jl:326-347): the hybridized
// operators are skew and the two-point flux is symmetric, so pair {i,j} contributes +c.F(i,j) to row i and
// -c.F(i,j) to row j.  One thread per hybridized row; in every round of the host-built schedule a thread evaluates at
// most one pair, keeps its own contribution in registers and subtracts the partner's from the partner's slot of
// s.QF -- no two threads address the same slot in a round, so plain read-modify-write with a barrier per round is
// race-free and the summation order is fixed (bitwise reproducible).  Halves the flux evaluations of the row-wise form.
template <int DIM>
__device__ __forceinline__ void hadamard_sym(const GenericParams& p, GenericSmem<DIM>& s, long long e0, int nb) {
    constexpr int NF = DIM + 2;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < nb * NF * p.Nh; i += nt) s.QF[i] = 0.0;
    __syncthreads();
    // elements with at most 32 hybridized rows (triangles up to N = 4): one WARP per element, so the rounds need a
    // warp barrier only (a row and its partner's slot belong to the same element)
    const bool warp_elems = p.Nh <= 32 && nt >= 32 * nb;
    const bool act = warp_elems ? ((tid >> 5) < nb && (tid & 31) < p.Nh) : tid < nb * p.Nh;
    const int eb = !act ? 0 : (warp_elems ? tid >> 5 : tid / p.Nh), r = !act ? 0 : (warp_elems ? tid & 31 : tid - eb * p.Nh);
    const long long e = e0 + eb;
    const double* w = s.W + (size_t)eb * (NF + 2) * p.Nh;
    Prim<DIM> Pi;
    Pi.rho = w[r];
#pragma unroll
    for (int d = 0; d < DIM; ++d) Pi.vel[d] = w[(1 + d) * p.Nh + r];
    Pi.beta = w[(1 + DIM) * p.Nh + r];
    Pi.lrho = w[(2 + DIM) * p.Nh + r];
    Pi.lbeta = w[(3 + DIM) * p.Nh + r];
    double acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.0;
    double Gr[DIM * DIM];
    {
        const double* G = p.geo + e * p.geo_stride;
#pragma unroll
        for (int q = 0; q < DIM * DIM; ++q) Gr[q] = G[q];
    }
    const bool curved = DIM == 3 && p.curved;
    const double* Ge = curved ? p.geoG + (size_t)e * (DIM * DIM) * p.Nh : nullptr;
    double Gi[DIM * DIM];
    if (curved) {
#pragma unroll
        for (int q = 0; q < DIM * DIM; ++q) Gi[q] = Ge[q * p.Nh + r];
    }
    double* qf = s.QF + (size_t)eb * NF * p.Nh;
    // partner and coefficients of a round are requested one round ahead (table reads from global memory / L1: their
    // latency sat on the critical path of every round, profiles/r2_tri_generic_rhs_ncu_summary.txt)
    int jn = act ? p.sym_dst[r] : -1;
    double cn[DIM];
#pragma unroll
    for (int kk = 0; kk < DIM; ++kk) cn[kk] = act ? p.sym_c[(size_t)r * 3 + kk] : 0.0;
    for (int t = 0; t < p.sym_rounds; ++t) {
        const int j = jn;
        double cq[DIM];
#pragma unroll
        for (int kk = 0; kk < DIM; ++kk) cq[kk] = cn[kk];
        if (act && t + 1 < p.sym_rounds) {
            jn = p.sym_dst[(t + 1) * p.Nh + r];
#pragma unroll
            for (int kk = 0; kk < DIM; ++kk) cn[kk] = p.sym_c[(size_t)((t + 1) * p.Nh + r) * 3 + kk];
        }
        if (j >= 0) {
            if (curved) {
#pragma unroll
                for (int q = 0; q < DIM * DIM; ++q) Gr[q] = .5 * (Gi[q] + Ge[q * p.Nh + j]);
            }
            double c[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double a = 0.0;
#pragma unroll
                for (int kk = 0; kk < DIM; ++kk) a = fma(Gr[d * DIM + kk], cq[kk], a);
                c[d] = a;
            }
            Prim<DIM> Pj;
            Pj.rho = w[j];
#pragma unroll
            for (int d = 0; d < DIM; ++d) Pj.vel[d] = w[(1 + d) * p.Nh + j];
            Pj.beta = w[(1 + DIM) * p.Nh + j];
            Pj.lrho = w[(2 + DIM) * p.Nh + j];
            Pj.lbeta = w[(3 + DIM) * p.Nh + j];
            double F[NF];
            ec_flux_dot<DIM>(Pi, Pj, c, p.igm1x2, F);
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                acc[f] += F[f];
                qf[f * p.Nh + j] -= F[f];
            }
        }
        if (warp_elems) __syncwarp(); else __syncthreads();
    }
    if (act) {
#pragma unroll
        for (int f = 0; f < NF; ++f) qf[f * p.Nh + r] += acc[f];
    }
}

// Kernel A: face traces of the (projected) conservative variables, Uf = u_vfun(face rows of the projection of
// v_ufun) (dg2D_euler_quad.jl:149-150 / dg2D_euler_tri.jl:138-140).  Only the face rows are evaluated (the volume rows,
// primitive variables and logs of project_batch are not needed here), with a compact shared-memory layout
// (generic_traces_smem_bytes: U, V and one face array).
inline size_t generic_traces_smem_bytes(int DIM, int EB, int Ns, int Nq, int Nfq) {
    return (size_t)EB * (DIM + 2) * (Ns + Nq + Nfq) * sizeof(double);
}

template <int DIM>
__global__ void __launch_bounds__(ESDG_GENERIC_MAX_THREADS) k_generic_traces(const __grid_constant__ GenericParams pin) {
    constexpr int NF = DIM + 2;
    extern __shared__ double smem[];
    const GenericParams& p = pin;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* sU = smem;                          // [EB][NF][Ns]
    double* sV = sU + p.EB * NF * p.Ns;         // [EB][NF][Nq]
    double* sF = sV + p.EB * NF * p.Nq;         // [EB][NF][Nfq]  (modal: projected entropy variables at the face points)
    const long long e0 = (long long)blockIdx.x * p.EB;
    const int nb = (int)min((long long)p.EB, p.K - e0);
    bool bad = false;
    for (int f = 0; f < NF; ++f) {
        const double* src = p.Q + f * p.q_stride + e0 * p.Ns;
        for (int i = tid; i < nb * p.Ns; i += nt) {
            const int eb = i / p.Ns, k = i - eb * p.Ns;
            sU[(eb * NF + f) * p.Ns + k] = src[i];
        }
    }
    __syncthreads();
    if (!p.collocated) {
        mma_apply<false>(p.dVq, sU, p.Ns, sV, p.Nq, nb * NF);
        __syncthreads();
    }
    const double* Uq = p.collocated ? sU : sV;  // state at the quadrature points
    const int ldq = p.collocated ? p.Ns : p.Nq;
    for (int i = tid; i < nb * p.Nq; i += nt) {
        const int eb = i / p.Nq, q = i - eb * p.Nq;
        double U[NF], V[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) U[f] = Uq[(eb * NF + f) * ldq + q];
        v_of_u<DIM>(U, p.gamma, p.lg1, V, bad);
#pragma unroll
        for (int f = 0; f < NF; ++f) sV[(eb * NF + f) * p.Nq + q] = V[f];
    }
    __syncthreads();
    if (!p.collocated) {
        mma_apply<false>(p.dProjF, sV, p.Nq, sF, p.Nfq, nb * NF);
        __syncthreads();
    }
    for (int i = tid; i < nb * p.Nfq; i += nt) {
        const int eb = i / p.Nfq, k = i - eb * p.Nfq;
        double U[NF], V[NF];
        if (!p.collocated) {
#pragma unroll
            for (int f = 0; f < NF; ++f) V[f] = sF[(eb * NF + f) * p.Nfq + k];
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f) V[f] = 0.0;
            const int n = p.Proj.cnt[k];
            for (int j = 0; j < n; ++j) {
                const int c = p.Proj.col[k * p.Proj.width + j];
                const double a = p.Proj.val[k * p.Proj.width + j];
#pragma unroll
                for (int f = 0; f < NF; ++f) V[f] = fma(a, sV[(eb * NF + f) * p.Nq + c], V[f]);
            }
        }
        u_of_v<DIM>(V, p.gamma, p.lg1, U);
#pragma unroll
        for (int f = 0; f < NF; ++f) p.traces[f * p.trace_stride + e0 * p.Nfq + i] = U[f];
    }
    if (bad) atomicOr(p.bad_flag, 1);
}

// Kernel B: surface + volume + lift + (rhs | LSRK stage).
// RCAP: register cap.  Residency decides on these table-driven kernels (profiles/r1_generic_register_caps.txt):
// collocated elements run 320-thread CTAs and are best at 96 registers (quads N=3: 3.88 ms against 4.28 ms at 80 / 64);
// modal elements run 128-thread CTAs and gain from 64 registers = 8 CTAs per SM (triangles N=4: 0.93 -> 0.84 ms).
template <int DIM, int RCAP>
__global__ void __maxnreg__(RCAP) k_generic_rhs(const __grid_constant__ GenericParams pin) {
    constexpr int NF = DIM + 2;
    extern __shared__ double smem[];
    GenericParams p = pin;
    stage_generic_tables(p, smem, ST_VQ | ST_PROJ | ST_PH | ST_LF | ST_QP);   // visible after the first barrier of project_batch
    GenericSmem<DIM> s(smem, p);
    const int tid = threadIdx.x, nt = blockDim.x;
    const long long e0 = (long long)blockIdx.x * p.EB;
    const int nb = (int)min((long long)p.EB, p.K - e0);
    bool bad = false;
    // LSRK register of this thread's first output item: requested now, consumed in phase 6
    double res_pre[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) res_pre[f] = 0.0;
    if (p.mode == 1 && tid < nb * p.Ns) {
#pragma unroll
        for (int f = 0; f < NF; ++f) res_pre[f] = p.res[f * p.q_stride + e0 * p.Ns + tid];
    }
    project_batch<DIM>(p, s, e0, nb, bad);

    // phase 3: surface flux at face points
    for (int i = tid; i < nb * p.Nfq; i += nt) {
        const int eb = i / p.Nfq, k = i - eb * p.Nfq;
        const long long e = e0 + eb;
        const int face = k / p.Nfp;
        const double* g = p.geo + e * p.geo_stride + DIM * DIM + 1 + face * (DIM + 1);
        double nJ[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) nJ[d] = g[d];
        double sJ = g[DIM];
        if (DIM == 3 && p.curved) {   // normals of the face point itself (dg3D_euler_hex.jl:80-84)
            const double* gn = p.geoN + (size_t)e * 4 * p.Nfq + k;
#pragma unroll
            for (int d = 0; d < DIM; ++d) nJ[d] = gn[d * p.Nfq];
            sJ = gn[DIM * p.Nfq];
        }
        const double isJ = fast_rcp(sJ);
        // exterior state
        const long long ip = p.mapP[e * p.Nfq + k];
        double UP[NF], UM[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            UP[f] = p.traces[f * p.trace_stride + ip];
            UM[f] = s.Uf[(eb * NF + f) * p.Nfq + k];
        }
        Prim<DIM> PM, PP;
        const double* w = s.W + (size_t)eb * (NF + 2) * p.Nh + p.Nq + k;
        PM.rho = w[0];
#pragma unroll
        for (int d = 0; d < DIM; ++d) PM.vel[d] = w[(1 + d) * p.Nh];
        PM.beta = w[(1 + DIM) * p.Nh];
        PM.lrho = w[(2 + DIM) * p.Nh];
        PM.lbeta = w[(3 + DIM) * p.Nh];
        prim_of_u<DIM>(UP, p.gamma, PP);
        double F[NF];
        ec_flux_dot<DIM>(PM, PP, nJ, p.igm1x2, F);
        if (p.lf_scale != 0.0) {
            double unM = 0.0, unP = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                unM = fma(UM[1 + d], nJ[d], unM);
                unP = fma(UP[1 + d], nJ[d], unP);
            }
            const double lamM = wavespeed_1d(UM[0], unM * isJ, UM[1 + DIM], p.gamma);
            const double lamP = wavespeed_1d(UP[0], unP * isJ, UP[1 + DIM], p.gamma);
            const double LFc = p.lf_scale * fmax(lamM, lamP) * sJ;
#pragma unroll
            for (int f = 0; f < NF; ++f) F[f] -= LFc * (UP[f] - UM[f]);
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) s.Fl[(eb * NF + f) * p.Nfq + k] = F[f];
    }

    // phase 4: Hadamard sum -- symmetric schedule when the host built one, else one hybridized row per item
    if (p.sym_rounds > 0) hadamard_sym<DIM>(p, s, e0, nb);
    else
    for (int i = tid; i < nb * p.Nh; i += nt) {
        const int eb = i / p.Nh, r = i - eb * p.Nh;
        const long long e = e0 + eb;
        const double* G = p.geo + e * p.geo_stride;
        const double* w = s.W + (size_t)eb * (NF + 2) * p.Nh;
        Prim<DIM> Pi;
        Pi.rho = w[r];
#pragma unroll
        for (int d = 0; d < DIM; ++d) Pi.vel[d] = w[(1 + d) * p.Nh + r];
        Pi.beta = w[(1 + DIM) * p.Nh + r];
        Pi.lrho = w[(2 + DIM) * p.Nh + r];
        Pi.lbeta = w[(3 + DIM) * p.Nh + r];
        double acc[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[f] = 0.0;
        double Gr[DIM * DIM];
#pragma unroll
        for (int q = 0; q < DIM * DIM; ++q) Gr[q] = G[q];
        const bool curved = DIM == 3 && p.curved;
        const double* Ge = curved ? p.geoG + (size_t)e * (DIM * DIM) * p.Nh : nullptr;
        double Gi[DIM * DIM];
        if (curved) {
#pragma unroll
            for (int q = 0; q < DIM * DIM; ++q) Gi[q] = Ge[q * p.Nh + r];
        }
        const int n = p.Qp.cnt[r];
        for (int k = 0; k < n; ++k) {
            const int j = p.Qp.col[r * p.Qp.width + k];
            const double* cq = p.Qp.c + (size_t)(r * p.Qp.width + k) * 3;
            if (curved) {   // geofacs of the pair averaged: .5*(rxJ[i]+rxJ[j]) ... (dg3D_euler_hex.jl:145-151)
#pragma unroll
                for (int q = 0; q < DIM * DIM; ++q) Gr[q] = .5 * (Gi[q] + Ge[q * p.Nh + j]);
            }
            // c_d = sum_k G[d][k] * Q_k[i,j]   (Fr = rxJ*Fx + ryJ*Fy ..., :126-129)
            double c[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double a = 0.0;
#pragma unroll
                for (int kk = 0; kk < DIM; ++kk) a = fma(Gr[d * DIM + kk], cq[kk], a);
                c[d] = a;
            }
            Prim<DIM> Pj;
            Pj.rho = w[j];
#pragma unroll
            for (int d = 0; d < DIM; ++d) Pj.vel[d] = w[(1 + d) * p.Nh + j];
            Pj.beta = w[(1 + DIM) * p.Nh + j];
            Pj.lrho = w[(2 + DIM) * p.Nh + j];
            Pj.lbeta = w[(3 + DIM) * p.Nh + j];
            double F[NF];
            ec_flux_dot<DIM>(Pi, Pj, c, p.igm1x2, F);
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[f] += F[f];
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) s.QF[(eb * NF + f) * p.Nh + r] = acc[f];
    }
    __syncthreads();

    // phase 5: test with Ph, lift, scale by -1/J; phase 6: output
    double tsum = 0.0;
    if (!p.collocated) {
        // modal: R = (vs Ph) QF + LIFT flux as two accumulating batch GEMMs (dg2D_euler_tri.jl:171-175)
        mma_apply<false>(p.dPh, s.QF, p.Nh, s.R, p.Ns, nb * NF);
        __syncthreads();
        mma_apply<true>(p.dLf, s.Fl, p.Nfq, s.R, p.Ns, nb * NF);
        __syncthreads();
    }
    for (int i = tid; i < nb * p.Ns; i += nt) {
        const int eb = i / p.Ns, r = i - eb * p.Ns;
        const long long e = e0 + eb;
        double J = p.geo[e * p.geo_stride + DIM * DIM];
        if (DIM == 3 && p.curved) J = p.geoJ[(size_t)e * p.Nq + r];   // collocated: Ns == Nq (rhsQ ./ J, :212)
        const double mJ = -fast_rcp(J);
        double rr[NF];
        if (!p.collocated) {
#pragma unroll
            for (int f = 0; f < NF; ++f) rr[f] = s.R[(eb * NF + f) * p.Ns + r] * mJ;
        } else {
            double a[NF], b[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) a[f] = b[f] = 0.0;
            int n = p.Ph.cnt[r];
            for (int k = 0; k < n; ++k) {
                const int c = p.Ph.col[r * p.Ph.width + k];
                const double v = p.Ph.val[r * p.Ph.width + k];
#pragma unroll
                for (int f = 0; f < NF; ++f) a[f] = fma(v, s.QF[(eb * NF + f) * p.Nh + c], a[f]);
            }
            n = p.Lf.cnt[r];
            for (int k = 0; k < n; ++k) {
                const int c = p.Lf.col[r * p.Lf.width + k];
                const double v = p.Lf.val[r * p.Lf.width + k];
#pragma unroll
                for (int f = 0; f < NF; ++f) b[f] = fma(v, s.Fl[(eb * NF + f) * p.Nfq + c], b[f]);
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) rr[f] = (b[f] + p.volume_scale * a[f]) * mJ;
        }
        const long long g = e * p.Ns + r;
        if (p.want_rhstest) {
            if (p.collocated) {
                const double wJ = p.wq[r] * J;
#pragma unroll
                for (int f = 0; f < NF; ++f) tsum += wJ * s.V[(eb * NF + f) * p.Nq + r] * rr[f];
            } else {
#pragma unroll
                for (int f = 0; f < NF; ++f) s.R[(eb * NF + f) * p.Ns + r] = rr[f];
            }
        }
        if (p.mode == 0) {
#pragma unroll
            for (int f = 0; f < NF; ++f) p.rhs[f * p.q_stride + g] = rr[f];
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const double res = p.rk_a * (i == tid ? res_pre[f] : p.res[f * p.q_stride + g]) + p.dt * rr[f];
                p.res[f * p.q_stride + g] = res;
                p.Qout[f * p.q_stride + g] = s.U[(eb * NF + f) * p.Ns + r] + p.rk_b * res;
            }
        }
    }
    if (p.want_rhstest) {
        __syncthreads();
        if (!p.collocated) {
            // tri: sum wJq .* VU[1:Nq] .* (Vq*rhs)   (dg2D_euler_tri.jl:177-183); Vq*rhs into s.V (free by now)
            mma_apply<false>(p.dVq, s.R, p.Ns, s.V, p.Nq, nb * NF);
            __syncthreads();
            for (int i = tid; i < nb * p.Nq; i += nt) {
                const int eb = i / p.Nq, q = i - eb * p.Nq;
                const long long e = e0 + eb;
                const double J = p.geo[e * p.geo_stride + DIM * DIM];
                const double wJ = p.wq[q] * J;
#pragma unroll
                for (int f = 0; f < NF; ++f) tsum += wJ * s.VP[(eb * NF + f) * p.Nq + q] * s.V[(eb * NF + f) * p.Nq + q];
            }
        }
        // deterministic block reduction -> one partial per CTA
        __shared__ double red[32];
        for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(0xffffffffu, tsum, o);
        if ((tid & 31) == 0) red[tid >> 5] = tsum;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int wv = 0; wv < (nt + 31) / 32; ++wv) t += red[wv];
            p.rhstest_partial[blockIdx.x] = t;
        }
    }
    if (bad) atomicOr(p.bad_flag, 1);
}

// Deterministic final sum of the per-CTA partials (fixed order).
__global__ void k_reduce_partials(const double* partial, int n, double* out) {
    __shared__ double sh[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) t += partial[i];
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

}  // namespace esdg
