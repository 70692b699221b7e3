// Compressible Navier-Stokes kernels (2D): inviscid flux differencing + BR1 viscous terms with
// entropy-stable wall boundary conditions.
//
// Reference call stack (examples/CompressibleNS/dg2D_CNS_cavity_optimized.jl):
//   rhsRK!            :955-972   rhs = rhs_inviscid! + rhs_viscous!, rhstest, rhstest_visc
//   rhs_inviscid!     :447-528   entropy projection, wall mirror state (:157-176), update_flux!
//                                (:308-324), flux_differencing! (:326-347), Ph / LIFT, -1/J
//   rhs_viscous!      :749-849   VU = Pq v(Vq Q), dg_grad! (:548-569), viscous_matrices! (:613-645),
//                                sigma = K theta, dg_div! (:590-611), penalisation (:817-840),
//                                impose_BCs_entropyvars! (:178-216), impose_BCs_stress! (:218-262)
//
// The two inter-element dependences (entropy-variable traces, sigma traces) split one rhs
// evaluation into three launches; everything between them is fused per element batch:
//   k_cns_traces : v(Vq Q) -> projected entropy variables at the face points      -> V traces
//   k_cns_sigma  : BR1 gradient, K(v) theta, Pq projection -> sigma (modal) + sigma traces,
//                  rhstest_visc volume term
//   k_cns_rhs    : inviscid surface + volume, BR1 divergence, penalisation, rhstest,
//                  rhs or fused LSRK stage
// The exterior conservative state of the inviscid flux is recomputed from the neighbour's
// entropy-variable trace (u_vfun of the same numbers the neighbour itself uses), so ONE trace
// array serves QP, UP, lam[mapP] and VUP.
#pragma once
#include "kernels_generic.cuh"

namespace esdg {

struct CnsParams {
    GenericParams g;
    Ell Pq, Vf, Dr, Ds;          // Np x Nq, Nfq x Np, Np x Np, Np x Np
    Dense dPq, dVf, dDr, dDs;    // the same, dense + padded for the tensor-core GEMMs
    const unsigned char* bc;     // [Nfq*K] bits 0-2: 0 interior, 1 wall, 2 lid, 3 inflow, 4 outflow ; bit 3: member of md.mapB
    double in_prim[6];           // inflow state: rho, u, v, beta, log rho, log beta   (dg2D_CNS_modalESDG.jl:172-175)
    double in_v[4];              // its entropy variables v_ufun(rhoL, rhoL uL, rhoL vL, EL)  (:191-197)
    const double* vlid;          // [Nfq*K] lid speed (read only where kind == 2) or nullptr
    int bctype;                  // 1 adiabatic no-slip, 2 isothermal, 3 slip   (:27)
    int viscous_dissp;           // :30
    double lam, mu, Pr, Re;      // lam is ALREADY sign-flipped (`let lambda=-lambda`, :614)
    double inv_theta2;           // 2/theta of the isothermal wall (:189), theta = 1/0.3^2/1.4/0.4
    double* sig;                 // 6 planes x Np*K : Pq-projected (sigma_x[2..4], sigma_y[2..4])
    long long sig_stride;
    double* sig_tr;              // 6 planes x trace_stride : the same at the face points
    double* partial_visc;        // per CTA: sum wJq (theta . sigma)                       (:802-806)
    double* partial_vrhs;        // per CTA: sum wJq VUq (Vq * visc_rhs)                   (:966)
};

struct CnsSmem {
    // generic block first (project_batch fills it), CNS extras after it; every array is [EB][4 fields][n]
    double *Vfc;   // [Nfq]  projected entropy variables at face points (own side)
    double *A;     // [Np]   VU (modal)
    double *B1;    // [Np]   Dr*VU            | rhs kernel: G0 sigma_x + G2 sigma_y
    double *B2;    // [Np]   Ds*VU            | rhs kernel: G1 sigma_x + G3 sigma_y
    double *Tx;    // [Np]   LIFT jump -> theta_x -> sigma_x (modal)
    double *Ty;    // [Np]
    double *Jx;    // [Nfq]  jump*nxJ | sigma_x at face points | rhs kernel: viscous surface term + J * penalisation
    double *Jy;    // [Nfq]  jump*nyJ | sigma_y at face points
    double *Q1;    // [Nq]   theta_x at quadrature points -> sigma_x   | rhs kernel: Vq * visc_rhs
    double *Q2;    // [Nq]
    double *Rv;    // [Np]   viscous rhs
};

inline size_t cns_extra_doubles(int EB, int Np, int Nq, int Nfq) {
    return (size_t)EB * 4 * (3 * Nfq + 6 * Np + 2 * Nq);
}
// k_cns_sigma: U, V of the generic block + the CNS arrays; k_cns_traces: U, V + one face array
inline size_t cns_sigma_smem_bytes(int EB, int Np, int Nq, int Nfq) {
    return ((size_t)EB * 4 * (Np + Nq) + cns_extra_doubles(EB, Np, Nq, Nfq)) * sizeof(double);
}
inline size_t cns_traces_smem_bytes(int EB, int Np, int Nq, int Nfq) {
    return (size_t)EB * 4 * (Np + Nq + Nfq) * sizeof(double);
}

__device__ __forceinline__ CnsSmem cns_carve(double* base, const GenericParams& p) {
    CnsSmem c;
    const int EB = p.EB;
    c.Vfc = base;
    c.A = c.Vfc + EB * 4 * p.Nfq;
    c.B1 = c.A + EB * 4 * p.Ns;
    c.B2 = c.B1 + EB * 4 * p.Ns;
    c.Tx = c.B2 + EB * 4 * p.Ns;
    c.Ty = c.Tx + EB * 4 * p.Ns;
    c.Jx = c.Ty + EB * 4 * p.Ns;
    c.Jy = c.Jx + EB * 4 * p.Nfq;
    c.Q1 = c.Jy + EB * 4 * p.Nfq;
    c.Q2 = c.Q1 + EB * 4 * p.Nq;
    c.Rv = c.Q2 + EB * 4 * p.Nq;
    return c;
}

// y[f] = sum_k A[row,k] x[f*ld + col_k],  f < NFLD
template <int NFLD>
__device__ __forceinline__ void ell_row(const Ell& A, int row, const double* x, int ld, double* y) {
#pragma unroll
    for (int f = 0; f < NFLD; ++f) y[f] = 0.0;
    const int n = A.cnt[row];
    for (int k = 0; k < n; ++k) {
        const int c = A.col[row * A.width + k];
        const double a = A.val[row * A.width + k];
#pragma unroll
        for (int f = 0; f < NFLD; ++f) y[f] = fma(a, x[f * ld + c], y[f]);
    }
}

// impose_BCs_entropyvars! (:178-216): exterior entropy variables at a wall (kind 1) / lid (kind 2) node
__device__ __forceinline__ void bc_entropy_vars(const CnsParams& p, int kind, double vlid, double nx, double ny,
                                                const double* Vf, double* VP) {
    if (p.bctype == 1) {
        VP[1] = kind == 2 ? -Vf[1] - 2 * vlid * Vf[3] : -Vf[1];
        VP[2] = -Vf[2];
        VP[3] = Vf[3];
    } else if (p.bctype == 2) {
        VP[1] = kind == 2 ? p.inv_theta2 - Vf[1] : -Vf[1];
        VP[2] = -Vf[2];
        VP[3] = -p.inv_theta2 - Vf[3];
    } else if (p.bctype == 3) {
        const double Vn = Vf[1] * nx + Vf[2] * ny;
        VP[3] = Vf[3];
        VP[1] = Vf[1] - 2 * Vn * nx;
        VP[2] = Vf[2] - 2 * Vn * ny;
    }
}

// impose_BCs_stress! (:218-262); s*[0..2] = fields 2..4
__device__ __forceinline__ void bc_stress(const CnsParams& p, int kind, double vlid, double nx, double ny,
                                          const double* sxf, const double* syf, double* sxP, double* syP) {
    if (p.bctype == 1) {
        sxP[0] = sxf[0]; syP[0] = syf[0];
        sxP[1] = sxf[1]; syP[1] = syf[1];
        sxP[2] = kind == 2 ? -sxf[2] + 2 * vlid * sxf[0] : -sxf[2];
        syP[2] = kind == 2 ? -syf[2] + 2 * vlid * syf[0] : -syf[2];
    } else if (p.bctype == 2) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { sxP[k] = sxf[k]; syP[k] = syf[k]; }
    } else if (p.bctype == 3) {
        const double snx = sxf[0] * nx + sxf[1] * ny;
        const double sny = syf[0] * nx + syf[1] * ny;
        sxP[0] = -sxf[0] + 2 * nx * snx;
        syP[0] = -syf[0] + 2 * nx * sny;
        sxP[1] = -sxf[1] + 2 * ny * snx;
        syP[1] = -syf[1] + 2 * ny * sny;
        sxP[2] = -sxf[2];
        syP[2] = -syf[2];
    }
}

// k_cns_rhs uses Vfc, B1, B2, Jx, Q1, Rv of the CNS arrays only: compact carve behind the generic block
inline size_t cns_rhs_extra_doubles(int EB, int Np, int Nq, int Nfq) {
    return (size_t)EB * 4 * (2 * Nfq + 3 * Np + Nq);
}
__device__ __forceinline__ CnsSmem cns_carve_rhs(double* base, const GenericParams& p) {
    CnsSmem c;
    const int EB = p.EB;
    c.Vfc = base;
    c.B1 = c.Vfc + EB * 4 * p.Nfq;
    c.B2 = c.B1 + EB * 4 * p.Ns;
    c.Jx = c.B2 + EB * 4 * p.Ns;
    c.Q1 = c.Jx + EB * 4 * p.Nfq;
    c.Rv = c.Q1 + EB * 4 * p.Nq;
    c.A = c.Tx = c.Ty = c.Jy = c.Q2 = nullptr;
    return c;
}

// Stage Q, evaluate v_ufun(Vq*Q) at the quadrature points (:763-771).  Fills s.U, s.V.
__device__ __forceinline__ void cns_entropy_vars(const GenericParams& p, GenericSmem<2>& s, long long e0, int nb, bool& bad) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int f = 0; f < 4; ++f) {
        const double* src = p.Q + f * p.q_stride + e0 * p.Ns;
        for (int i = tid; i < nb * p.Ns; i += nt) {
            const int eb = i / p.Ns, k = i - eb * p.Ns;
            s.U[(eb * 4 + f) * p.Ns + k] = src[i];
        }
    }
    __syncthreads();
    mma_apply<false>(p.dVq, s.U, p.Ns, s.V, p.Nq, nb * 4);
    __syncthreads();
    for (int i = tid; i < nb * p.Nq; i += nt) {
        const int eb = i / p.Nq, q = i - eb * p.Nq;
        double U[4], V[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) U[f] = s.V[(eb * 4 + f) * p.Nq + q];
        v_of_u<2>(U, p.gamma, p.lg1, V, bad);
#pragma unroll
        for (int f = 0; f < 4; ++f) s.V[(eb * 4 + f) * p.Nq + q] = V[f];
    }
    __syncthreads();
}

// Kernel 1: entropy-variable traces  VUf = (Vh Pq)[face rows] * v(Vq Q)
__global__ void __launch_bounds__(ESDG_GENERIC_MAX_THREADS) k_cns_traces(const __grid_constant__ CnsParams Pin) {
    extern __shared__ double smem[];
    const GenericParams& p = Pin.g;
    GenericSmem<2> s(smem, p);
    s.Uf = s.W;   // compact layout (cns_traces_smem_bytes): only U, V and one face array are used here
    const long long e0 = (long long)blockIdx.x * p.EB;
    const int nb = (int)min((long long)p.EB, p.K - e0);
    bool bad = false;
    cns_entropy_vars(p, s, e0, nb, bad);
    mma_apply<false>(p.dProjF, s.V, p.Nq, s.Uf, p.Nfq, nb * 4);
    __syncthreads();
    for (int f = 0; f < 4; ++f) {
        double* dst = p.traces + f * p.trace_stride + e0 * p.Nfq;
        for (int i = threadIdx.x; i < nb * p.Nfq; i += blockDim.x) {
            const int eb = i / p.Nfq, k = i - eb * p.Nfq;
            dst[i] = s.Uf[(eb * 4 + f) * p.Nfq + k];
        }
    }
    if (bad) atomicOr(p.bad_flag, 1);
}

// Kernel 2: BR1 gradient of the projected entropy variables, sigma = K(v) theta, projection.
// Every operator application is a batch GEMM (kernels_mma.cuh); the pointwise steps between them are the
// BCs / jumps at face points, the metric combination at nodes and K(v) theta at quadrature points.
__global__ void __launch_bounds__(ESDG_GENERIC_MAX_THREADS) k_cns_sigma(const __grid_constant__ CnsParams P) {
    extern __shared__ double smem[];
    const GenericParams& p = P.g;
    GenericSmem<2> s(smem, p);
    // compact layout: this kernel touches only U and V of the generic block, so the CNS arrays start where W would
    // (cns_sigma_smem_bytes: 27 KB instead of 45 KB per 4-element CTA on quads N=3 -> 8 instead of 5 CTAs per SM)
    CnsSmem c = cns_carve(s.W, p);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int Np = p.Ns, Nq = p.Nq, Nfq = p.Nfq;
    const long long e0 = (long long)blockIdx.x * p.EB;
    const int nb = (int)min((long long)p.EB, p.K - e0);
    const int NC = nb * 4;
    bool bad = false;
    cns_entropy_vars(p, s, e0, nb, bad);
    mma_apply<false>(P.dPq, s.V, Nq, c.A, Np, NC);                 // VU = Pq * VUq  (:772)
    __syncthreads();
    mma_apply<false>(P.dVf, c.A, Np, c.Vfc, Nfq, NC);              // VUf = Vf * VU  (:775)
    mma_apply<false>(P.dDr, c.A, Np, c.B1, Np, NC);                // Dr VU, Ds VU   (:553-554)
    mma_apply<false>(P.dDs, c.A, Np, c.B2, Np, NC);
    __syncthreads();
    // VUP = VUf[mapP] + BCs ; jump terms .5*(vP - vf)*nJ  (:776-777, :558-559)
    for (int i = tid; i < nb * Nfq; i += nt) {
        const int eb = i / Nfq, k = i - eb * Nfq;
        const long long e = e0 + eb;
        const double* g = p.geo + e * p.geo_stride + 5 + (k / p.Nfp) * 3;
        const double nxJ = g[0], nyJ = g[1], sJ = g[2];
        double Vf[4], VP[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) Vf[f] = c.Vfc[(eb * 4 + f) * Nfq + k];
        const long long self = e * Nfq + k;
        const long long ip = p.mapP[self];
        if (ip == self) {
#pragma unroll
            for (int f = 0; f < 4; ++f) VP[f] = Vf[f];
        } else {
#pragma unroll
            for (int f = 0; f < 4; ++f) VP[f] = p.traces[f * p.trace_stride + ip];
        }
        const int code = P.bc[self];
        const int kind = code & 7;
        if (kind == 1 || kind == 2) {
            const double isJ = fast_rcp(sJ);
            bc_entropy_vars(P, kind, kind == 2 ? P.vlid[self] : 0.0, nxJ * isJ, nyJ * isJ, Vf, VP);
        } else if (kind == 3) {            // inflow: entropy variables of the prescribed state (dg2D_CNS_modalESDG.jl:191-197)
#pragma unroll
            for (int f = 0; f < 4; ++f) VP[f] = P.in_v[f];
        } else if (kind == 4) {            // outflow: exterior = interior (:199-202)
#pragma unroll
            for (int f = 0; f < 4; ++f) VP[f] = Vf[f];
        }
        c.Jx[(eb * 4) * Nfq + k] = 0.0;
        c.Jy[(eb * 4) * Nfq + k] = 0.0;
#pragma unroll
        for (int f = 1; f < 4; ++f) {
            const double h = .5 * (VP[f] - Vf[f]);
            c.Jx[(eb * 4 + f) * Nfq + k] = h * nxJ;
            c.Jy[(eb * 4 + f) * Nfq + k] = h * nyJ;
        }
    }
    __syncthreads();
    mma_apply<false>(p.dLf, c.Jx, Nfq, c.Tx, Np, NC);              // LIFT * jump
    mma_apply<false>(p.dLf, c.Jy, Nfq, c.Ty, Np, NC);
    __syncthreads();
    // dg_grad! (:548-569): theta = (G [Dr;Ds] v + LIFT jump)/J
    for (int i = tid; i < nb * Np; i += nt) {
        const int eb = i / Np, r = i - eb * Np;
        const double* G = p.geo + (e0 + eb) * p.geo_stride;   // rxJ, sxJ, ryJ, syJ, J
        const double G0 = G[0], G1 = G[1], G2 = G[2], G3 = G[3], iJ = fast_rcp(G[4]);
#pragma unroll
        for (int f = 1; f < 4; ++f) {
            const int o = (eb * 4 + f) * Np + r;
            const double vr = c.B1[o], vs = c.B2[o];
            c.Tx[o] = (G0 * vr + G1 * vs + c.Tx[o]) * iJ;
            c.Ty[o] = (G2 * vr + G3 * vs + c.Ty[o]) * iJ;
        }
    }
    __syncthreads();
    mma_apply<false>(p.dVq, c.A, Np, s.V, Nq, NC);                 // VUq, VUx, VUy at the quadrature points (:780-782)
    mma_apply<false>(p.dVq, c.Tx, Np, c.Q1, Nq, NC);
    mma_apply<false>(p.dVq, c.Ty, Np, c.Q2, Nq, NC);
    __syncthreads();
    // quadrature points: sigma = K(v) theta (:786-800), rhstest_visc volume term (:802-806)
    double tsum = 0.0;
    for (int i = tid; i < nb * Nq; i += nt) {
        const int eb = i / Nq, q = i - eb * Nq;
        const long long e = e0 + eb;
        double tx[3], ty[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            tx[f] = c.Q1[(eb * 4 + 1 + f) * Nq + q];
            ty[f] = c.Q2[(eb * 4 + 1 + f) * Nq + q];
        }
        const double v2 = s.V[(eb * 4 + 1) * Nq + q], v3 = s.V[(eb * 4 + 2) * Nq + q], v4 = s.V[(eb * 4 + 3) * Nq + q];
        const double lam = P.lam, mu = P.mu, l2m = lam + 2.0 * mu;
        const double ic = fast_rcp(v4 * v4 * v4);
        const double gmp = p.gamma * mu / P.Pr;
        const double v44 = v4 * v4;
        // Kxx, Kxy, Kyy rows/cols 2..4 (indices 0..2 here), :613-645
        const double xx00 = ic * -l2m * v44, xx02 = ic * l2m * v2 * v4, xx11 = ic * -mu * v44, xx12 = ic * mu * v3 * v4;
        const double xx22 = ic * -(l2m * v2 * v2 + mu * v3 * v3 - gmp * v4);
        const double xy01 = ic * -lam * v44, xy02 = ic * lam * v3 * v4, xy10 = ic * -mu * v44, xy12 = ic * mu * v2 * v4;
        const double xy20 = ic * mu * v3 * v4, xy21 = ic * lam * v2 * v4, xy22 = ic * (lam + mu) * (-v2 * v3);
        const double yy00 = ic * -mu * v44, yy02 = ic * mu * v2 * v4, yy11 = ic * -l2m * v44, yy12 = ic * l2m * v3 * v4;
        const double yy22 = ic * -(l2m * v3 * v3 + mu * v2 * v2 - gmp * v4);
        double sx[3], sy[3];
        // sigma_x[row] = sum_col Kxx[row,col] tx[col] + Kxy[row,col] ty[col]
        sx[0] = xx00 * tx[0] + xy01 * ty[1] + (xx02 * tx[2] + xy02 * ty[2]);
        sx[1] = xy10 * ty[0] + xx11 * tx[1] + (xx12 * tx[2] + xy12 * ty[2]);
        sx[2] = (xx02 * tx[0] + xy20 * ty[0]) + (xx12 * tx[1] + xy21 * ty[1]) + (xx22 * tx[2] + xy22 * ty[2]);
        // sigma_y[row] = sum_col Kxy[col,row] tx[col] + Kyy[row,col] ty[col]
        sy[0] = yy00 * ty[0] + xy10 * tx[1] + (xy20 * tx[2] + yy02 * ty[2]);
        sy[1] = xy01 * tx[0] + yy11 * ty[1] + (xy21 * tx[2] + yy12 * ty[2]);
        sy[2] = (xy02 * tx[0] + yy02 * ty[0]) + (xy12 * tx[1] + yy12 * ty[1]) + (xy22 * tx[2] + yy22 * ty[2]);
        const double wJ = p.wq[q] * p.geo[e * p.geo_stride + 4];
        c.Q1[(eb * 4) * Nq + q] = 0.0;
        c.Q2[(eb * 4) * Nq + q] = 0.0;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            tsum += wJ * tx[f] * sx[f];
            tsum += wJ * ty[f] * sy[f];
            c.Q1[(eb * 4 + 1 + f) * Nq + q] = sx[f];
            c.Q2[(eb * 4 + 1 + f) * Nq + q] = sy[f];
        }
    }
    __syncthreads();
    mma_apply<false>(P.dPq, c.Q1, Nq, c.Tx, Np, NC);               // sigma = Pq * sigma (:808-809)
    mma_apply<false>(P.dPq, c.Q2, Nq, c.Ty, Np, NC);
    __syncthreads();
    for (int i = tid; i < nb * Np; i += nt) {                      // modal sigma -> HBM (read again by k_cns_rhs)
        const int eb = i / Np, r = i - eb * Np;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            P.sig[f * P.sig_stride + e0 * Np + i] = c.Tx[(eb * 4 + 1 + f) * Np + r];
            P.sig[(3 + f) * P.sig_stride + e0 * Np + i] = c.Ty[(eb * 4 + 1 + f) * Np + r];
        }
    }
    mma_apply<false>(P.dVf, c.Tx, Np, c.Jx, Nfq, NC);              // sigma_f = Vf * sigma (:811-812) -> sigma traces
    mma_apply<false>(P.dVf, c.Ty, Np, c.Jy, Nfq, NC);
    __syncthreads();
    for (int i = tid; i < nb * Nfq; i += nt) {
        const int eb = i / Nfq, k = i - eb * Nfq;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            P.sig_tr[f * p.trace_stride + e0 * Nfq + i] = c.Jx[(eb * 4 + 1 + f) * Nfq + k];
            P.sig_tr[(3 + f) * p.trace_stride + e0 * Nfq + i] = c.Jy[(eb * 4 + 1 + f) * Nfq + k];
        }
    }
    // deterministic per-CTA partial of the viscous entropy term
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(0xffffffffu, tsum, o);
    if ((tid & 31) == 0) red[tid >> 5] = tsum;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int wv = 0; wv < (nt + 31) / 32; ++wv) t += red[wv];
        P.partial_visc[blockIdx.x] = t;
    }
    if (bad) atomicOr(p.bad_flag, 1);
}

// Kernel 3: inviscid rhs + BR1 divergence + penalisation -> rhs or LSRK stage.
// Residency is the lever on this path (profiles/r1_cns_register_caps.txt, quad N=3, 128-thread CTAs, compact shared
// memory of 35 KB): 134 registers = 3 CTAs per SM 2.8 ms, 128 = 4 CTAs 2.12 ms, 96 = 5 CTAs 1.91 ms, 80 = 6 CTAs 1.76 ms
// per launch (20 bytes of spills)
#ifndef ESDG_CNS_RHS_REGS
#define ESDG_CNS_RHS_REGS 80
#endif
__global__ void __maxnreg__(ESDG_CNS_RHS_REGS) k_cns_rhs(const __grid_constant__ CnsParams P) {
    const GenericParams& p = P.g;
    constexpr int NF = 4;
    extern __shared__ double smem[];
    GenericSmem<2> s(smem, p);
    CnsSmem c = cns_carve_rhs(s.R + p.EB * 4 * p.Ns, p);
    s.Vface = c.Vfc;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int Np = p.Ns, Nq = p.Nq, Nfq = p.Nfq;
    const long long e0 = (long long)blockIdx.x * p.EB;
    const int nb = (int)min((long long)p.EB, p.K - e0);
    const int NC = nb * 4;
    bool bad = false;
    project_batch<2>(p, s, e0, nb, bad);
    // sigma (modal) of this batch, combined with the metric terms so that dg_div! (:590-611) becomes
    //   Dr (rxJ sx + ryJ sy) + Ds (sxJ sx + syJ sy):  B1 = G0 sx + G2 sy, B2 = G1 sx + G3 sy
    for (int i = tid; i < nb * Np; i += nt) {
        const int eb = i / Np, r = i - eb * Np;
        const double* G = p.geo + (e0 + eb) * p.geo_stride;
        const double G0 = G[0], G1 = G[1], G2 = G[2], G3 = G[3];
        c.B1[(eb * 4) * Np + r] = 0.0;
        c.B2[(eb * 4) * Np + r] = 0.0;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const double sx = P.sig[f * P.sig_stride + e0 * Np + i], sy = P.sig[(3 + f) * P.sig_stride + e0 * Np + i];
            c.B1[(eb * 4 + 1 + f) * Np + r] = G0 * sx + G2 * sy;
            c.B2[(eb * 4 + 1 + f) * Np + r] = G1 * sx + G3 * sy;
        }
    }
    // surface terms at the face points
    for (int i = tid; i < nb * Nfq; i += nt) {
        const int eb = i / Nfq, k = i - eb * Nfq;
        const long long e = e0 + eb;
        const double* g = p.geo + e * p.geo_stride + 5 + (k / p.Nfp) * 3;
        const double nJ[2] = {g[0], g[1]};
        const double sJ = g[2], isJ = fast_rcp(sJ);
        const double nx = nJ[0] * isJ, ny = nJ[1] * isJ;
        const long long self = e * Nfq + k;
        const long long ip = p.mapP[self];
        const int code = P.bc[self];
        const int kind = code & 7;
        const double vlid = kind == 2 ? P.vlid[self] : 0.0;
        double UM[NF], UP[NF], VM[NF], VPv[NF];
        Prim<2> PM, PP;
        const double* w = s.W + (size_t)eb * (NF + 2) * p.Nh + Nq + k;
        PM.rho = w[0]; PM.vel[0] = w[p.Nh]; PM.vel[1] = w[2 * p.Nh];
        PM.beta = w[3 * p.Nh]; PM.lrho = w[4 * p.Nh]; PM.lbeta = w[5 * p.Nh];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            UM[f] = s.Uf[(eb * NF + f) * Nfq + k];
            VM[f] = c.Vfc[(eb * NF + f) * Nfq + k];
        }
        if (ip == self) {
#pragma unroll
            for (int f = 0; f < NF; ++f) { UP[f] = UM[f]; VPv[f] = VM[f]; }
            PP = PM;
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f) VPv[f] = p.traces[f * p.trace_stride + ip];
            u_of_v<2>(VPv, p.gamma, p.lg1, UP);
            prim_of_u<2>(UP, p.gamma, PP);
        }
        if (kind == 1 || kind == 2) {
            // impose_BCs_inviscid! (:157-176): rho+ = rho, beta+ = beta, u+ = u - 2 (u.n) n
            const double Un = PM.vel[0] * nx + PM.vel[1] * ny;
            PP.rho = PM.rho; PP.beta = PM.beta; PP.lrho = PM.lrho; PP.lbeta = PM.lbeta;
            PP.vel[0] = PM.vel[0] - 2 * Un * nx;
            PP.vel[1] = PM.vel[1] - 2 * Un * ny;
        } else if (kind == 3) {            // shock-tube inflow: prescribed primitive state (dg2D_CNS_modalESDG.jl:172-175)
            PP.rho = P.in_prim[0]; PP.vel[0] = P.in_prim[1]; PP.vel[1] = P.in_prim[2];
            PP.beta = P.in_prim[3]; PP.lrho = P.in_prim[4]; PP.lbeta = P.in_prim[5];
        } else if (kind == 4) {            // outflow: exterior = interior (:177-180)
            PP = PM;
        }
        double F[NF];
        ec_flux_dot<2>(PP, PM, nJ, p.igm1x2, F);          // update_flux! evaluates (QP, QM) (:311)
        if (p.lf_scale != 0.0 && kind < 3) {   // impose_BCs_lam! zeroes the penalty at inflow / outflow nodes (:183-188)
            const double unM = (UM[1] * nJ[0] + UM[2] * nJ[1]) * isJ;
            const double unP = (UP[1] * nJ[0] + UP[2] * nJ[1]) * isJ;
            const double lamM = wavespeed_1d(UM[0], unM, UM[3], p.gamma);
            const double lamP = wavespeed_1d(UP[0], unP, UP[3], p.gamma);
            const double LFc = p.lf_scale * fmax(lamM, lamP) * sJ;
#pragma unroll
            for (int f = 0; f < NF; ++f) F[f] -= LFc * (UP[f] - UM[f]);
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) s.Fl[(eb * NF + f) * Nfq + k] = F[f];
        // viscous surface term .5*((sxP-sxf)nxJ + (syP-syf)nyJ)  (:604) with stress BCs
        double sxf[3], syf[3], sxP[3], syP[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            sxf[f] = P.sig_tr[f * p.trace_stride + self];
            syf[f] = P.sig_tr[(3 + f) * p.trace_stride + self];
            sxP[f] = P.sig_tr[f * p.trace_stride + ip];
            syP[f] = P.sig_tr[(3 + f) * p.trace_stride + ip];
        }
        if (kind == 1 || kind == 2) bc_stress(P, kind, vlid, nx, ny, sxf, syf, sxP, syP);
        else if (kind >= 3) {              // sigma+ = sigma on `vwall` (dg2D_CNS_modalESDG.jl:205-215)
#pragma unroll
            for (int f = 0; f < 3; ++f) { sxP[f] = sxf[f]; syP[f] = syf[f]; }
        }
        // viscous face term of this point: surf + J * pen, so that (vol + LIFT surf)/J + LIFT pen (:843-846) is one lift
        double sv[3], pen[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int f = 0; f < 3; ++f) sv[f] = .5 * ((sxP[f] - sxf[f]) * nJ[0] + (syP[f] - syf[f]) * nJ[1]);
        // penalisation tau*[[v]] (:817-840)
        if (P.viscous_dissp) {
            if (kind == 1 || kind == 2) bc_entropy_vars(P, kind, vlid, nx, ny, VM, VPv);
            else if (kind == 3) {
#pragma unroll
                for (int f = 0; f < 4; ++f) VPv[f] = P.in_v[f];
            } else if (kind == 4) {
#pragma unroll
                for (int f = 0; f < 4; ++f) VPv[f] = VM[f];
            }
            const double iv4 = fast_rcp(VM[3]);
            const double tau = -iv4 / P.Re;
            const double d1 = VPv[1] - VM[1], d2 = VPv[2] - VM[2], d3 = VPv[3] - VM[3];
            double p3 = tau * d3;
            if (code & 8) {
                const double a1 = .5 * (VPv[1] + VM[1]), a2 = .5 * (VPv[2] + VM[2]);
                p3 = P.bctype == 1 ? -tau * (a1 * d1 + a2 * d2) * iv4
                                   : -tau * (a1 * d1 + a2 * d2 + d3 * d3 * .5) * iv4;
            }
            pen[0] = tau * d1;
            pen[1] = tau * d2;
            pen[2] = p3;
        }
        const double Je = p.geo[e * p.geo_stride + 4];
        c.Jx[(eb * 4) * Nfq + k] = 0.0;
#pragma unroll
        for (int f = 0; f < 3; ++f) c.Jx[(eb * 4 + 1 + f) * Nfq + k] = fma(Je, pen[f], sv[f]);
    }
    // flux differencing (:326-347): symmetric schedule (kernels_generic.cuh: hadamard_sym) or one hybridized row per item
    if (p.sym_rounds > 0) hadamard_sym<2>(p, s, e0, nb);
    else
    for (int i = tid; i < nb * p.Nh; i += nt) {
        const int eb = i / p.Nh, r = i - eb * p.Nh;
        const long long e = e0 + eb;
        const double* G = p.geo + e * p.geo_stride;
        const double* w = s.W + (size_t)eb * (NF + 2) * p.Nh;
        Prim<2> Pi;
        Pi.rho = w[r]; Pi.vel[0] = w[p.Nh + r]; Pi.vel[1] = w[2 * p.Nh + r];
        Pi.beta = w[3 * p.Nh + r]; Pi.lrho = w[4 * p.Nh + r]; Pi.lbeta = w[5 * p.Nh + r];
        double acc[NF] = {0.0, 0.0, 0.0, 0.0};
        const double G0 = G[0], G1 = G[1], G2 = G[2], G3 = G[3];
        const int n = p.Qp.cnt[r];
        for (int k = 0; k < n; ++k) {
            const int j = p.Qp.col[r * p.Qp.width + k];
            const double* cq = p.Qp.c + (size_t)(r * p.Qp.width + k) * 3;
            double cc[2];
            cc[0] = fma(G1, cq[1], G0 * cq[0]);
            cc[1] = fma(G3, cq[1], G2 * cq[0]);
            Prim<2> Pj;
            Pj.rho = w[j]; Pj.vel[0] = w[p.Nh + j]; Pj.vel[1] = w[2 * p.Nh + j];
            Pj.beta = w[3 * p.Nh + j]; Pj.lrho = w[4 * p.Nh + j]; Pj.lbeta = w[5 * p.Nh + j];
            double F[NF];
            ec_flux_dot<2>(Pi, Pj, cc, p.igm1x2, F);
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[f] += F[f];
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) s.QF[(eb * NF + f) * p.Nh + r] = acc[f];
    }
    __syncthreads();
    // assemble with batch GEMMs: inviscid numerator R = (vs Ph) QF + LIFT flux (:514-517); viscous numerator
    // Rv = Dr B1 + Ds B2 + LIFT (surf + J pen)  [= J * ((G D sigma + LIFT surf)/J + LIFT pen), :590-611, :843-846]
    mma_apply<false>(p.dPh, s.QF, p.Nh, s.R, Np, NC);
    mma_apply<true>(p.dLf, s.Fl, Nfq, s.R, Np, NC);
    mma_apply<false>(P.dDr, c.B1, Np, c.Rv, Np, NC);
    mma_apply<true>(P.dDs, c.B2, Np, c.Rv, Np, NC);
    mma_apply<true>(p.dLf, c.Jx, Nfq, c.Rv, Np, NC);
    __syncthreads();
    for (int i = tid; i < nb * Np; i += nt) {
        const int eb = i / Np, r = i - eb * Np;
        const long long e = e0 + eb;
        const double iJ = fast_rcp(p.geo[e * p.geo_stride + 4]);
        double rr[NF], rv[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            rv[f] = c.Rv[(eb * NF + f) * Np + r] * iJ;
            rr[f] = rv[f] - s.R[(eb * NF + f) * Np + r] * iJ;
        }
        const long long gidx = e * Np + r;
        if (p.want_rhstest) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                s.R[(eb * NF + f) * Np + r] = rr[f];
                c.Rv[(eb * NF + f) * Np + r] = rv[f];
            }
        }
        if (p.mode == 0) {
#pragma unroll
            for (int f = 0; f < NF; ++f) p.rhs[f * p.q_stride + gidx] = rr[f];
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const double res = p.rk_a * p.res[f * p.q_stride + gidx] + p.dt * rr[f];
                p.res[f * p.q_stride + gidx] = res;
                p.Qout[f * p.q_stride + gidx] = s.U[(eb * NF + f) * Np + r] + p.rk_b * res;
            }
        }
    }
    if (p.want_rhstest) {
        __syncthreads();
        // rhstest = sum wJq VUq (Vq*rhs), rhstest_visc = sum wJq VUq (Vq*visc_rhs) + visc_test (:960-971)
        mma_apply<false>(p.dVq, s.R, Np, s.V, Nq, NC);
        mma_apply<false>(p.dVq, c.Rv, Np, c.Q1, Nq, NC);
        __syncthreads();
        double ta = 0.0, tv = 0.0;
        for (int i = tid; i < nb * Nq; i += nt) {
            const int eb = i / Nq, q = i - eb * Nq;
            const long long e = e0 + eb;
            const double wJ = p.wq[q] * p.geo[e * p.geo_stride + 4];
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const double vq = s.VP[(eb * NF + f) * Nq + q];
                ta += wJ * vq * s.V[(eb * NF + f) * Nq + q];
                tv += wJ * vq * c.Q1[(eb * NF + f) * Nq + q];
            }
        }
        __shared__ double red[2][32];
        for (int o = 16; o > 0; o >>= 1) {
            ta += __shfl_down_sync(0xffffffffu, ta, o);
            tv += __shfl_down_sync(0xffffffffu, tv, o);
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = ta; red[1][tid >> 5] = tv; }
        __syncthreads();
        if (tid == 0) {
            double t0 = 0.0, t1 = 0.0;
            for (int wv = 0; wv < (nt + 31) / 32; ++wv) { t0 += red[0][wv]; t1 += red[1][wv]; }
            p.rhstest_partial[blockIdx.x] = t0;
            P.partial_vrhs[blockIdx.x] = t1;
        }
    }
    if (bad) atomicOr(p.bad_flag, 1);
}

// rhstest[0] = sum a ; rhstest[1] = sum b + sum c   (fixed order: deterministic)
__global__ void k_reduce_partials_cns(const double* a, const double* b, const double* c, int n, double* out) {
    __shared__ double sh[2][256];
    double t0 = 0.0, t1 = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { t0 += a[i]; t1 += b[i] + c[i]; }
    sh[0][threadIdx.x] = t0;
    sh[1][threadIdx.x] = t1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sh[0][0]; out[1] = sh[1][0]; }
}

}  // namespace esdg
