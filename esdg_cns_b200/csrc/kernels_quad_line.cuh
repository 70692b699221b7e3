// Tensor-product QUAD path (collocated Gauss nodes, affine elements; BASELINE config 1): the line decomposition of the
// hex kernels (hex_common.cuh) in two dimensions.  The hybridized SBP operators of a Gauss-collocated quad
// (dg2D_euler_quad.jl:47-78) couple a volume node only to the nodes of its two grid lines and to the two face points each
// line ends in, so `sparse_hadamard_sum` (:102-138) splits into 2n independent LINE problems per element:
// n volume nodes + 2 face nodes -> n(n-1)/2 + 2n unordered pairs.  One thread owns one line (nodes and accumulators in
// registers, every unordered pair once, surface flux at its own two face points, lift onto its own nodes); one launch
// per LSRK stage does everything `rhs` + the stage update do (:141-194, :198-207):
//   A  per volume node : (rho, u/2, v/2, beta) from Q, (log rho, log beta) from the stash          -> shared WV[6]
//   C  per line        : pair fluxes, F_EC(M,P).nJ - LFc (U+ - U), face rows, lift                  -> shared A2[2][4]
//   E  per volume node : rhs = -(A2_r + A2_s)/J ; rhstest ; LSRK stage ; v_ufun of the new state (shared V[4]) ; stash
//   F  per line        : entropy projection of the new state to the line's two face points -> 7-plane traces
// Trace planes: rho, u/2, v/2, beta, log rho, log beta, lambda (the owner's wave speed = the script's lam[mapP]).
// Node k = i + n j with s = x[i] (fastest), r = x[j]; faces counter-clockwise s=-1, r=+1, s=+1, r=-1, the last two with
// reversed point order (esdg_cns_b200/setup_dg.py:reference_quad).  esdg_finalize proves that the caller's operators ARE
// these Kronecker products before selecting the path; otherwise the generic table-driven kernels run.
#pragma once
#include "hex_common.cuh"

namespace esdg {

constexpr int QUAD_TP = 7;     // trace planes
constexpr int QUAD_GEO = 17;   // G[4] = rxJ, sxJ, ryJ, syJ ; J ; 4 x (nxJ, nyJ, sJ)

struct QuadParams {
    long long K, e_begin, e_count;
    long long q_stride, t_stride;
    const double* Q;
    double* Qout;
    double* res;
    double* rhs;
    const double* Lcur;      // stashed (log rho, log beta): 2 planes of q_stride
    double* Lnext;
    const double* Tcur;
    double* Tnext;
    const int* mapP;
    const double* geo;
    double gamma, lf_scale, lg1, l2g1;
    double rk_a, rk_b, dt;
    int mode, want_rhstest;
    double* partial;
    int* bad;
    double T[64], lm[8], lp[8], w[8], lmw[8], lpw[8];
    // compressible Navier-Stokes extension (kernels_cnsq.cuh; k_quad_stage_line<n, true>)
    double* sig;             // 6 planes x q_stride: sigma_x[2..4], sigma_y[2..4] at the volume nodes
    double* sig_tr;          // 6 planes x t_stride: the same at the face points
    const unsigned char* bc; // [Nfq*K] bits 0-2: 0 interior, 1 wall, 2 lid; bit 3: member of md.mapB
    const double* vlid;      // [Nfq*K] lid speed (read where kind == 2)
    int bctype, viscous_dissp;
    double lam, mu, Pr, Re, inv_theta2;   // lam already sign-flipped (dg2D_CNS_cavity_optimized.jl:614)
    double D[64];            // 1-D differentiation matrix at the Gauss nodes
    double* partial_visc;    // per CTA: sum wJq (theta . sigma)            (:802-806)
    double* partial_vrhs;    // per CTA: sum wJq VUq . visc_rhs             (:966)
};

#ifndef ESDG_QUAD_PF_AHEAD
#define ESDG_QUAD_PF_AHEAD 222   // half a wave of CTAs ahead (3 x 148 resident): measured 0 / 222 / 444 / 888 -> 1.287 / 1.236 / 1.243 / 1.294 ms
#endif
#ifndef ESDG_QUAD_THREADS
#define ESDG_QUAD_THREADS 128
#endif
#ifndef ESDG_QUAD_OWN_SMEM
#define ESDG_QUAD_OWN_SMEM 1   // Euler: own traces staged in shared memory by cp.async (1.28 ms per stage on quads N=3 1024^2) instead of being
#endif                         // read by the line thread (1.37 ms); profiles/r2_quad_variants.txt



template <int n>
struct QuadDims {
    static constexpr int Nq = n * n, Nfq = 4 * n, NL = 2 * n;
    static constexpr int THREADS = ESDG_QUAD_THREADS;
    static constexpr int EB = THREADS / NL;                   // n = 4: 16 elements, 60 KB of shared memory, 3 CTAs per SM
    static constexpr int MIN_CTAS = 384 / THREADS;            // 12 warps per SM at 168 registers
    static constexpr int EN = EB * Nq, EF = EB * Nfq;
    // WV[6][EN] WF[7][EF] P[7][EF] A2[2][4][EN] (V[4][EN] aliases A2 block 0) geo[EB][17] lm lp w
    static constexpr int SMEM_DOUBLES = 6 * EN + 2 * QUAD_TP * EF + 8 * EN + EB * QUAD_GEO + 3 * n;
    // CNS: own traces come straight from global memory; instead sigma at the nodes S[6][EN] and the exterior sigma traces SP[6][EF]
    static constexpr int SMEM_DOUBLES_CNS = 6 * EN + QUAD_TP * EF + 8 * EN + 6 * EN + 6 * EF + EB * QUAD_GEO + 3 * n;
};

template <int W>
struct Node2V {
    VD<W> rho, u[2], beta, lr, lb;
};
struct Node2 {
    double rho, u[2], beta, lr, lb;
};
template <int W>
__device__ __forceinline__ void set_lane(Node2V<W>& d, int q, const Node2& s) {
    d.rho[q] = s.rho; d.u[0][q] = s.u[0]; d.u[1][q] = s.u[1]; d.beta[q] = s.beta; d.lr[q] = s.lr; d.lb[q] = s.lb;
}
__device__ __forceinline__ Node2 ld_node2(const double* Wp, int S, int idx) {
    Node2 a;
    a.rho = Wp[idx]; a.u[0] = Wp[S + idx]; a.u[1] = Wp[2 * S + idx]; a.beta = Wp[3 * S + idx]; a.lr = Wp[4 * S + idx]; a.lb = Wp[5 * S + idx];
    return a;
}

// euler_fluxes.jl:23-48 with logmean.jl:14-28, contracted with g = (g_x, g_y); see pair_flux_v (hex_common.cuh) for the
// arithmetic choices (sums instead of means, one reciprocal per pair, branch-free Taylor branch, half velocities)
template <int W>
__device__ __forceinline__ void pair_flux2_v(const Node2V<W>& A, const Node2V<W>& B, const VD<W>* g, double c_igm, VD<W>* F) {
    const VD<W> dar = B.rho - A.rho, sr = A.rho + B.rho;
    const VD<W> dab = B.beta - A.beta, sb = A.beta + B.beta;
    const VB<W> tr = vlt(vabs(dar), 5e-5 * sr), tb = vlt(vabs(dab), 5e-5 * sb);
    const VD<W> X = vsel(tr, sr, B.lr - A.lr);
    const VD<W> Y = vsel(tb, sb, dab);
    const VD<W> XY = X * Y;
    const VD<W> R = vrcp(XY * sb);
    const VD<W> iX = R * (Y * sb), iY = R * (X * sb), iZ = R * XY;
    const VD<W> qr = dar * iX, fb = dab * iY;
    const VD<W> vr = qr * qr, vb = fb * fb;
    const VD<W> rholog = vsel(tr, sr * vfma(vr, -.4, .5), qr);
    const VD<W> ibl = iY * vsel(tb, vfma(vb, 1.6, 2.0), B.lb - A.lb);
    const VD<W> pa = (.5 * sr) * iZ;
    const VD<W> ua0 = A.u[0] + B.u[0], ua1 = A.u[1] + B.u[1];
    const VD<W> unorm = vfma(A.u[1], B.u[1], A.u[0] * B.u[0]);
    const VD<W> Ep = vfma(rholog, vfma(unorm, 2.0, ibl * c_igm), pa);
    const VD<W> un = vfma(g[1], ua1, g[0] * ua0);
    const VD<W> m = rholog * un;
    F[0] = m;
    F[1] = vfma(m, ua0, pa * g[0]);
    F[2] = vfma(m, ua1, pa * g[1]);
    F[3] = Ep * un;
}

template <int W>
__device__ __forceinline__ void vol_prim2_v(const VD<W>* U, double gamma, VD<W>* hu, VD<W>& beta, VD<W>& rhoe) {
    const VD<W> hir = .5 * vrcp(U[0]);
    hu[0] = U[1] * hir;
    hu[1] = U[2] * hir;
    const VD<W> ke = vfma(U[2], hu[1], U[1] * hu[0]);
    rhoe = U[3] - ke;
    beta = .5 * U[0] * vrcp((gamma - 1.0) * rhoe);
}
template <int W>
__device__ __forceinline__ void entropy_vars2_v(const VD<W>* U, const VD<W>& rhoe, const VD<W>& lr, const VD<W>& lb, double gamma, VD<W>* V) {
    const VD<W> s = vfma(lr, 1.0 - gamma, -lb) - 0.6931471805599453094;
    const VD<W> ire = vrcp(rhoe);
    V[0] = vfma(-U[3], ire, (gamma + 1.0) - s);
    V[1] = U[1] * ire;
    V[2] = U[2] * ire;
    V[3] = -(U[0] * ire);
}
// u_vfun + betafun + logs + wave speed from the projected entropy variables (see trace_from_entropy_v, hex_common.cuh)
template <int W>
__device__ __forceinline__ void trace_from_entropy2_v(const VD<W>* V, const VD<W>* nh, double gamma, double lg1, double l2g1, VD<W>* Tv) {
    const VD<W> nvE = -V[3];
    const VD<W> inv = vrcp(nvE);
    const VD<W> v2 = vfma(V[1], V[1], V[2] * V[2]);
    const VD<W> s = (gamma - V[0]) - .5 * v2 * inv;
    const VD<W> lv = vlog_pos(nvE);
    const double ig = 1.0 / (gamma - 1.0);
    const VD<W> lrhoe = ((lg1 - gamma * lv) - s) * ig;
    const VD<W> rhoeV = vexp_mod(lrhoe);
    Tv[0] = rhoeV * nvE;
    const VD<W> hinv = .5 * inv;
    Tv[1] = V[1] * hinv;
    Tv[2] = V[2] * hinv;
    Tv[3] = (.5 * ig) * nvE;
    Tv[4] = vlog_pos(Tv[0]);
    Tv[5] = lv - l2g1;
    const VD<W> un = vfma(Tv[2], nh[1], Tv[1] * nh[0]);
    const VD<W> hq = vfma(Tv[1], Tv[1], Tv[2] * Tv[2]);
    const VD<W> c2 = (gamma * (gamma - 1.0)) * (inv + vfma(un * un, -.5, 2.0 * hq));
    Tv[6] = vsqrt_fast(vabs(un)) + vsqrt_fast(c2);
}

__device__ __forceinline__ Node2 ld_node2_g(const double* Tp, long long S, long long idx) {
    Node2 a;
    a.rho = Tp[idx]; a.u[0] = Tp[S + idx]; a.u[1] = Tp[2 * S + idx]; a.beta = Tp[3 * S + idx]; a.lr = Tp[4 * S + idx]; a.lb = Tp[5 * S + idx];
    return a;
}
// entropy variables 2..4 from the primitive trace (u/2, v/2, beta): v4 = -rho/rhoe = -2 (gamma-1) beta, v2 = -v4 u, v3 = -v4 v
__device__ __forceinline__ void v234_from_prim(double hu, double hv, double beta, double gamma, double* V) {
    const double c = 2.0 * (gamma - 1.0) * beta;
    V[1] = 2.0 * c * hu;
    V[2] = 2.0 * c * hv;
    V[3] = -c;
}
// impose_BCs_entropyvars! (dg2D_CNS_cavity_optimized.jl:178-216): exterior entropy variables at a wall (kind 1) / lid (kind 2) node
__device__ __forceinline__ void bc_entropy_vars_q(int bctype, int kind, double vlid, double inv_theta2, double nx, double ny, const double* Vf,
                                                  double* VP) {
    if (bctype == 1) {
        VP[1] = kind == 2 ? -Vf[1] - 2 * vlid * Vf[3] : -Vf[1];
        VP[2] = -Vf[2];
        VP[3] = Vf[3];
    } else if (bctype == 2) {
        VP[1] = kind == 2 ? inv_theta2 - Vf[1] : -Vf[1];
        VP[2] = -Vf[2];
        VP[3] = -inv_theta2 - Vf[3];
    } else if (bctype == 3) {
        const double Vn = Vf[1] * nx + Vf[2] * ny;
        VP[3] = Vf[3];
        VP[1] = Vf[1] - 2 * Vn * nx;
        VP[2] = Vf[2] - 2 * Vn * ny;
    }
}
// impose_BCs_stress! (:218-262); s*[0..2] = fields 2..4
__device__ __forceinline__ void bc_stress_q(int bctype, int kind, double vlid, double nx, double ny, const double* sxf, const double* syf,
                                            double* sxP, double* syP) {
    if (bctype == 1) {
        sxP[0] = sxf[0]; syP[0] = syf[0];
        sxP[1] = sxf[1]; syP[1] = syf[1];
        sxP[2] = kind == 2 ? -sxf[2] + 2 * vlid * sxf[0] : -sxf[2];
        syP[2] = kind == 2 ? -syf[2] + 2 * vlid * syf[0] : -syf[2];
    } else if (bctype == 2) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { sxP[k] = sxf[k]; syP[k] = syf[k]; }
    } else if (bctype == 3) {
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

// line l of an element: direction d (0: r, nodes i + n a at fixed i = m; 1: s, nodes a + n m), first node, stride and the
// trace indices of its two end points (side 0: coordinate -1, side 1: +1)
template <int n>
__device__ __forceinline__ void quad_line(int ll, int& d, int& base, int& stride, int& f0, int& f1) {
    d = ll / n;
    const int m = ll - d * n;
    if (d == 0) { base = m; stride = n; f0 = 3 * n + (n - 1 - m); f1 = n + m; }          // r = -1: face 4 (reversed), r = +1: face 2
    else { base = n * m; stride = 1; f0 = m; f1 = 2 * n + (n - 1 - m); }                  // s = -1: face 1, s = +1: face 3 (reversed)
}

// CNS = true: the inviscid part of rhsRK! with the wall mirror state (dg2D_CNS_cavity_optimized.jl:157-176, :447-528) plus, per line,
// the BR1 divergence of sigma (k_cnsq_sigma, kernels_cnsq.cuh), its surface term with the stress BCs and the penalisation
// (:590-611, :817-846); final rhs = (viscous numerator - inviscid numerator) / J.
template <int n, bool CNS = false>
__global__ void __launch_bounds__(QuadDims<n>::THREADS, QuadDims<n>::MIN_CTAS) k_quad_stage_line(const __grid_constant__ QuadParams p) {
    using D = QuadDims<n>;
    constexpr int Nq = D::Nq, Nfq = D::Nfq, NL = D::NL, EB = D::EB, NT = D::THREADS, EN = D::EN, EF = D::EF;
    constexpr PairTab<n> PT{};
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[2][NT / 32];
    double* sWV = sm;                       // [6][EN]
    double* sWF = sWV + 6 * EN;             // [7][EF] own traces (Euler only)
    constexpr bool OWN_SMEM = !CNS && ESDG_QUAD_OWN_SMEM;   // own traces staged in shared memory or read by the line thread itself
    double* sP = CNS ? sWF : sWF + QUAD_TP * EF;   // [7][EF] exterior traces
    double* sA = sP + QUAD_TP * EF;         // [2][4][EN]; block 0 later holds V[4][EN]
    double* sS = sA + 8 * EN;               // CNS: [6][EN] sigma at the nodes, [6][EF] exterior sigma traces
    double* sSP = sS + 6 * EN;
    double* sGeo = CNS ? sSP + 6 * EF : sS; // [EB][17]
    double* s_lm = sGeo + EB * QUAD_GEO;
    double* s_lp = s_lm + n;
    double* s_w = s_lp + n;
    const int tid = threadIdx.x;
    const long long e0 = p.e_begin + (long long)blockIdx.x * EB;
    const int nb = (int)min((long long)EB, p.e_begin + p.e_count - e0);
    const long long gq0 = e0 * Nq, gf0 = e0 * Nfq;
    // ---- loads: own traces, exterior traces (mapP gather), geometry by cp.async; A from global -----------------
    // every independent request first (mapP indices, Q and the stashed logs of this thread's nodes), their uses after
    constexpr int FIT = (EF + NT - 1) / NT, AIT = (EN + NT - 1) / NT;
    int ipv[FIT];
#pragma unroll
    for (int k = 0; k < FIT; ++k) {
        const int it = tid + k * NT;
        ipv[k] = it < nb * Nfq ? p.mapP[gf0 + it] : 0;
    }
    double Ua[AIT][4], La[AIT][2];
#pragma unroll
    for (int k = 0; k < AIT; ++k) {
        const int it = tid + k * NT;
        const bool on = it < nb * Nq;
#pragma unroll
        for (int f = 0; f < 4; ++f) Ua[k][f] = on ? p.Q[f * p.q_stride + gq0 + it] : 1.0;
        La[k][0] = on ? p.Lcur[gq0 + it] : 0.0;
        La[k][1] = on ? p.Lcur[p.q_stride + gq0 + it] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < FIT; ++k) {
        const int it = tid + k * NT;
        if (it < nb * Nfq) {
            const int ip = ipv[k];
#pragma unroll
            for (int q = 0; q < QUAD_TP; ++q) {
                if (OWN_SMEM) cp_async8(sWF + q * EF + it, p.Tcur + q * p.t_stride + gf0 + it);
                cp_async8(sP + q * EF + it, p.Tcur + q * p.t_stride + ip);
            }
            if (CNS) {
#pragma unroll
                for (int q = 0; q < 6; ++q) cp_async8(sSP + q * EF + it, p.sig_tr + q * p.t_stride + ip);
            }
        }
    }
    for (int it = tid; it < nb * QUAD_GEO; it += NT) cp_async8(sGeo + it, p.geo + e0 * QUAD_GEO + it);
    if (CNS) {
        for (int it = tid; it < nb * Nq; it += NT) {
#pragma unroll
            for (int q = 0; q < 6; ++q) cp_async8(sS + q * EN + it, p.sig + q * p.q_stride + gq0 + it);
        }
    }
    cp_async_commit();
    if (p.mode == 1) {   // the LSRK register is needed last: pull it into L2 meanwhile
        constexpr int LPF = (EN * 8 + 127) / 128;
        for (int it = tid; it < 4 * LPF; it += NT) prefetch_l2(p.res + (it / LPF) * p.q_stride + gq0 + (it % LPF) * 16);
    }
#if ESDG_QUAD_PF_AHEAD > 0
    {   // pull the inputs of the batch that will run in this CTA's slot next (ESDG_QUAD_PF_AHEAD CTAs further on) into L2
        const long long ea = e0 + (long long)ESDG_QUAD_PF_AHEAD * EB;
        if (ea + EB <= p.e_begin + p.e_count) {
            constexpr int LQ = (EN * 8 + 127) / 128, LF = (EF * 8 + 127) / 128;
            for (int it = tid; it < 4 * LQ; it += NT) prefetch_l2(p.Q + (it / LQ) * p.q_stride + ea * Nq + (it % LQ) * 16);
            for (int it = tid; it < 2 * LQ; it += NT) prefetch_l2(p.Lcur + (it / LQ) * p.q_stride + ea * Nq + (it % LQ) * 16);
            for (int it = tid; it < QUAD_TP * LF; it += NT) prefetch_l2(p.Tcur + (it / LF) * p.t_stride + ea * Nfq + (it % LF) * 16);
        }
    }
#endif
    if (tid < n) { s_lm[tid] = p.lm[tid]; s_lp[tid] = p.lp[tid]; s_w[tid] = p.w[tid]; }
    bool bad = false;
    const double c_igm = 1.0 / (2.0 * (p.gamma - 1.0));
#pragma unroll
    for (int k = 0; k < AIT; ++k) {
        const int it = tid + k * NT;
        if (it < nb * Nq) {
            VD<1> U[4], hu[2], beta, rhoe;
#pragma unroll
            for (int f = 0; f < 4; ++f) U[f][0] = Ua[k][f];
            vol_prim2_v<1>(U, p.gamma, hu, beta, rhoe);
            bad = bad || !(U[0][0] > 0.0) || !(rhoe[0] > 0.0);
            sWV[it] = U[0][0];
            sWV[EN + it] = hu[0][0];
            sWV[2 * EN + it] = hu[1][0];
            sWV[3 * EN + it] = beta[0];
            sWV[4 * EN + it] = La[k][0];
            sWV[5 * EN + it] = La[k][1];
        }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- C: one line per thread ------------------------------------------------------------------------------
    double tvisc = 0.0;
    if (tid < nb * NL) {
        const int eb = tid / NL, ll = tid - eb * NL;
        int d, base, stride, f0, f1;
        quad_line<n>(ll, d, base, stride, f0, f1);
        base += eb * Nq;
        const int fidx[2] = {eb * Nfq + f0, eb * Nfq + f1};
        const double* Gm = sGeo + eb * QUAD_GEO;
        const double g0 = Gm[d], g1 = Gm[2 + d];                         // (d_x, d_y) J of direction d
        Node2 nd[n];
#pragma unroll
        for (int a = 0; a < n; ++a) nd[a] = ld_node2(sWV, EN, base + a * stride);
        double acc[n][4];
#pragma unroll
        for (int a = 0; a < n; ++a)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[a][f] = 0.0;
        const VD<2> gg[2] = {vset<2>(g0), vset<2>(g1)};
        // volume-volume pairs, two at a time
#pragma unroll
        for (int k0 = 0; k0 < PT.NP; k0 += 2) {
            Node2V<2> A, B;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k = k0 + q < PT.NP ? k0 + q : k0;
                set_lane(A, q, nd[PT.a[k]]);
                set_lane(B, q, nd[PT.b[k]]);
            }
            VD<2> F[4];
            pair_flux2_v<2>(A, B, gg, c_igm, F);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (k0 + q < PT.NP) {
                    const int a = PT.a[k0 + q], b = PT.b[k0 + q];
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        acc[a][f] = fma(p.T[a * n + b], F[f][q], acc[a][f]);
                        acc[b][f] = fma(p.T[b * n + a], F[f][q], acc[b][f]);
                    }
                }
            }
        }
        // volume-face pairs (lane = side)
        Node2 fn[2];
        double lamOwn[2];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (!OWN_SMEM) {
                fn[side] = ld_node2_g(p.Tcur, p.t_stride, gf0 + fidx[side]);
                lamOwn[side] = p.Tcur[6 * p.t_stride + gf0 + fidx[side]];
            } else {
                fn[side] = ld_node2(sWF, EF, fidx[side]);
                lamOwn[side] = sWF[6 * EF + fidx[side]];
            }
        }
        double h[2][4];
#pragma unroll
        for (int f = 0; f < 4; ++f) h[0][f] = h[1][f] = 0.0;
#pragma unroll
        for (int a = 0; a < n; ++a) {
            Node2V<2> A, B;
            set_lane(A, 0, nd[a]); set_lane(A, 1, nd[a]);
            set_lane(B, 0, fn[0]); set_lane(B, 1, fn[1]);
            VD<2> F[4];
            pair_flux2_v<2>(A, B, gg, c_igm, F);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                acc[a][f] = fma(-p.lmw[a], F[f][0], acc[a][f]);
                acc[a][f] = fma(p.lpw[a], F[f][1], acc[a][f]);
                h[0][f] = fma(p.lm[a], F[f][0], h[0][f]);
                h[1][f] = fma(p.lp[a], F[f][1], h[1][f]);
            }
        }
        // surface flux at the two face points: F_EC(M,P).nJ - LFc (U+ - U)        (dg2D_euler_quad.jl:158-170)
        VD<2> FS[4];
        {
            Node2V<2> M, P;
            VD<2> nJ[2], sJ, lamM, lamP;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                set_lane(M, side, fn[side]);
                set_lane(P, side, ld_node2(sP, EF, fidx[side]));
                const int face = (side ? f1 : f0) / n;
                const double* gf = Gm + 5 + face * 3;
                nJ[0][side] = gf[0]; nJ[1][side] = gf[1]; sJ[side] = gf[2];
                lamM[side] = lamOwn[side];
                lamP[side] = sP[6 * EF + fidx[side]];
            }
            VD<2> jump[4];
            const bool lf = p.lf_scale != 0.0;
            if (lf) {   // U+ - U of the GATHERED exterior state (a boundary point maps to itself: no penalty there, :500-507)
                const VD<2> ibM = vrcp(M.beta), ibP = vrcp(P.beta);
                const VD<2> qM = vfma(M.u[0], M.u[0], M.u[1] * M.u[1]), qP = vfma(P.u[0], P.u[0], P.u[1] * P.u[1]);
                const VD<2> tM = 2.0 * M.rho, tP = 2.0 * P.rho;
                jump[0] = P.rho - M.rho;
                jump[1] = vfma(tP, P.u[0], -(tM * M.u[0]));
                jump[2] = vfma(tP, P.u[1], -(tM * M.u[1]));
                jump[3] = vfma(P.rho, vfma(ibP, c_igm, 2.0 * qP), -(M.rho * vfma(ibM, c_igm, 2.0 * qM)));
            }
            if (CNS) {   // impose_BCs_inviscid! (:157-176): rho+ = rho, beta+ = beta, u+ = u - 2 (u.n) n at wall / lid points
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const int kd = p.bc[gf0 + fidx[side]] & 7;
                    if (kd == 1 || kd == 2) {
                        const double isJ = fast_rcp(sJ[side]);
                        const double nx = nJ[0][side] * isJ, ny = nJ[1][side] * isJ;
                        const double Un = fn[side].u[0] * nx + fn[side].u[1] * ny;
                        P.rho[side] = fn[side].rho; P.beta[side] = fn[side].beta; P.lr[side] = fn[side].lr; P.lb[side] = fn[side].lb;
                        P.u[0][side] = fn[side].u[0] - 2 * Un * nx;
                        P.u[1][side] = fn[side].u[1] - 2 * Un * ny;
                    }
                }
            }
            pair_flux2_v<2>(M, P, nJ, c_igm, FS);
            if (lf) {
                const VD<2> LFc = p.lf_scale * vmax(lamM, lamP) * sJ;
#pragma unroll
                for (int f = 0; f < 4; ++f) FS[f] = vfma(-LFc, jump[f], FS[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const double Hm = FS[f][0] + h[0][f], Hp = FS[f][1] - h[1][f];
#pragma unroll
            for (int a = 0; a < n; ++a) acc[a][f] = fma(p.lpw[a], Hp, fma(p.lmw[a], Hm, acc[a][f]));
        }
        double* A = sA + d * 4 * EN;
#pragma unroll
        for (int a = 0; a < n; ++a)
#pragma unroll
            for (int f = 0; f < 4; ++f) A[f * EN + base + a * stride] = acc[a][f];
        if constexpr (CNS) {
            // viscous numerator of this line: D (g_x sigma_x + g_y sigma_y) + lift(.5 [[sigma]].nJ + J tau [[v]])
            const double Je = Gm[4];
            double ft[2][3];
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const double* l = side ? p.lp : p.lm;
                double sxf[3] = {0, 0, 0}, syf[3] = {0, 0, 0}, sxP[3], syP[3];
#pragma unroll
                for (int a = 0; a < n; ++a)
#pragma unroll
                    for (int f = 0; f < 3; ++f) {
                        sxf[f] = fma(l[a], sS[f * EN + base + a * stride], sxf[f]);
                        syf[f] = fma(l[a], sS[(3 + f) * EN + base + a * stride], syf[f]);
                    }
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    sxP[f] = sSP[f * EF + fidx[side]];
                    syP[f] = sSP[(3 + f) * EF + fidx[side]];
                }
                const double* gf = Gm + 5 + ((side ? f1 : f0) / n) * 3;
                const double isJ = fast_rcp(gf[2]);
                const double nx = gf[0] * isJ, ny = gf[1] * isJ;
                const long long self = gf0 + fidx[side];
                const int code = p.bc[self];
                const int kd = code & 7;
                const double vlid = kd == 2 ? p.vlid[self] : 0.0;
                if (kd == 1 || kd == 2) bc_stress_q(p.bctype, kd, vlid, nx, ny, sxf, syf, sxP, syP);
#pragma unroll
                for (int f = 0; f < 3; ++f) ft[side][f] = .5 * ((sxP[f] - sxf[f]) * gf[0] + (syP[f] - syf[f]) * gf[1]);
                if (p.viscous_dissp) {   // penalisation tau [[v]] (:817-840)
                    double VM[4], VPv[4];
                    v234_from_prim(fn[side].u[0], fn[side].u[1], fn[side].beta, p.gamma, VM);
                    v234_from_prim(sP[EF + fidx[side]], sP[2 * EF + fidx[side]], sP[3 * EF + fidx[side]], p.gamma, VPv);
                    if (kd == 1 || kd == 2) bc_entropy_vars_q(p.bctype, kd, vlid, p.inv_theta2, nx, ny, VM, VPv);
                    const double iv4 = fast_rcp(VM[3]);
                    const double tau = -iv4 / p.Re;
                    const double d1 = VPv[1] - VM[1], d2 = VPv[2] - VM[2], d3 = VPv[3] - VM[3];
                    double p3 = tau * d3;
                    if (code & 8) {
                        const double a1 = .5 * (VPv[1] + VM[1]), a2 = .5 * (VPv[2] + VM[2]);
                        p3 = p.bctype == 1 ? -tau * (a1 * d1 + a2 * d2) * iv4 : -tau * (a1 * d1 + a2 * d2 + d3 * d3 * .5) * iv4;
                    }
                    ft[side][0] = fma(Je, tau * d1, ft[side][0]);
                    ft[side][1] = fma(Je, tau * d2, ft[side][1]);
                    ft[side][2] = fma(Je, p3, ft[side][2]);
                }
            }
            const int m = ll - d * n;
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                double accv[n];
#pragma unroll
                for (int a = 0; a < n; ++a) accv[a] = fma(p.lmw[a], ft[0][f], p.lpw[a] * ft[1][f]);
#pragma unroll
                for (int b = 0; b < n; ++b) {
                    const double S = fma(g1, sS[(3 + f) * EN + base + b * stride], g0 * sS[f * EN + base + b * stride]);
#pragma unroll
                    for (int a = 0; a < n; ++a) accv[a] = fma(p.D[a * n + b], S, accv[a]);
                }
#pragma unroll
                for (int a = 0; a < n; ++a) {
                    if (p.want_rhstest) {   // sum wJq VUq . visc_rhs (:966): wJ (accv / J) = w accv; v_{2+f} from the node primitives
                        const double c = 2.0 * (p.gamma - 1.0) * sWV[3 * EN + base + a * stride];
                        const double v = f == 2 ? -c : 2.0 * c * sWV[(1 + f) * EN + base + a * stride];
                        tvisc = fma(p.w[a] * p.w[m] * v, accv[a], tvisc);
                    }
                    A[(1 + f) * EN + base + a * stride] -= accv[a];   // this thread's own slots
                }
            }
        }
    }
    // E operands: requested before the barrier, all in flight together (the LSRK register is touched nowhere else)
    constexpr int EIT = (EN + NT - 1) / NT;
    double Ue[EIT][4], Re[EIT][4];
#pragma unroll
    for (int k = 0; k < EIT; ++k) {
        const int it = tid + k * NT;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            Ue[k][f] = it < nb * Nq ? p.Q[f * p.q_stride + gq0 + it] : 1.0;
            Re[k][f] = (it < nb * Nq && p.mode == 1) ? p.res[f * p.q_stride + gq0 + it] : 0.0;
        }
    }
    __syncthreads();
    // ---- E: update ------------------------------------------------------------------------------------------
    double tsum = 0.0;
#pragma unroll
    for (int k = 0; k < EIT; ++k) {
        const int it = tid + k * NT;
        if (it >= nb * Nq) break;
        const long long g = gq0 + it;
        const int eb = it / Nq, i = it - eb * Nq;
        const double J = sGeo[eb * QUAD_GEO + 4];
        VD<1> U[4], rr[4], Jv;
        Jv[0] = J;
        const VD<1> mJ = -vrcp(Jv);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            U[f][0] = Ue[k][f];
            rr[f][0] = (sA[f * EN + it] + sA[(4 + f) * EN + it]) * mJ[0];
        }
        if (p.want_rhstest) {
            VD<1> V[4], lr, lb, hu[2], beta, rhoe;
            lr[0] = sWV[4 * EN + it];
            lb[0] = sWV[5 * EN + it];
            vol_prim2_v<1>(U, p.gamma, hu, beta, rhoe);
            entropy_vars2_v<1>(U, rhoe, lr, lb, p.gamma, V);
            const double wJ = s_w[i % n] * s_w[i / n] * J;
#pragma unroll
            for (int f = 0; f < 4; ++f) tsum = fma(wJ * V[f][0], rr[f][0], tsum);
        }
        if (p.mode == 0) {
#pragma unroll
            for (int f = 0; f < 4; ++f) p.rhs[f * p.q_stride + g] = rr[f][0];
        } else {
            VD<1> Un[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const double rs = fma(p.rk_a, Re[k][f], p.dt * rr[f][0]);
                Un[f][0] = fma(p.rk_b, rs, U[f][0]);
                p.res[f * p.q_stride + g] = rs;
                p.Qout[f * p.q_stride + g] = Un[f][0];
            }
            VD<1> hu[2], beta, rhoen, V[4];
            vol_prim2_v<1>(Un, p.gamma, hu, beta, rhoen);
            const VD<1> lr = vlog_pos<1>(Un[0]), lb = vlog_pos<1>(beta);
            entropy_vars2_v<1>(Un, rhoen, lr, lb, p.gamma, V);
            bad = bad || !(Un[0][0] > 0.0) || !(rhoen[0] > 0.0);
            p.Lnext[g] = lr[0];
            p.Lnext[p.q_stride + g] = lb[0];
#pragma unroll
            for (int f = 0; f < 4; ++f) sA[f * EN + it] = V[f][0];
        }
    }
    if (p.mode == 1) {
        __syncthreads();
        // ---- F: traces of the new state, both ends of a line per thread -----------------------------------------
        if (tid < nb * NL) {
            const int eb = tid / NL, ll = tid - eb * NL;
            int d, base, stride, f0, f1;
            quad_line<n>(ll, d, base, stride, f0, f1);
            const double* v = sA + eb * Nq + base;
            VD<2> V[4], nh[2], Tv[QUAD_TP];
#pragma unroll
            for (int f = 0; f < 4; ++f) V[f][0] = V[f][1] = 0.0;
#pragma unroll
            for (int a = 0; a < n; ++a) {
                const double la = s_lm[a], lb = s_lp[a];
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const double x = v[f * EN + a * stride];
                    V[f][0] = fma(la, x, V[f][0]);
                    V[f][1] = fma(lb, x, V[f][1]);
                }
            }
            const double* ge = sGeo + eb * QUAD_GEO;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const double* gf = ge + 5 + ((side ? f1 : f0) / n) * 3;
                const double s2 = 2.0 * fast_rcp(gf[2]);
                nh[0][side] = gf[0] * s2; nh[1][side] = gf[1] * s2;
            }
            trace_from_entropy2_v<2>(V, nh, p.gamma, p.lg1, p.l2g1, Tv);
            double* T0 = p.Tnext + (e0 + eb) * Nfq;
#pragma unroll
            for (int q = 0; q < QUAD_TP; ++q) {
                T0[q * p.t_stride + f0] = Tv[q][0];
                T0[q * p.t_stride + f1] = Tv[q][1];
            }
        }
    }
    if (p.want_rhstest) {
        for (int o = 16; o > 0; o >>= 1) {
            tsum += __shfl_down_sync(0xffffffffu, tsum, o);
            if (CNS) tvisc += __shfl_down_sync(0xffffffffu, tvisc, o);
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = tsum; red[1][tid >> 5] = tvisc; }
        __syncthreads();
        if (tid == 0) {
            double t = 0.0, tv = 0.0;
            for (int wv = 0; wv < NT / 32; ++wv) { t += red[0][wv]; tv += red[1][wv]; }
            if (CNS) {
                p.partial[blockIdx.x] = t;
                p.partial_vrhs[blockIdx.x] = tv;
            } else {
                p.partial[e0] = t;   // one slot per element, as on the hex path (esdg_ctx.cu:total_partials)
            }
        }
    }
    if (bad) atomicOr(p.bad, 1);
}

// traces + log stash of a given state (after an upload / for the rhs-only entry point)
template <int n>
__global__ void __launch_bounds__(QuadDims<n>::THREADS) k_quad_traces(const __grid_constant__ QuadParams p) {
    using D = QuadDims<n>;
    constexpr int Nq = D::Nq, Nfq = D::Nfq, NL = D::NL, EB = D::EB, NT = D::THREADS, EN = D::EN;
    extern __shared__ __align__(16) double sm[];
    double* sV = sm;                        // [4][EN]
    double* sGeo = sV + 4 * EN;
    double* s_lm = sGeo + EB * QUAD_GEO;
    double* s_lp = s_lm + n;
    const int tid = threadIdx.x;
    const long long e0 = p.e_begin + (long long)blockIdx.x * EB;
    const int nb = (int)min((long long)EB, p.e_begin + p.e_count - e0);
    if (tid < n) { s_lm[tid] = p.lm[tid]; s_lp[tid] = p.lp[tid]; }
    for (int it = tid; it < nb * QUAD_GEO; it += NT) sGeo[it] = p.geo[e0 * QUAD_GEO + it];
    bool bad = false;
    for (int it = tid; it < nb * Nq; it += NT) {
        const long long g = e0 * Nq + it;
        VD<1> U[4], hu[2], beta, rhoe, V[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) U[f][0] = p.Q[f * p.q_stride + g];
        vol_prim2_v<1>(U, p.gamma, hu, beta, rhoe);
        const bool ok = U[0][0] > 0.0 && rhoe[0] > 0.0;
        bad = bad || !ok;
        const VD<1> lr = vlog_pos<1>(vset<1>(ok ? U[0][0] : 1.0)), lb = vlog_pos<1>(vset<1>(ok ? beta[0] : 1.0));
        entropy_vars2_v<1>(U, rhoe, lr, lb, p.gamma, V);
        p.Lnext[g] = lr[0];
        p.Lnext[p.q_stride + g] = lb[0];
#pragma unroll
        for (int f = 0; f < 4; ++f) sV[f * EN + it] = V[f][0];
    }
    __syncthreads();
    if (tid < nb * NL) {
        const int eb = tid / NL, ll = tid - eb * NL;
        int d, base, stride, f0, f1;
        quad_line<n>(ll, d, base, stride, f0, f1);
        const double* v = sV + eb * Nq + base;
        VD<2> V[4], nh[2], Tv[QUAD_TP];
#pragma unroll
        for (int f = 0; f < 4; ++f) V[f][0] = V[f][1] = 0.0;
#pragma unroll
        for (int a = 0; a < n; ++a) {
            const double la = s_lm[a], lb = s_lp[a];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const double x = v[f * EN + a * stride];
                V[f][0] = fma(la, x, V[f][0]);
                V[f][1] = fma(lb, x, V[f][1]);
            }
        }
        const double* ge = sGeo + eb * QUAD_GEO;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const double* gf = ge + 5 + ((side ? f1 : f0) / n) * 3;
            const double s2 = 2.0 * fast_rcp(gf[2]);
            nh[0][side] = gf[0] * s2; nh[1][side] = gf[1] * s2;
        }
        trace_from_entropy2_v<2>(V, nh, p.gamma, p.lg1, p.l2g1, Tv);
        double* T0 = p.Tnext + (e0 + eb) * Nfq;
#pragma unroll
        for (int q = 0; q < QUAD_TP; ++q) {
            T0[q * p.t_stride + f0] = Tv[q][0];
            T0[q * p.t_stride + f1] = Tv[q][1];
        }
    }
    if (bad) atomicOr(p.bad, 1);
}

}  // namespace esdg
