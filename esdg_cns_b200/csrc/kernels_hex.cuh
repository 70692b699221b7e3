// Tensor-product hex kernels (collocated Gauss nodes, affine elements) -- the product path
// for BASELINE configs 4/5.  One launch per LSRK stage does everything the reference's
// rhs() + stage update do (examples/dg3D_euler_hex.jl:167-222, dg2D_euler_quad.jl:198-207):
//
//   A  stage Q; (rho,u,beta,logs) at volume nodes; own face traces -> shared memory
//   B  surface: exterior trace gather through mapP, EC flux . nJ, Lax-Friedrichs penalty   (:182-197)
//   C  flux differencing along the r, s, t grid lines.  Every UNORDERED pair of the
//      hybridized operator is evaluated once (the symmetric loop of
//      CompressibleNS/dg2D_CNS_cavity_optimized.jl:326-347): each volume thread owns the
//      pairs (x, x+k mod n) of its line plus its 2 face partners per direction, and hands
//      the contracted flux to the partner through shared memory
//   D  partner gather; per face node H = -nrJ*sum_a l[a]*F + flux*   (Ph / Lf applied as 1-D sweeps)
//   E  rhs = -(...)/J ; rhstest ; res = a*res + dt*rhs ; Q += b*res                       (:205-206)
//   F  entropy projection of the NEW state to the faces -> traces for the next stage      (:175-178)
//
// Operators never appear as matrices: the kernel holds the 1-D pair coefficients
// T[a][a'] = (w_a D[a,a'] - w_a' D[a',a]) / w_a, the end-point interpolation vectors lm, lp and
// the Gauss weights.  esdg_finalize proves (to 1e-10) that the caller's Qrh/Qsh/Qth, Ef, Ph, Lf
// ARE the Kronecker products of these before selecting this path; otherwise the generic
// table-driven kernels run.
//
// Trace format (7 planes): rho, u/2, v/2, w/2, beta, log(rho), log(beta) at every face point -- what
// the neighbour's surface flux needs, so the exterior logs are not recomputed.
#pragma once
#include "fastmath.cuh"
#include "physics.cuh"

namespace esdg {

constexpr int HEX_TP = 7;   // trace planes (== ESDG_HEX_TRACE_PLANES in hex_fast.h)

struct HexParams {
    long long K, e_begin, e_count;
    long long q_stride;      // Nq*K
    long long t_stride;      // Nfq*K + n_halo
    const double* Q;
    double* Qout;
    double* res;
    double* rhs;
    const double* Lcur;      // stashed (log rho, log beta) of Q at volume nodes: 2 planes of q_stride
    double* Lnext;
    const double* Tcur;
    double* Tnext;
    const int* mapP;
    const double* geo;
    int geo_stride;
    double gamma, lf_scale, lg1, l2g1;   // lg1 = log(gamma-1), l2g1 = log(2(gamma-1))
    double rk_a, rk_b, dt;
    int mode;                // 0: write rhs ; 1: fused LSRK stage + traces of the new state
    int want_rhstest;
    double* partial;
    int* bad;
    double T[64], lm[8], lp[8], w[8];
    double lmw[8], lpw[8];   // lm[a]/w[a], lp[a]/w[a]
};

template <int n>
struct HexDims {
    static constexpr int Nq = n * n * n, nfp = n * n, Nfq = 6 * nfp, Nh = Nq + Nfq;
    static constexpr int KH = n / 2;                         // owned offsets per line
    static constexpr int EB = (128 / Nq) > 0 ? (128 / Nq) : 1;
    static constexpr int THREADS = ((EB * Nq + 31) / 32) * 32;
    static constexpr int XV = KH * 5 * Nq, XF = 2 * 5 * nfp * n;
    static constexpr int XSZ = (XV + XF) > 5 * Nq ? (XV + XF) : 5 * Nq;
    static constexpr int TAB = n * n + 5 * n + EB * 34;   // 1-D tables + per-element geometry records
#ifndef ESDG_HEX_STAGE_RES
#define ESDG_HEX_STAGE_RES 1   // 1: LSRK register staged in shared memory by cp.async; 0: read in phase E (smaller footprint)
#endif
    static constexpr int SMEM_DOUBLES = EB * (7 * Nh + XSZ + 7 * Nfq + (ESDG_HEX_STAGE_RES ? 5 * Nq : 0)) + TAB;
};

struct Node {
    double rho, u[3], beta, lr, lb;
};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* q) { asm volatile("prefetch.global.L2 [%0];" ::"l"(q)); }

__device__ __forceinline__ Node ld_node(const double* W, int Nh, int idx) {
    Node a;
    a.rho = W[idx];
    a.u[0] = W[Nh + idx];
    a.u[1] = W[2 * Nh + idx];
    a.u[2] = W[3 * Nh + idx];
    a.beta = W[4 * Nh + idx];
    a.lr = W[5 * Nh + idx];
    a.lb = W[6 * Nh + idx];
    return a;
}

// Entropy-conservative flux (euler_fluxes.jl:51-89 with logmean.jl:14-28) contracted with a
// direction vector g: F[f] = sum_d g[d]*F_d[f](A,B).  DIAG: only g[D] is non-zero.
// One reciprocal serves the three quotients rholog = da/dlog, 1/betalog and pa = rhoavg/(bA+bB).
template <bool DIAG, int D>
__device__ __forceinline__ void pair_flux(const Node& A, const Node& B, const double* g, double c_igm, double* F) {
    const double dar = B.rho - A.rho, hr = .5 * (A.rho + B.rho);
    const double dab = B.beta - A.beta, sb = A.beta + B.beta, hb = .5 * sb;
    const bool tr = fabs(dar) < 1e-4 * hr, tb = fabs(dab) < 1e-4 * hb;
    const double X = tr ? hr : (B.lr - A.lr);
    const double Y = tb ? hb : dab;
    const double XY = X * Y;
    const double R = fast_rcp(XY * sb);
    const double iX = R * (Y * sb), iY = R * (X * sb), iZ = R * XY;
    const double qr = dar * iX;   // f (Taylor branch) or the log mean itself
    const double vr = qr * qr;
    const double rholog = tr ? hr * (1.0 + vr * (-.2 - vr * (.0512 - vr * 0.026038857142857))) : qr;
    const double fb = dab * iY, vb = fb * fb;
    const double xb = vb * (-.2 - vb * (.0512 - vb * 0.026038857142857));
    const double ibl = iY * (tb ? (1.0 - xb) : (B.lb - A.lb));   // 1/betalog
    const double pa = hr * iZ;                                    // rhoavg/(betaL+betaR)
    double uavg[3], unorm = 0.0, un;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        uavg[d] = A.u[d] + B.u[d];                 // Node.u holds HALF velocities
        unorm = fma(A.u[d], B.u[d], unorm);        // = uL.uR / 4
    }
    if (DIAG) {
        un = g[D] * uavg[D];
    } else {
        un = g[0] * uavg[0];
        un = fma(g[1], uavg[1], un);
        un = fma(g[2], uavg[2], un);
    }
    const double Ep = fma(rholog, fma(unorm, 2.0, ibl * c_igm), pa);   // rholog/(2(g-1)betalog)+pa+.5 rholog uL.uR
    const double m = rholog * un;
    F[0] = m;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (DIAG)
            F[1 + d] = (d == D) ? fma(m, uavg[d], pa * g[d]) : m * uavg[d];
        else
            F[1 + d] = fma(m, uavg[d], pa * g[d]);
    }
    F[4] = Ep * un;
}

// Two independent two-point fluxes against the same node A, written statement-interleaved so that the
// scheduler sees two dependency chains (the FP64 pipe is latency-bound with 4 warps per scheduler).
#define ESDG_Q2 for (int q = 0; q < 2; ++q)
template <bool DIAG, int D>
__device__ __forceinline__ void pair_flux2(const Node& A, const Node* B, const double* g, double c_igm, double (*F)[5]) {
    double dar[2], hr[2], dab[2], sb[2], hb[2], X[2], Y[2], XY[2], R[2], iX[2], iY[2], iZ[2];
    bool tr[2], tb[2];
#pragma unroll
    ESDG_Q2 { dar[q] = B[q].rho - A.rho; hr[q] = .5 * (A.rho + B[q].rho); }
#pragma unroll
    ESDG_Q2 { dab[q] = B[q].beta - A.beta; sb[q] = A.beta + B[q].beta; hb[q] = .5 * sb[q]; }
#pragma unroll
    ESDG_Q2 { tr[q] = fabs(dar[q]) < 1e-4 * hr[q]; tb[q] = fabs(dab[q]) < 1e-4 * hb[q]; }
#pragma unroll
    ESDG_Q2 { X[q] = tr[q] ? hr[q] : (B[q].lr - A.lr); Y[q] = tb[q] ? hb[q] : dab[q]; }
#pragma unroll
    ESDG_Q2 XY[q] = X[q] * Y[q];
    {   // two reciprocals, Newton steps interleaved
        double y[2], e[2], x[2];
#pragma unroll
        ESDG_Q2 { x[q] = XY[q] * sb[q]; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[q]) : "d"(x[q])); }
#pragma unroll
        ESDG_Q2 e[q] = fma(-x[q], y[q], 1.0);
#pragma unroll
        ESDG_Q2 y[q] = fma(y[q], e[q], y[q]);
#pragma unroll
        ESDG_Q2 e[q] = fma(-x[q], y[q], 1.0);
#pragma unroll
        ESDG_Q2 R[q] = fma(y[q], e[q], y[q]);
    }
#pragma unroll
    ESDG_Q2 { iX[q] = R[q] * (Y[q] * sb[q]); iY[q] = R[q] * (X[q] * sb[q]); iZ[q] = R[q] * XY[q]; }
    double qr[2], vr[2], rholog[2], fb[2], vb[2], xb[2], ibl[2], pa[2];
#pragma unroll
    ESDG_Q2 { qr[q] = dar[q] * iX[q]; fb[q] = dab[q] * iY[q]; }
#pragma unroll
    ESDG_Q2 { vr[q] = qr[q] * qr[q]; vb[q] = fb[q] * fb[q]; }
#pragma unroll
    ESDG_Q2 rholog[q] = tr[q] ? hr[q] * (1.0 + vr[q] * (-.2 - vr[q] * (.0512 - vr[q] * 0.026038857142857))) : qr[q];
#pragma unroll
    ESDG_Q2 xb[q] = vb[q] * (-.2 - vb[q] * (.0512 - vb[q] * 0.026038857142857));
#pragma unroll
    ESDG_Q2 { ibl[q] = iY[q] * (tb[q] ? (1.0 - xb[q]) : (B[q].lb - A.lb)); pa[q] = hr[q] * iZ[q]; }
    double uavg[2][3], unorm[2] = {0.0, 0.0}, un[2];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        ESDG_Q2 { uavg[q][d] = A.u[d] + B[q].u[d]; unorm[q] = fma(A.u[d], B[q].u[d], unorm[q]); }
    }
#pragma unroll
    ESDG_Q2 {
        if (DIAG) {
            un[q] = g[D] * uavg[q][D];
        } else {
            un[q] = g[0] * uavg[q][0];
            un[q] = fma(g[1], uavg[q][1], un[q]);
            un[q] = fma(g[2], uavg[q][2], un[q]);
        }
    }
    double Ep[2], m[2];
#pragma unroll
    ESDG_Q2 { Ep[q] = fma(rholog[q], fma(unorm[q], 2.0, ibl[q] * c_igm), pa[q]); m[q] = rholog[q] * un[q]; }
#pragma unroll
    ESDG_Q2 {
        F[q][0] = m[q];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (DIAG)
                F[q][1 + d] = (d == D) ? fma(m[q], uavg[q][d], pa[q] * g[d]) : m[q] * uavg[q][d];
            else
                F[q][1 + d] = fma(m[q], uavg[q][d], pa[q] * g[d]);
        }
        F[q][4] = Ep[q] * un[q];
    }
}
#undef ESDG_Q2

// volume-node primitives from conservative state; logs supplied (stash) or computed
__device__ __forceinline__ void vol_prim(const double* U, double gamma, Node& P, double& rhoe, bool& bad) {
    const double hir = .5 * fast_rcp(U[0]);
    double ke = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        P.u[d] = U[1 + d] * hir;                    // half velocity
        ke = fma(U[1 + d], P.u[d], ke);            // = .5 |rhoU|^2 / rho
    }
    rhoe = U[4] - ke;
    bad = bad || !(U[0] > 0.0) || !(rhoe > 0.0);
    P.rho = U[0];
    P.beta = .5 * U[0] * fast_rcp((gamma - 1.0) * rhoe);
}

// v_ufun (euler_variables.jl:79-86) with s = log((g-1)rhoe/rho^g) = (1-g) log rho - log beta - log 2
__device__ __forceinline__ void entropy_vars(const double* U, double rhoe, double lr, double lb, double gamma, double* V) {
    const double s = (1.0 - gamma) * lr - lb - 0.6931471805599453094;
    const double ire = fast_rcp(rhoe);
    V[0] = fma(-U[4], ire, gamma + 1.0 - s);
    V[1] = U[1] * ire;
    V[2] = U[2] * ire;
    V[3] = U[3] * ire;
    V[4] = -U[0] * ire;
}

// u_vfun (euler_variables.jl:95-113) followed by betafun / velocity / logs, all from the projected
// entropy variables: rhoeV = exp((log(g-1) - g log(-vE) - s)/(g-1)), rho = rhoeV(-vE),
// u = vU/(-vE), beta = -vE/(2(g-1)) (identically rho/(2p)), log rho and log beta follow without
// further transcendental calls.
__device__ __forceinline__ void trace_from_entropy(const double* V, double gamma, double lg1, double l2g1, double* Tv) {
    const double nvE = -V[4];
    const double inv = fast_rcp(nvE);
    const double v2 = fma(V[1], V[1], fma(V[2], V[2], V[3] * V[3]));
    const double s = gamma - V[0] - .5 * v2 * inv;
    const double lv = log_pos(nvE);
    const double ig = 1.0 / (gamma - 1.0);
    const double lrhoe = (lg1 - gamma * lv - s) * ig;   // lg1 = log(gamma-1), host-computed
    const double rhoeV = exp_mod(lrhoe);
    Tv[0] = rhoeV * nvE;
    const double hinv = .5 * inv;                        // traces carry HALF velocities, like Node.u
    Tv[1] = V[1] * hinv;
    Tv[2] = V[2] * hinv;
    Tv[3] = V[3] * hinv;
    Tv[4] = .5 * nvE * ig;
    Tv[5] = lrhoe + lv;
    Tv[6] = lv - l2g1;                                   // l2g1 = log(2(gamma-1))
}

// face point (face = 2*d + side, m) -> first volume node of its line and the line stride
template <int n>
__device__ __forceinline__ void face_line(int d, int m, int& base, int& stride) {
    const int ip = m % n, jp = m / n;
    if (d == 0) { base = n * jp + n * n * ip; stride = 1; }
    else if (d == 1) { base = jp + n * n * ip; stride = n; }
    else { base = jp + n * ip; stride = n * n; }
}

// Phase F: project entropy variables (shared V[5][Nq] per element) to the faces, write 7-plane traces
template <int n>
__device__ __forceinline__ void write_traces(const HexParams& p, const double* sV, const double* s_lm, const double* s_lp,
                                             long long e0, int nb, int ftid) {
    using D = HexDims<n>;
    for (int it = ftid; it < nb * D::Nfq; it += blockDim.x) {
        const int eb = it / D::Nfq, k = it - eb * D::Nfq;
        const int face = k / D::nfp, m = k - face * D::nfp;
        int base, stride;
        face_line<n>(face >> 1, m, base, stride);
        const double* l = (face & 1) ? s_lp : s_lm;
        const double* v = sV + eb * 5 * D::Nq + base;
        double V[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int a = 0; a < n; ++a) {
            const double la = l[a];
#pragma unroll
            for (int f = 0; f < 5; ++f) V[f] = fma(la, v[f * D::Nq + a * stride], V[f]);
        }
        double Tv[HEX_TP];
        trace_from_entropy(V, p.gamma, p.lg1, p.l2g1, Tv);
        const long long g = (e0 + eb) * D::Nfq + k;
#pragma unroll
        for (int q = 0; q < HEX_TP; ++q) p.Tnext[q * p.t_stride + g] = Tv[q];
    }
}

// Traces (+ log stash) of a given state: used once after upload and by the rhs-only entry point.
template <int n>
__global__ void __launch_bounds__(HexDims<n>::THREADS) k_hex_traces(const __grid_constant__ HexParams p) {
    using D = HexDims<n>;
    extern __shared__ double sm[];
    double* sV = sm;                       // [EB][5][Nq]
    double* s_lm = sm + D::EB * 5 * D::Nq;
    double* s_lp = s_lm + n;
    const int tid = threadIdx.x;
    if (tid < n) { s_lm[tid] = p.lm[tid]; s_lp[tid] = p.lp[tid]; }
    const long long e0 = p.e_begin + (long long)blockIdx.x * D::EB;
    const int nb = (int)min((long long)D::EB, p.e_begin + p.e_count - e0);
    bool bad = false;
    if (tid < nb * D::Nq) {
        const int eb = tid / D::Nq, i = tid - eb * D::Nq;
        const long long g = (e0 + eb) * D::Nq + i;
        double U[5], V[5], rhoe;
#pragma unroll
        for (int f = 0; f < 5; ++f) U[f] = p.Q[f * p.q_stride + g];
        Node P;
        vol_prim(U, p.gamma, P, rhoe, bad);
        const double lr = log_pos(P.rho), lb = log_pos(P.beta);
        p.Lnext[g] = lr;
        p.Lnext[p.q_stride + g] = lb;
        entropy_vars(U, rhoe, lr, lb, p.gamma, V);
#pragma unroll
        for (int f = 0; f < 5; ++f) sV[(eb * 5 + f) * D::Nq + i] = V[f];
    }
    __syncthreads();
    write_traces<n>(p, sV, s_lm, s_lp, e0, nb, threadIdx.x);
    if (bad) atomicOr(p.bad, 1);
}

template <int n, bool DIAG>
#ifdef ESDG_HEX_MIN_CTAS   // occupancy experiments; by default ptxas picks ~122 registers (4 CTAs/SM), which measured fastest
__global__ void __launch_bounds__(HexDims<n>::THREADS, ESDG_HEX_MIN_CTAS) k_hex_stage(
#else
__global__ void __launch_bounds__(HexDims<n>::THREADS) k_hex_stage(
#endif
    const __grid_constant__ HexParams p) {
    using D = HexDims<n>;
    constexpr int Nq = D::Nq, nfp = D::nfp, Nfq = D::Nfq, Nh = D::Nh, KH = D::KH;
    extern __shared__ double sm[];
    double* sW = sm;                              // [EB][7][Nh]
    double* sX = sW + D::EB * 7 * Nh;             // [EB][XSZ]   (XV | XF), later V[5][Nq]
    double* sP = sX + D::EB * D::XSZ;             // [EB][7][Nfq] exterior traces ...
    double* sFL = sP;                             // ... overwritten in place by the surface flux (first 5 planes)
    double* sR = sP + D::EB * 7 * Nfq;            // [EB][5][Nq] LSRK register (mode 1)
    double* s_T = sR + (ESDG_HEX_STAGE_RES ? D::EB * 5 * Nq : 0);   // [n][n]
    double* s_lm = s_T + n * n;
    double* s_lp = s_lm + n;
    double* s_lmw = s_lp + n;                     // lm[a]/w[a]
    double* s_lpw = s_lmw + n;
    double* s_w = s_lpw + n;
    double* sGeo = s_w + n;                       // [EB][34]: G[9], J, 6 x (nxJ,nyJ,nzJ,sJ)
    const int tid = threadIdx.x;
    // face-item loops use a thread id rotated by a whole number of warps per CTA: the leftover pass
    // (Nfq is not a multiple of the block size) then lands on a different SM sub-partition for each
    // co-resident CTA instead of always on warp 0
    const int ftid = (tid + 32 * (int)(blockIdx.x % (D::THREADS / 32))) % D::THREADS;
    if (tid < n * n) s_T[tid] = p.T[tid];
    if (tid < n) {
        s_lm[tid] = p.lm[tid];
        s_lp[tid] = p.lp[tid];
        s_lmw[tid] = p.lm[tid] / p.w[tid];
        s_lpw[tid] = p.lp[tid] / p.w[tid];
        s_w[tid] = p.w[tid];
    }
    const long long e0 = p.e_begin + (long long)blockIdx.x * D::EB;
    const int nb = (int)min((long long)D::EB, p.e_begin + p.e_count - e0);
    const bool vol = tid < nb * Nq;
    const int eb = vol ? tid / Nq : 0;
    const int i = tid - eb * Nq;
    const int xa = i % n, xb = (i / n) % n, xc = i / (n * n);
    const long long e = e0 + eb;
    const long long gq = e * Nq + i;
    const double c_igm = 1.0 / (2.0 * (p.gamma - 1.0));
    bool bad = false;

    // ---- prologue: every global read of this element batch is put in flight before any math ----
    // (1) exterior-trace indices, (2) own traces -> sW face rows and (3) exterior traces -> sP with
    // cp.async (no register staging), (4) L2 prefetch of the LSRK register read in phase E.
    constexpr int MAXIT = (D::EB * Nfq + D::THREADS - 1) / D::THREADS;
    constexpr int PF_DIST = 592;   // CTAs in flight on a B200 (148 SMs x 4): distance of the L2 warm-up
    int ipv[MAXIT];
#pragma unroll
    for (int r = 0; r < MAXIT; ++r) {
        const int it = ftid + r * D::THREADS;
        ipv[r] = it < nb * Nfq ? p.mapP[e0 * Nfq + it] : -1;
    }
    double U[5] = {1, 0, 0, 0, 1}, rhoe = 1.0;
    Node Pi;
    Pi.lr = Pi.lb = 0.0;
    if (vol) {
#pragma unroll
        for (int f = 0; f < 5; ++f) U[f] = p.Q[f * p.q_stride + gq];
        Pi.lr = p.Lcur[gq];
        Pi.lb = p.Lcur[p.q_stride + gq];
        if (p.mode == 1) {
#pragma unroll
            for (int f = 0; f < 5; ++f) {
                if (ESDG_HEX_STAGE_RES) cp_async8(sR + (eb * 5 + f) * Nq + i, p.res + f * p.q_stride + gq);
                else prefetch_l2(p.res + f * p.q_stride + gq);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < MAXIT; ++r) {
        const int it = ftid + r * D::THREADS;
        if (it < nb * Nfq) {
            const int b2 = it / Nfq, k = it - b2 * Nfq;
#pragma unroll
            for (int q = 0; q < HEX_TP; ++q)
                cp_async8(sW + (b2 * 7 + q) * Nh + Nq + k, p.Tcur + q * p.t_stride + e0 * Nfq + it);
        }
    }
#pragma unroll
    for (int r = 0; r < MAXIT; ++r) {
        const int it = ftid + r * D::THREADS;
        if (it < nb * Nfq) {
            const int b2 = it / Nfq, k = it - b2 * Nfq;
#pragma unroll
            for (int q = 0; q < HEX_TP; ++q)
                cp_async8(sP + (b2 * 7 + q) * Nfq + k, p.Tcur + q * p.t_stride + ipv[r]);
        }
    }
    for (int it = tid; it < nb * 34; it += D::THREADS) cp_async8(sGeo + it, p.geo + e0 * 34 + it);
    cp_async_commit();
    // (5) warm L2 for the CTA that will run ~one grid-wave later (it starts when this one retires)
    {
        const long long e1 = e0 + (long long)PF_DIST * D::EB;
        if (e1 + D::EB <= p.e_begin + p.e_count) {
            if (tid < D::EB * Nq) {
                const long long g1 = e1 * Nq + tid;
#pragma unroll
                for (int f = 0; f < 5; ++f) prefetch_l2(p.Q + f * p.q_stride + g1);
                prefetch_l2(p.Lcur + g1);
                prefetch_l2(p.Lcur + p.q_stride + g1);
                if (p.mode == 1) {
#pragma unroll
                    for (int f = 0; f < 5; ++f) prefetch_l2(p.res + f * p.q_stride + g1);
                }
            }
#pragma unroll
            for (int r = 0; r < MAXIT; ++r) {
                const int it = ftid + r * D::THREADS;
                if (it < D::EB * Nfq) {
                    prefetch_l2(p.mapP + e1 * Nfq + it);
#pragma unroll
                    for (int q = 0; q < HEX_TP; ++q) prefetch_l2(p.Tcur + q * p.t_stride + e1 * Nfq + it);
                }
            }
        }
    }

    // ---- A: volume nodes -> shared --------------------------------------------------------------
    double* W = sW + eb * 7 * Nh;
    if (vol) {
        vol_prim(U, p.gamma, Pi, rhoe, bad);
        W[i] = Pi.rho;
        W[Nh + i] = Pi.u[0];
        W[2 * Nh + i] = Pi.u[1];
        W[3 * Nh + i] = Pi.u[2];
        W[4 * Nh + i] = Pi.beta;
        W[5 * Nh + i] = Pi.lr;
        W[6 * Nh + i] = Pi.lb;
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- B: surface flux at face points (dg3D_euler_hex.jl:182-197) ----------------------------
    for (int it = ftid; it < nb * Nfq; it += blockDim.x) {
        const int b2 = it / Nfq, k = it - b2 * Nfq;
        const long long ee = e0 + b2;
        const int face = k / nfp;
        const double* gf = sGeo + b2 * 34 + 10 + face * 4;
        const double nJ[3] = {gf[0], gf[1], gf[2]};
        const double sJ = gf[3];
        const Node M = ld_node(sW + b2 * 7 * Nh, Nh, Nq + k);
        const Node P = ld_node(sP + b2 * 7 * Nfq, Nfq, k);
        double F[5];
        pair_flux<false, 0>(M, P, nJ, c_igm, F);
        if (p.lf_scale != 0.0) {
            // conservative states from the primitive traces: E = p/(g-1) + .5 rho |u|^2, p = rho/(2 beta)
            const double isJ = fast_rcp(sJ);
            const double g1 = p.gamma - 1.0;
            double UM[5], UP[5], unM = 0.0, unP = 0.0, kM = 0.0, kP = 0.0;
            UM[0] = M.rho;
            UP[0] = P.rho;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double uM = 2.0 * M.u[d], uP = 2.0 * P.u[d];
                UM[1 + d] = M.rho * uM;
                UP[1 + d] = P.rho * uP;
                kM = fma(UM[1 + d], uM, kM);
                kP = fma(UP[1 + d], uP, kP);
                unM = fma(UM[1 + d], nJ[d], unM);
                unP = fma(UP[1 + d], nJ[d], unP);
            }
            const double pM = .5 * M.rho * fast_rcp(M.beta), pP = .5 * P.rho * fast_rcp(P.beta);
            UM[4] = fma(.5, kM, pM / g1);
            UP[4] = fma(.5, kP, pP / g1);
            const double lamM = wavespeed_1d(UM[0], unM * isJ, UM[4], p.gamma);
            const double lamP = wavespeed_1d(UP[0], unP * isJ, UP[4], p.gamma);
            const double LFc = p.lf_scale * fmax(lamM, lamP) * sJ;
#pragma unroll
            for (int f = 0; f < 5; ++f) F[f] = fma(-LFc, UP[f] - UM[f], F[f]);
        }
#pragma unroll
        for (int f = 0; f < 5; ++f) sFL[(b2 * 7 + f) * Nfq + k] = F[f];
    }

    // ---- C/D: flux differencing, one reference direction at a time ------------------------------
    // The direction loop is rolled (3x smaller code: the unrolled kernel overflowed the instruction
    // cache); velocity components, momentum accumulators, line coordinates and strides are kept in
    // a frame ROTATED with the direction so that "component 0" is always the line direction.
    double acc[5] = {0, 0, 0, 0, 0};
    const double* G = sGeo + eb * 34;
    double* XVb = sX + eb * D::XSZ;
    double* XFb = XVb + D::XV;
    int x0 = xa, x1 = xb, x2 = xc;                 // rotated line coordinates: x0 runs along direction d
    int s0 = 1, s1 = n, s2 = n * n;                // matching node strides
    int m0 = xc + n * xb, m1 = xc + n * xa, m2 = xb + n * xa;   // on-face index of my line per direction
    // work items of the per-direction face reduction (independent of the direction)
    constexpr int MAXH = (D::EB * 10 * nfp + D::THREADS - 1) / D::THREADS;
    int h_xoff[MAXH], h_dst[MAXH];
    bool h_side[MAXH];
#pragma unroll
    for (int r = 0; r < MAXH; ++r) {
        const int it = ftid + r * D::THREADS;
        const int b2 = it / (10 * nfp), q = it - b2 * 10 * nfp;
        const int sf = q / nfp, mm = q - sf * nfp;          // sf = side*5 + field
        h_side[r] = sf >= 5;
        h_xoff[r] = it < nb * 10 * nfp ? b2 * D::XSZ + D::XV + (sf * nfp + mm) * n : -1;
        h_dst[r] = b2 * 7 * Nfq + (sf >= 5 ? nfp : 0) + mm + (sf >= 5 ? sf - 5 : sf) * Nfq;
    }
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
        const int c1 = d == 2 ? 0 : d + 1, c2 = c1 == 2 ? 0 : c1 + 1;   // physical components in rotated order d,c1,c2
        const int pl0 = (1 + d) * Nh, pl1 = (1 + c1) * Nh, pl2 = (1 + c2) * Nh;
        if (vol) {
            // (r_x, r_y, r_z) J of direction d, rotated
            const double g[3] = {G[3 * d + d], G[3 * c1 + d], G[3 * c2 + d]};
            auto ldn = [&](int idx) {
                Node q;
                q.rho = W[idx];
                q.u[0] = W[pl0 + idx];
                q.u[1] = W[pl1 + idx];
                q.u[2] = W[pl2 + idx];
                q.beta = W[4 * Nh + idx];
                q.lr = W[5 * Nh + idx];
                q.lb = W[6 * Nh + idx];
                return q;
            };
            // partners of this direction: KH owned volume pairs (x, x+k mod n) and the two face nodes
            // of my line, evaluated two at a time
            constexpr int NP = KH + 2;
            int pj[NP];
            double pc[NP];
            double* px[NP];          // exchange slot of field 0 (field stride pxs)
            int pxs[NP];
            bool pst[NP];
#pragma unroll
            for (int k = 1; k <= KH; ++k) {
                const bool own = !((n % 2 == 0) && (k == n / 2) && (x0 >= n / 2));
                int ap = x0 + k;
                ap = ap >= n ? ap - n : ap;
                pj[k - 1] = i + (ap - x0) * s0;
                pc[k - 1] = own ? s_T[x0 * n + ap] : 0.0;
                px[k - 1] = XVb + (k - 1) * 5 * Nq + i;
                pxs[k - 1] = Nq;
                pst[k - 1] = own;
            }
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                pj[KH + side] = Nq + (2 * d + side) * nfp + m0;
                pc[KH + side] = side ? s_lpw[x0] : -s_lmw[x0];
                px[KH + side] = XFb + (side * 5 * nfp + m0) * n + x0;
                pxs[KH + side] = nfp * n;
                pst[KH + side] = true;
            }
#pragma unroll
            for (int pp = 0; pp + 1 < NP; pp += 2) {
                const Node B2[2] = {ldn(pj[pp]), ldn(pj[pp + 1])};
                double F2[2][5];
                pair_flux2<DIAG, 0>(Pi, B2, g, c_igm, F2);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
#pragma unroll
                    for (int f = 0; f < 5; ++f) {
                        acc[f] = fma(pc[pp + q], F2[q][f], acc[f]);
                        // face slots are written in PHYSICAL field order (the face reduction is not rotated)
                        const int fs = (pp + q >= KH) ? (f == 1 ? 1 + d : (f == 2 ? 1 + c1 : (f == 3 ? 1 + c2 : f))) : f;
                        if (pst[pp + q]) px[pp + q][fs * pxs[pp + q]] = F2[q][f];
                    }
                }
            }
            if (NP % 2 == 1) {
                const Node Pj = ldn(pj[NP - 1]);
                double F[5];
                pair_flux<DIAG, 0>(Pi, Pj, g, c_igm, F);
#pragma unroll
                for (int f = 0; f < 5; ++f) {
                    acc[f] = fma(pc[NP - 1], F[f], acc[f]);
                    const int fs = (f == 1 ? 1 + d : (f == 2 ? 1 + c1 : (f == 3 ? 1 + c2 : f)));   // last partner is a face
                    if (pst[NP - 1]) px[NP - 1][fs * pxs[NP - 1]] = F[f];
                }
            }
        }
        __syncthreads();
        if (vol) {
#pragma unroll
            for (int k = 1; k <= KH; ++k) {
                int am = x0 - k;
                am = am < 0 ? am + n : am;
                const bool has = !((n % 2 == 0) && (k == n / 2) && (am >= n / 2));
                if (has) {
                    const double t = s_T[x0 * n + am];
                    const int j = i + (am - x0) * s0;
#pragma unroll
                    for (int f = 0; f < 5; ++f) acc[f] = fma(t, XVb[((k - 1) * 5 + f) * Nq + j], acc[f]);
                }
            }
        }
        // face nodes of this direction: H = -nrJ * sum_a l[a] F(a) + flux*, one (face point, field)
        // per work item so that all warps share the reduction (item decode hoisted out of the loop)
#pragma unroll
        for (int r = 0; r < MAXH; ++r) {
            if (h_xoff[r] >= 0) {
                const double* xf = sX + h_xoff[r];
                const double* l = h_side[r] ? s_lp : s_lm;
                double h = 0.0;
#pragma unroll
                for (int a = 0; a < n; ++a) h = fma(l[a], xf[a], h);
                double* dst = sFL + h_dst[r] + 2 * d * nfp;
                *dst = h_side[r] ? *dst - h : *dst + h;
            }
        }
        __syncthreads();
        // rotate the frame: direction d+1 becomes component 0
        {
            const double t = Pi.u[0]; Pi.u[0] = Pi.u[1]; Pi.u[1] = Pi.u[2]; Pi.u[2] = t;
            const double ta = acc[1]; acc[1] = acc[2]; acc[2] = acc[3]; acc[3] = ta;
            const int tx = x0; x0 = x1; x1 = x2; x2 = tx;
            const int ts = s0; s0 = s1; s1 = s2; s2 = ts;
            const int tm = m0; m0 = m1; m1 = m2; m2 = tm;
        }
    }

    // ---- E: lift, scale, rhstest, LSRK stage ----------------------------------------------------
    double tsum = 0.0;
    if (vol) {
        const double* FL = sFL + eb * 7 * Nfq;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int xd = d == 0 ? xa : (d == 1 ? xb : xc);
            const int m = d == 0 ? (xc + n * xb) : (d == 1 ? (xc + n * xa) : (xb + n * xa));
            const double cm = s_lmw[xd], cp = s_lpw[xd];
#pragma unroll
            for (int f = 0; f < 5; ++f) {
                acc[f] = fma(cm, FL[f * Nfq + (2 * d) * nfp + m], acc[f]);
                acc[f] = fma(cp, FL[f * Nfq + (2 * d + 1) * nfp + m], acc[f]);
            }
        }
        const double J = G[9];
        const double mJ = -1.0 / J;
        double r[5];
#pragma unroll
        for (int f = 0; f < 5; ++f) r[f] = acc[f] * mJ;
        if (p.want_rhstest) {
            double V[5];
            entropy_vars(U, rhoe, Pi.lr, Pi.lb, p.gamma, V);
            const double wJ = s_w[xa] * s_w[xb] * s_w[xc] * J;
#pragma unroll
            for (int f = 0; f < 5; ++f) tsum = fma(wJ * V[f], r[f], tsum);
        }
        if (p.mode == 0) {
#pragma unroll
            for (int f = 0; f < 5; ++f) p.rhs[f * p.q_stride + gq] = r[f];
        } else {
            double Un[5];
#pragma unroll
            for (int f = 0; f < 5; ++f) {
                const double rs = fma(p.rk_a, ESDG_HEX_STAGE_RES ? sR[(eb * 5 + f) * Nq + i] : p.res[f * p.q_stride + gq], p.dt * r[f]);
                p.res[f * p.q_stride + gq] = rs;
                Un[f] = fma(p.rk_b, rs, U[f]);
                p.Qout[f * p.q_stride + gq] = Un[f];
            }
            // ---- F (volume part): entropy variables of the new state -> shared -------------------
            Node Pn;
            double rhoen, V[5];
            vol_prim(Un, p.gamma, Pn, rhoen, bad);
            const double lr = log_pos(Pn.rho), lb = log_pos(Pn.beta);
            p.Lnext[gq] = lr;
            p.Lnext[p.q_stride + gq] = lb;
            entropy_vars(Un, rhoen, lr, lb, p.gamma, V);
            double* sV = sX + eb * 5 * Nq;   // element-contiguous V blocks: index (eb*5+f)*Nq + i
#pragma unroll
            for (int f = 0; f < 5; ++f) sV[f * Nq + i] = V[f];
        }
    }
    if (p.mode == 1) {
        __syncthreads();
        write_traces<n>(p, sX, s_lm, s_lp, e0, nb, ftid);
    }
    if (p.want_rhstest) {
        __shared__ double red[32];
        for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(0xffffffffu, tsum, o);
        if ((tid & 31) == 0) red[tid >> 5] = tsum;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int wv = 0; wv < (int)(blockDim.x + 31) / 32; ++wv) t += red[wv];
            p.partial[blockIdx.x] = t;
        }
    }
    if (bad) atomicOr(p.bad, 1);
}

}  // namespace esdg
