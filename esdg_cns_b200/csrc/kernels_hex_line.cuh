// Tensor-product hex stage kernel, LINE-PER-THREAD flux differencing.
//
// Same fused stage as k_hex_stage (kernels_hex.cuh: surface + volume + lift + LSRK + next traces,
// examples/dg3D_euler_hex.jl:167-222 + dg2D_euler_quad.jl:198-207) with a different work
// decomposition of the dominant phase.  The hybridized SBP operator of a Gauss-collocated hex
// couples a volume node only to the nodes of its three grid lines and to the two face points
// each line ends in, so the Hadamard sum  sum_j Q_ij o F(u_i,u_j)  (dg3D_euler_hex.jl:122-164)
// splits into 3*n^2 independent LINE problems per element:
//     n volume nodes + 2 face nodes  ->  n(n-1)/2 + 2n unordered pairs.
// One thread owns one line: it loads the n+2 nodes ONCE into registers, evaluates every unordered
// pair exactly once (the symmetric loop of CompressibleNS/dg2D_CNS_cavity_optimized.jl:326-347),
// keeps all n*5 + 2*5 accumulators in registers and writes each result once.  Compared with the
// node-per-thread kernel this removes the shared-memory exchange of pair fluxes, all per-pair
// index arithmetic (the pair table is compile-time; T[a][b] are constant-bank operands of the
// DFMAs) and 4 of the 9 block barriers.  Every phase processes TWO independent items per thread
// statement-interleaved (vecmath.cuh), because occupancy is 2.5 warps per scheduler and the FP64
// pipe must be fed from instruction-level parallelism.
//
// Phases per CTA (EB elements):
//   0  cp.async: Q, LSRK register, log stash, own traces, exterior traces (mapP gather), geometry
//   A  volume primitives (rho, u/2, beta) -> shared
//   B  surface flux F_EC(M,P).nJ - LFc (U+ - U)                       dg3D_euler_hex.jl:182-197
//   C  line problems; face rows get  flux* -/+ sum_a l[a] F(a,face)   (Ph / Lf as 1-D sweeps)
//   E  rhs = -(sum_d acc_d + lift)/J ; rhstest ; res = a res + dt rhs ; Q += b res
//   F  entropy projection of the new state -> traces + log stash for the next stage   (:175-178)
#pragma once
#include "kernels_hex.cuh"
#include "vecmath.cuh"

namespace esdg {

#ifdef ESDG_PHASE_TIMING   // debug build only: per-phase cycle counts summed over CTAs (thread 0 of each CTA)
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_MARK(k)                                                                      \
    do {                                                                                   \
        if (threadIdx.x == 0) {                                                            \
            const long long t_ = clock64();                                                \
            atomicAdd(&g_phase_cycles[k], (unsigned long long)(t_ - t_phase));             \
            t_phase = t_;                                                                  \
        }                                                                                  \
    } while (0)
#else
#define PHASE_MARK(k)
#endif

// ---- TMA bulk copy (cp.async.bulk, completion on an mbarrier): one instruction moves a whole plane ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(double* smem_dst, const double* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int n>
struct LineDims {
    static constexpr int Nq = n * n * n, nfp = n * n, Nfq = 6 * nfp, Nh = Nq + Nfq, NL = 3 * nfp;
#ifndef ESDG_LINE_EB5
#define ESDG_LINE_EB5 2
#endif
#ifndef ESDG_LINE_EB4
#define ESDG_LINE_EB4 2
#endif
    static constexpr int EB = n == 5 ? ESDG_LINE_EB5 : (n == 4 ? ESDG_LINE_EB4 : (n == 3 ? 4 : 8));
    static constexpr int THREADS = ((EB * NL + 31) / 32) * 32;   // line threads
#ifndef ESDG_LINE_HELPER
#define ESDG_LINE_HELPER 0
#endif
    // N = 3, 4: an odd number of line warps leaves one warp slot of the register file unused (warps are allocated in
    // pairs); with ESDG_LINE_HELPER=1 a helper warp in that slot computes the surface flux concurrently with the pair
    // loops (mbarrier hand-over).  Measured on B200: N=4 5.11 ms against 5.02 ms without it, N=3 2.57 against 2.34 --
    // one warp needs as long for all the face points as the line warps for their pairs, so it stays off.
#ifndef ESDG_LINE_HELPER4
#define ESDG_LINE_HELPER4 1
#endif
    static constexpr bool HELPER = ESDG_LINE_HELPER && ((THREADS / 32) % 2 == 1) && (n == 5 || ESDG_LINE_HELPER4);
    static constexpr int THREADS_LAUNCH = THREADS + (HELPER ? 32 : 0);
#ifndef ESDG_LINE_MINB5
#define ESDG_LINE_MINB5 2
#endif
#ifndef ESDG_LINE_MINB4
#define ESDG_LINE_MINB4 4
#endif
    static constexpr int MINB = (n == 5 && ESDG_LINE_EB5 == 2) ? ESDG_LINE_MINB5 : (n == 4 ? ESDG_LINE_MINB4 : 4);   // CTAs per SM the register budget is sized for
    static constexpr int EN = EB * Nq, EF = EB * Nfq;
    static constexpr int TAB = 5 * n + EB * 34;
    // WV[7][EN] WF[7][EF] P[7][EF] A[15][EN] Q[5][EN] R[5][EN]
    static constexpr int SMEM_DOUBLES = 7 * EN + 14 * EF + 15 * EN + 10 * EN + TAB;
};

// unordered pairs (a,b), a < b, of one line in a fixed order
template <int n>
struct PairTab {
    static constexpr int NP = n * (n - 1) / 2;
    int a[NP > 0 ? NP : 1], b[NP > 0 ? NP : 1];
    constexpr PairTab() : a(), b() {
        int k = 0;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) { a[k] = i; b[k] = j; ++k; }
    }
};

template <int W>
struct NodeV {
    VD<W> rho, u[3], beta, lr, lb;
};

template <int W>
__device__ __forceinline__ void set_lane(NodeV<W>& d, int q, const Node& s) {
    d.rho[q] = s.rho; d.u[0][q] = s.u[0]; d.u[1][q] = s.u[1]; d.u[2][q] = s.u[2];
    d.beta[q] = s.beta; d.lr[q] = s.lr; d.lb[q] = s.lb;
}

// pair_flux (kernels_hex.cuh) on W independent pairs: EC flux (euler_fluxes.jl:51-89, logmean.jl:14-28)
// contracted with g; one reciprocal per pair serves rholog, 1/betalog and pa.
// Instruction diet (all within an ulp of the reference formulas):
//  * sums instead of means: with s = aL + aR the Taylor argument is f/2 = da/s, f^2 = 4 (da/s)^2, and the
//    branch test |da| < 1e-4 (s/2) is the same predicate written as |da| < 5e-5 s (exact scaling by 2);
//  * in the Taylor branch |f| < 1e-4, so f^4 terms of logmean.jl:18-27 are < 5.2e-18, a twentieth of an ulp of
//    the leading 1: 1 - .2 f^2 is the reference polynomial to double precision;
//  * DIAG (axis-aligned elements): every flux component is proportional to the single metric term g[0]; it is
//    factored out here and applied once to the line's accumulated sums (GSCALED = false: F is F/g[0]).
template <int W, bool DIAG>
__device__ __forceinline__ void pair_flux_v(const NodeV<W>& A, const NodeV<W>& B, const VD<W>* g, double c_igm, VD<W>* F) {
    const VD<W> dar = B.rho - A.rho, sr = A.rho + B.rho;
    const VD<W> dab = B.beta - A.beta, sb = A.beta + B.beta;
    const VB<W> tr = vlt(vabs(dar), 5e-5 * sr), tb = vlt(vabs(dab), 5e-5 * sb);
    const VD<W> X = vsel(tr, sr, B.lr - A.lr);
    const VD<W> Y = vsel(tb, sb, dab);
    const VD<W> XY = X * Y;
    const VD<W> R = vrcp(XY * sb);
    const VD<W> iX = R * (Y * sb), iY = R * (X * sb), iZ = R * XY;
    const VD<W> qr = dar * iX;                     // f/2 (Taylor branch) or the log mean itself
    const VD<W> fb = dab * iY;
    const VD<W> vr = qr * qr, vb = fb * fb;        // f^2 / 4
    const VD<W> rholog = vsel(tr, sr * vfma(vr, -.4, .5), qr);                    // (s/2)(1 - .2 f^2)
    const VD<W> ibl = iY * vsel(tb, vfma(vb, 1.6, 2.0), B.lb - A.lb);             // 1/betalog = (2/s)(1 + .2 f^2)
    const VD<W> pa = (.5 * sr) * iZ;                                              // rhoavg/(betaL+betaR)
    VD<W> uavg[3], unorm, un;
    uavg[0] = A.u[0] + B.u[0];                                 // Node.u holds HALF velocities
    uavg[1] = A.u[1] + B.u[1];
    uavg[2] = A.u[2] + B.u[2];
    unorm = A.u[0] * B.u[0];
    unorm = vfma(A.u[1], B.u[1], unorm);
    unorm = vfma(A.u[2], B.u[2], unorm);                       // = uL.uR / 4
    const VD<W> Ep = vfma(rholog, vfma(unorm, 2.0, ibl * c_igm), pa);   // rholog/(2(g-1)betalog)+pa+.5 rholog uL.uR
    if (DIAG) {
        const VD<W> m = rholog * uavg[0];                      // fluxes / g[0]
        F[0] = m;
        F[1] = vfma(m, uavg[0], pa);
        F[2] = m * uavg[1];
        F[3] = m * uavg[2];
        F[4] = Ep * uavg[0];
    } else {
        un = g[0] * uavg[0];
        un = vfma(g[1], uavg[1], un);
        un = vfma(g[2], uavg[2], un);
        const VD<W> m = rholog * un;
        F[0] = m;
        F[1] = vfma(m, uavg[0], pa * g[0]);
        F[2] = vfma(m, uavg[1], pa * g[1]);
        F[3] = vfma(m, uavg[2], pa * g[2]);
        F[4] = Ep * un;
    }
}

// log_pos / exp_mod of kernels_hex.cuh on W lanes
template <int W>
__device__ __forceinline__ VD<W> vlog_pos(const VD<W>& x) {
    VD<W> f, dk;
    VFOR {
        int hi = __double2hiint(x.v[q_]);
        const int lo = __double2loint(x.v[q_]);
        int k = (hi >> 20) - 1023;
        hi &= 0x000fffff;
        const int i = (hi + 0x95f64) & 0x100000;
        hi |= (i ^ 0x3ff00000);
        k += (i >> 20);
        f.v[q_] = __hiloint2double(hi, lo) - 1.0;
        dk.v[q_] = (double)k;
    }
    const VD<W> s = f * vrcp(f + 2.0);
    const VD<W> z = s * s, w = z * z;
    const VD<W> t1 = w * vfma(w, vfma(w, c_lg[5], c_lg[3]), c_lg[1]);
    const VD<W> t2 = z * vfma(w, vfma(w, vfma(w, c_lg[6], c_lg[4]), c_lg[2]), c_lg[0]);
    const VD<W> R = t2 + t1;
    const VD<W> hfsq = .5 * f * f;
    return dk * 6.93147180369123816490e-01 - ((hfsq - vfma(s, hfsq + R, dk * 1.90821492927058770002e-10)) - f);
}

template <int W>
__device__ __forceinline__ VD<W> vexp_mod(const VD<W>& x) {
    VD<W> kd;
    VFOR kd.v[q_] = rint(x.v[q_] * 1.44269504088896338700e+00);
    const VD<W> hi = vfma(kd, -6.93147180369123816490e-01, x);
    const VD<W> lo = kd * 1.90821492927058770002e-10;
    const VD<W> r = hi - lo;
    const VD<W> t = r * r;
    const VD<W> c = r - t * vfma(t, vfma(t, vfma(t, vfma(t, c_ex[4], c_ex[3]), c_ex[2]), c_ex[1]), c_ex[0]);
    const VD<W> y = 1.0 - ((lo - (r * c) * vrcp(2.0 - c)) - hi);
    VD<W> out;
    VFOR out.v[q_] = __hiloint2double(__double2hiint(y.v[q_]) + ((int)kd.v[q_] << 20), __double2loint(y.v[q_]));
    return out;
}

// vol_prim / entropy_vars / trace_from_entropy of kernels_hex.cuh on W lanes
template <int W>
__device__ __forceinline__ void vol_prim_v(const VD<W>* U, double gamma, VD<W>& rho, VD<W>* hu, VD<W>& beta, VD<W>& rhoe) {
    const VD<W> hir = .5 * vrcp(U[0]);
    hu[0] = U[1] * hir;
    hu[1] = U[2] * hir;
    hu[2] = U[3] * hir;
    VD<W> ke = U[1] * hu[0];
    ke = vfma(U[2], hu[1], ke);
    ke = vfma(U[3], hu[2], ke);                 // = .5 |rhoU|^2 / rho
    rhoe = U[4] - ke;
    rho = U[0];
    beta = .5 * U[0] * vrcp((gamma - 1.0) * rhoe);
}

template <int W>
__device__ __forceinline__ void entropy_vars_v(const VD<W>* U, const VD<W>& rhoe, const VD<W>& lr, const VD<W>& lb, double gamma, VD<W>* V) {
    const VD<W> s = vfma(lr, 1.0 - gamma, -lb) - 0.6931471805599453094;
    const VD<W> ire = vrcp(rhoe);
    V[0] = vfma(-U[4], ire, (gamma + 1.0) - s);
    V[1] = U[1] * ire;
    V[2] = U[2] * ire;
    V[3] = U[3] * ire;
    V[4] = -(U[0] * ire);
}

template <int W>
__device__ __forceinline__ void trace_from_entropy_v(const VD<W>* V, double gamma, double lg1, double l2g1, VD<W>* Tv) {
    const VD<W> nvE = -V[4];
    const VD<W> inv = vrcp(nvE);
    const VD<W> v2 = vfma(V[1], V[1], vfma(V[2], V[2], V[3] * V[3]));
    const VD<W> s = (gamma - V[0]) - .5 * v2 * inv;
    const VD<W> lv = vlog_pos(nvE);
    const double ig = 1.0 / (gamma - 1.0);
    const VD<W> lrhoe = ((lg1 - gamma * lv) - s) * ig;
    const VD<W> rhoeV = vexp_mod(lrhoe);
    Tv[0] = rhoeV * nvE;
    const VD<W> hinv = .5 * inv;
    Tv[1] = V[1] * hinv;
    Tv[2] = V[2] * hinv;
    Tv[3] = V[3] * hinv;
    Tv[4] = (.5 * ig) * nvE;
    Tv[5] = lrhoe + lv;
    Tv[6] = lv - l2g1;
}

// |wavespeed(rho, rhoU_n, E)|, euler_variables.jl:7-10 (sic: sqrt of |u_n|), ir = 1/rho
template <int W>
__device__ __forceinline__ VD<W> wavespeed_v(const VD<W>& ir, const VD<W>& rhoun, const VD<W>& E, double gamma) {
    const VD<W> pg = (gamma * (gamma - 1.0)) * (E - .5 * rhoun * rhoun * ir);
    return vabs(vsqrt(vabs(rhoun * ir)) + vsqrt(pg * ir));
}

// W pairs of one line evaluated together (statement-interleaved): volume-volume pairs K0..K0+W-1 of PairTab<n>
template <int n, bool DIAG, int K0, int W>
__device__ __forceinline__ void line_vv_group(const Node (&nd)[n], double (&acc)[n][5], double g0, double g1, double g2, double c_igm,
                                              const HexParams& p) {
    constexpr PairTab<n> PT{};
    const VD<W> g[3] = {vset<W>(g0), vset<W>(g1), vset<W>(g2)};
    NodeV<W> A, B;
#pragma unroll
    for (int q = 0; q < W; ++q) {
        set_lane(A, q, nd[PT.a[K0 + q]]);
        set_lane(B, q, nd[PT.b[K0 + q]]);
    }
    VD<W> F[5];
    pair_flux_v<W, DIAG>(A, B, g, c_igm, F);
#pragma unroll
    for (int q = 0; q < W; ++q) {
        const int a = PT.a[K0 + q], b = PT.b[K0 + q];
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            acc[a][f] = fma(p.T[a * n + b], F[f].v[q], acc[a][f]);
            acc[b][f] = fma(p.T[b * n + a], F[f].v[q], acc[b][f]);
        }
    }
}
template <int n, bool DIAG, int K0, int WC>
__device__ __forceinline__ void line_vv_all(const Node (&nd)[n], double (&acc)[n][5], double g0, double g1, double g2, double c_igm,
                                            const HexParams& p) {
    constexpr int NP = PairTab<n>::NP;
    if constexpr (K0 < NP) {
        constexpr int W = NP - K0 < WC ? NP - K0 : WC;
        line_vv_group<n, DIAG, K0, W>(nd, acc, g0, g1, g2, c_igm, p);
        line_vv_all<n, DIAG, K0 + W, WC>(nd, acc, g0, g1, g2, c_igm, p);
    }
}
// volume-face pairs J0..J0+W-1, item j = (node j/2, side j%2)
template <int n, bool DIAG, int J0, int W>
__device__ __forceinline__ void line_vf_group(const Node (&nd)[n], const Node (&fn)[2], double (&acc)[n][5], double (&h)[2][5], double g0,
                                              double g1, double g2, double c_igm, const HexParams& p) {
    const VD<W> g[3] = {vset<W>(g0), vset<W>(g1), vset<W>(g2)};
    NodeV<W> A, B;
#pragma unroll
    for (int q = 0; q < W; ++q) {
        set_lane(A, q, nd[(J0 + q) / 2]);
        set_lane(B, q, fn[(J0 + q) % 2]);
    }
    VD<W> F[5];
    pair_flux_v<W, DIAG>(A, B, g, c_igm, F);
#pragma unroll
    for (int q = 0; q < W; ++q) {
        const int a = (J0 + q) / 2, side = (J0 + q) % 2;
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            acc[a][f] = fma(side ? p.lpw[a] : -p.lmw[a], F[f].v[q], acc[a][f]);
            h[side][f] = fma(side ? p.lp[a] : p.lm[a], F[f].v[q], h[side][f]);
        }
    }
}
template <int n, bool DIAG, int J0, int WC>
__device__ __forceinline__ void line_vf_all(const Node (&nd)[n], const Node (&fn)[2], double (&acc)[n][5], double (&h)[2][5], double g0,
                                            double g1, double g2, double c_igm, const HexParams& p) {
    if constexpr (J0 < 2 * n) {
        constexpr int W = 2 * n - J0 < WC ? 2 * n - J0 : WC;
        line_vf_group<n, DIAG, J0, W>(nd, fn, acc, h, g0, g1, g2, c_igm, p);
        line_vf_all<n, DIAG, J0 + W, WC>(nd, fn, acc, h, g0, g1, g2, c_igm, p);
    }
}

#ifdef ESDG_LINE_MAXNREG   // register-budget experiments (n = 5 only): explicit cap instead of the CTAs-per-SM hint
#define ESDG_LINE_BOUNDS(n) __launch_bounds__(LineDims<n>::THREADS_LAUNCH) __maxnreg__(n == 5 ? ESDG_LINE_MAXNREG : 128)
#else
#ifndef ESDG_LINE_REGTHREADS4
#define ESDG_LINE_REGTHREADS4 0   // N = 3: thread count the register budget is computed for (0 = the launch size)
#endif
#define ESDG_LINE_BOUNDS(n) \
    __launch_bounds__((n == 4 && ESDG_LINE_REGTHREADS4) ? ESDG_LINE_REGTHREADS4 : LineDims<n>::THREADS_LAUNCH, LineDims<n>::MINB)
#endif

template <int n, bool DIAG>
__global__ void ESDG_LINE_BOUNDS(n) k_hex_stage_line(
    const __grid_constant__ HexParams p) {
    using D = LineDims<n>;
    constexpr int Nq = D::Nq, nfp = D::nfp, Nfq = D::Nfq, NL = D::NL, EB = D::EB, NT = D::THREADS;
    constexpr int EN = D::EN, EF = D::EF;
#ifndef ESDG_LINE_WI
#define ESDG_LINE_WI 1
#endif
    constexpr int W = ESDG_LINE_WI;   // items per thread interleaved in the per-node / per-face-point phases (2 measured no faster)
    constexpr int NTP = D::THREADS_LAUNCH;   // all threads share the per-point phases and the loads
    extern __shared__ __align__(16) double sm[];
    __shared__ unsigned long long mbar, bbar;
    double* sWV = sm;                   // [7][EN]  rho, u/2, v/2, w/2, beta, log rho, log beta at volume nodes
    double* sWF = sWV + 7 * EN;         // [7][EF]  the same at the face points (own traces)
    double* sP = sWF + 7 * EF;          // [7][EF]  exterior traces, overwritten by the surface flux (5 planes)
    double* sA = sP + 7 * EF;           // [3][5][EN] per-direction volume accumulators; later V[5][EN]
    double* sQ = sA + 15 * EN;          // [5][EN]
    double* sR = sQ + 5 * EN;           // [5][EN]  LSRK register (mode 1)
    double* s_lm = sR + 5 * EN;
    double* s_lp = s_lm + n;
    double* s_lmw = s_lp + n;
    double* s_lpw = s_lmw + n;
    double* s_w = s_lpw + n;
    double* sGeo = s_w + n;             // [EB][34]: G[9], J, 6 x (nxJ,nyJ,nzJ,sJ)
    const int tid = threadIdx.x;
#ifdef ESDG_PHASE_TIMING
    long long t_phase = clock64();
#endif
    const long long e0 = p.e_begin + (long long)blockIdx.x * EB;
    const int nb = (int)min((long long)EB, p.e_begin + p.e_count - e0);
    constexpr int MAXF = (EF + NTP - 1) / NTP, MAXV = (EN + NTP - 1) / NTP;
    constexpr int GF = (MAXF + W - 1) / W, GV = (MAXV + W - 1) / W;   // item groups of W per thread
    const long long gq0 = e0 * Nq, gf0 = e0 * Nfq;
    // exterior-trace indices first: their (L2) latency overlaps everything up to the gather below
    int ipv[MAXF];
#pragma unroll
    for (int r = 0; r < MAXF; ++r) {
        const int it = tid + r * NTP;
#ifndef ESDG_LINE_NOLOAD
        ipv[r] = it < nb * Nfq ? p.mapP[gf0 + it] : -1;
#else
        ipv[r] = -1;
#endif
    }
    PHASE_MARK(12);
    if (tid == 0) {
        mbar_init(&mbar, 1);
        mbar_init(&bbar, 32);   // helper warp -> line threads: surface flux ready
    }
    PHASE_MARK(13);
    if (tid < n) {
        s_lm[tid] = p.lm[tid];
        s_lp[tid] = p.lp[tid];
        s_lmw[tid] = p.lmw[tid];
        s_lpw[tid] = p.lpw[tid];
        s_w[tid] = p.w[tid];
    }
    __syncthreads();   // mbarrier initialised before any warp issues a bulk copy on it
    PHASE_MARK(14);
    const double c_igm = 1.0 / (2.0 * (p.gamma - 1.0));
    bool bad = false;

    // ---- 0: every global read of this batch goes in flight -----------------------------------------
    // Contiguous planes (Q, LSRK register, log stash, own traces: 19 runs of 2 - 2.4 KB) move as TMA bulk
    // copies issued by one thread; only the exterior-trace gather through mapP needs per-point cp.async.
    // 16-byte alignment of every run (uniform per CTA); otherwise the 8-byte cp.async path
    const bool bulk = (((p.q_stride | gq0 | (long long)(nb * Nq) | p.t_stride | gf0 | (long long)(nb * Nfq)) & 1) == 0) &&
                      ((((size_t)p.Q | (size_t)p.res | (size_t)p.Lcur | (size_t)p.Tcur) & 15) == 0);
#ifndef ESDG_LINE_NOLOAD   // NOLOAD: upper-bound experiment without any global read (garbage in shared memory)
    if (bulk) {
        // 19 (mode 1) / 14 runs, one per lane-0 of the warps in turn: TMA issue costs ~90 cycles per copy when
        // one thread issues them back to back
        const unsigned bq = (unsigned)(nb * Nq * 8), bf = (unsigned)(nb * Nfq * 8);
        if (tid == 0) mbar_expect_tx(&mbar, (p.mode == 1 ? 12u : 7u) * bq + 7u * bf);
        if ((tid & 31) == 0) {
            const int wv = tid >> 5;
            constexpr int NW = NTP / 32;
            for (int c = wv; c < 19; c += NW) {
                if (c < 5) bulk_g2s(sQ + c * EN, p.Q + c * p.q_stride + gq0, bq, &mbar);
                else if (c < 7) bulk_g2s(sWV + c * EN, p.Lcur + (c - 5) * p.q_stride + gq0, bq, &mbar);
                else if (c < 14) bulk_g2s(sWF + (c - 7) * EF, p.Tcur + (c - 7) * p.t_stride + gf0, bf, &mbar);
                else if (p.mode == 1) bulk_g2s(sR + (c - 14) * EN, p.res + (c - 14) * p.q_stride + gq0, bq, &mbar);
            }
        }
    } else {
        for (int it = tid; it < nb * Nq; it += NTP) {
            const long long g = gq0 + it;
#pragma unroll
            for (int f = 0; f < 5; ++f) cp_async8(sQ + f * EN + it, p.Q + f * p.q_stride + g);
            cp_async8(sWV + 5 * EN + it, p.Lcur + g);
            cp_async8(sWV + 6 * EN + it, p.Lcur + p.q_stride + g);
            if (p.mode == 1) {
#pragma unroll
                for (int f = 0; f < 5; ++f) cp_async8(sR + f * EN + it, p.res + f * p.q_stride + g);
            }
        }
        for (int it = tid; it < nb * Nfq; it += NTP) {
#pragma unroll
            for (int q = 0; q < HEX_TP; ++q) cp_async8(sWF + q * EF + it, p.Tcur + q * p.t_stride + gf0 + it);
        }
    }
    PHASE_MARK(10);
#pragma unroll
    for (int r = 0; r < MAXF; ++r) {
        const int it = tid + r * NTP;
        if (it < nb * Nfq) {
#pragma unroll
            for (int q = 0; q < HEX_TP; ++q) cp_async8(sP + q * EF + it, p.Tcur + q * p.t_stride + ipv[r]);
        }
    }
    for (int it = tid; it < nb * 34; it += NTP) cp_async8(sGeo + it, p.geo + e0 * 34 + it);
    cp_async_commit();
    PHASE_MARK(11);
    // warm L2 for the CTA that starts when this one retires (one grid wave ahead)
    {
        constexpr int PF_DIST = 148 * D::MINB;
        const long long e1 = e0 + (long long)PF_DIST * EB;
        if (e1 + EB <= p.e_begin + p.e_count) {
            for (int it = tid * 16; it < EN; it += NTP * 16) {   // one 128-byte line per 16 doubles
                const long long g1 = e1 * Nq + it;
#pragma unroll
                for (int f = 0; f < 5; ++f) prefetch_l2(p.Q + f * p.q_stride + g1);
                prefetch_l2(p.Lcur + g1);
                prefetch_l2(p.Lcur + p.q_stride + g1);
                if (p.mode == 1) {
#pragma unroll
                    for (int f = 0; f < 5; ++f) prefetch_l2(p.res + f * p.q_stride + g1);
                }
            }
            for (int it = tid * 16; it < EF; it += NTP * 16) {
                prefetch_l2(p.mapP + e1 * Nfq + it);
#pragma unroll
                for (int q = 0; q < HEX_TP; ++q) prefetch_l2(p.Tcur + q * p.t_stride + e1 * Nfq + it);
            }
        }
    }
#endif
    PHASE_MARK(0);
    cp_async_wait_all();
    __syncthreads();
#ifndef ESDG_LINE_NOLOAD
    if (bulk) mbar_wait(&mbar, 0);
#endif
    PHASE_MARK(1);

#define LINE_CSYNC() __syncthreads()
#define LINE_SGEO sGeo
#define LINE_PARTIAL_IDX blockIdx.x
#define LINE_AFTER_C()
#define LINE_AFTER_E()
#define LINE_NTP NTP
#define LINE_HELPER D::HELPER
#define LINE_B_PARITY 0
#include "hex_line_phases.inc"
#undef LINE_CSYNC
#undef LINE_SGEO
#undef LINE_PARTIAL_IDX
#undef LINE_AFTER_C
#undef LINE_AFTER_E
#undef LINE_NTP
#undef LINE_HELPER
#undef LINE_B_PARITY
#ifndef ESDG_LINE_NOLOAD
    if (bad) atomicOr(p.bad, 1);
#endif
}

// ---- persistent variant with a producer warp -----------------------------------------------------------
// The one-shot kernel spends 16 % of its time issuing and awaiting its own loads (measured with ESDG_LINE_NOLOAD),
// and the register file leaves a sixth warp per CTA unused (5 compute warps are allocated as 6).  Here every CTA
// stays resident and walks over its batches; warp 5 is a PRODUCER that fetches the next batch into the shared
// arrays the compute warps have released, so loads overlap phases E/F of the previous batch:
//   full_T  : own traces (7 TMA bulk copies), log stash (2), exterior traces (mapP gather, cp.async), geometry
//   full_Q  : Q (5 TMA bulk copies) and the LSRK register (5)
//   empty_T : compute threads arrive after the line phase (traces, exterior traces / flux*, stash consumed)
//   empty_Q : compute threads arrive after the update phase
// Full batches with 16-byte aligned runs only (the host sends a ragged tail to the one-shot kernel).

template <int n>
struct LineDimsP {
    using D = LineDims<n>;
    static constexpr int THREADS = D::THREADS + 32;                      // compute warps + the producer warp
    static constexpr int SMEM_DOUBLES = D::SMEM_DOUBLES + D::EB * 34;    // geometry records double-buffered
};

template <int n, bool DIAG>
__global__ void __launch_bounds__(LineDimsP<n>::THREADS, LineDims<n>::MINB) k_hex_stage_line_p(const __grid_constant__ HexParams p) {
    using D = LineDims<n>;
    constexpr int Nq = D::Nq, nfp = D::nfp, Nfq = D::Nfq, NL = D::NL, EB = D::EB, NT = D::THREADS;
    constexpr int EN = D::EN, EF = D::EF;
    constexpr int W = ESDG_LINE_WI;
    extern __shared__ __align__(16) double sm[];
    __shared__ unsigned long long pbar[4];   // full_T, full_Q, empty_T, empty_Q
    __shared__ unsigned long long bbar;      // (helper-warp hand-over of the one-shot kernel: unused here)
    double* sWV = sm;
    double* sWF = sWV + 7 * EN;
    double* sP = sWF + 7 * EF;
    double* sA = sP + 7 * EF;
    double* sQ = sA + 15 * EN;
    double* sR = sQ + 5 * EN;
    double* s_lm = sR + 5 * EN;
    double* s_lp = s_lm + n;
    double* s_lmw = s_lp + n;
    double* s_lpw = s_lmw + n;
    double* s_w = s_lpw + n;
    double* sGeo2 = s_w + n;            // [2][EB][34]
    const int tid = threadIdx.x;
#ifdef ESDG_PHASE_TIMING
    long long t_phase = clock64();
#endif
    if (tid == 0) {
        mbar_init(&pbar[0], 33);        // expect_tx arrive of producer lane 0 + one cp.async arrive per producer lane
        mbar_init(&pbar[1], 1);
        mbar_init(&pbar[2], NT);
        mbar_init(&pbar[3], NT);
    }
    if (tid < n) {
        s_lm[tid] = p.lm[tid];
        s_lp[tid] = p.lp[tid];
        s_lmw[tid] = p.lmw[tid];
        s_lpw[tid] = p.lpw[tid];
        s_w[tid] = p.w[tid];
    }
    __syncthreads();
    const long long nbatch = p.e_count / EB;
    constexpr unsigned bq = (unsigned)(EB * Nq * 8), bf = (unsigned)(EB * Nfq * 8);

    if (tid >= NT) {
        // ---- producer warp ------------------------------------------------------------------------
        const int lane = tid - NT;
        constexpr int MAXFP = (EF + 31) / 32;
        unsigned it = 0;
        for (long long bi = blockIdx.x; bi < nbatch; bi += gridDim.x, ++it) {
            const long long e0 = p.e_begin + bi * EB;
            const long long gq0 = e0 * Nq, gf0 = e0 * Nfq;
            int ipv[MAXFP];
#pragma unroll
            for (int r = 0; r < MAXFP; ++r) {
                const int k = lane + r * 32;
                ipv[r] = k < EF ? p.mapP[gf0 + k] : -1;
            }
            if (it > 0) mbar_wait(&pbar[2], (it - 1) & 1);
            if (lane == 0) mbar_expect_tx(&pbar[0], 7u * bf + 2u * bq);
            __syncwarp();
            if (lane < 7) bulk_g2s(sWF + lane * EF, p.Tcur + lane * p.t_stride + gf0, bf, &pbar[0]);
            else if (lane < 9) bulk_g2s(sWV + (lane - 2) * EN, p.Lcur + (lane - 7) * p.q_stride + gq0, bq, &pbar[0]);
#pragma unroll
            for (int r = 0; r < MAXFP; ++r) {
                const int k = lane + r * 32;
                if (k < EF) {
#pragma unroll
                    for (int q = 0; q < HEX_TP; ++q) cp_async8(sP + q * EF + k, p.Tcur + q * p.t_stride + ipv[r]);
                }
            }
            double* sG = sGeo2 + (it & 1) * (EB * 34);
            for (int k = lane; k < EB * 34; k += 32) cp_async8(sG + k, p.geo + e0 * 34 + k);
            cp_async_mbar_arrive(&pbar[0]);
            if (it > 0) mbar_wait(&pbar[3], (it - 1) & 1);
            if (lane == 0) mbar_expect_tx(&pbar[1], (p.mode == 1 ? 10u : 5u) * bq);
            __syncwarp();
            if (lane < 5) bulk_g2s(sQ + lane * EN, p.Q + lane * p.q_stride + gq0, bq, &pbar[1]);
            else if (lane < 10 && p.mode == 1) bulk_g2s(sR + (lane - 5) * EN, p.res + (lane - 5) * p.q_stride + gq0, bq, &pbar[1]);
            // warm L2 for this CTA's next batch (fetched one batch time from now)
            const long long b1 = bi + gridDim.x;
            if (b1 < nbatch) {
                const long long e1 = p.e_begin + b1 * EB;
                for (int k = lane * 16; k < EN; k += 32 * 16) {
                    const long long g1 = e1 * Nq + k;
#pragma unroll
                    for (int f = 0; f < 5; ++f) prefetch_l2(p.Q + f * p.q_stride + g1);
                    prefetch_l2(p.Lcur + g1);
                    prefetch_l2(p.Lcur + p.q_stride + g1);
                    if (p.mode == 1) {
#pragma unroll
                        for (int f = 0; f < 5; ++f) prefetch_l2(p.res + f * p.q_stride + g1);
                    }
                }
                for (int k = lane * 16; k < EF; k += 32 * 16) {
                    prefetch_l2(p.mapP + e1 * Nfq + k);
#pragma unroll
                    for (int q = 0; q < HEX_TP; ++q) prefetch_l2(p.Tcur + q * p.t_stride + e1 * Nfq + k);
                }
            }
        }
        return;
    }

    // ---- compute warps ------------------------------------------------------------------------------
    constexpr int MAXF = (EF + NT - 1) / NT, MAXV = (EN + NT - 1) / NT;
    constexpr int GF = (MAXF + W - 1) / W, GV = (MAXV + W - 1) / W;
    const double c_igm = 1.0 / (2.0 * (p.gamma - 1.0));
    bool bad = false;
    constexpr int nb = EB;
    unsigned iter = 0;
    for (long long bi = blockIdx.x; bi < nbatch; bi += gridDim.x, ++iter) {
        const long long e0 = p.e_begin + bi * EB;
        const double* sGeoC = sGeo2 + (iter & 1) * (EB * 34);
        mbar_wait(&pbar[1], iter & 1);    // Q (phase A) and the LSRK register
        mbar_wait(&pbar[0], iter & 1);    // traces, exterior traces, log stash, geometry
        PHASE_MARK(1);
#define LINE_CSYNC() asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory")
#define LINE_SGEO sGeoC
#define LINE_PARTIAL_IDX bi
#define LINE_AFTER_C() do { if (!p.want_rhstest) mbar_arrive(&pbar[2]); } while (0)
#define LINE_AFTER_E() do { if (p.want_rhstest) mbar_arrive(&pbar[2]); mbar_arrive(&pbar[3]); } while (0)
#define LINE_NTP NT
#define LINE_HELPER false
#define LINE_B_PARITY 0
#include "hex_line_phases.inc"
#undef LINE_CSYNC
#undef LINE_SGEO
#undef LINE_PARTIAL_IDX
#undef LINE_AFTER_C
#undef LINE_AFTER_E
#undef LINE_NTP
#undef LINE_HELPER
#undef LINE_B_PARITY
    }
    if (bad) atomicOr(p.bad, 1);
}

}  // namespace esdg
