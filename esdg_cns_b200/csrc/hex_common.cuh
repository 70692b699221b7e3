// Tensor-product hex path (collocated Gauss nodes, affine elements; BASELINE configs 4 / 5): shared device code.
//
// One launch per LSRK stage does everything the reference's rhs() + stage update do
// (examples/dg3D_euler_hex.jl:167-222, dg2D_euler_quad.jl:198-207).  The hybridized SBP operator of a
// Gauss-collocated hex couples a volume node only to the nodes of its three grid lines and to the two face
// points each line ends in, so the Hadamard sum  sum_j Q_ij o F(u_i,u_j)  (dg3D_euler_hex.jl:122-164) splits
// into 3 n^2 independent LINE problems per element:  n volume nodes + 2 face nodes -> n(n-1)/2 + 2n unordered
// pairs.  ONE THREAD OWNS ONE LINE: it loads the n+2 nodes once into registers, evaluates every unordered pair
// exactly once (the symmetric loop of CompressibleNS/dg2D_CNS_cavity_optimized.jl:326-347), keeps all 5n + 10
// accumulators in registers, evaluates the surface flux at its own two face points, lifts it onto its own nodes
// (Ph / Lf as 1-D sweeps) and writes each result once.
//
// Work items per element and the phases built from them (kernels_hex_line.cuh: all threads run the phases one
// after another; kernels_hex_ws.cuh: line warps and point warps run them concurrently on different batches):
//   A  per volume node : (rho, u/2, beta) from Q, (log rho, log beta) from the stash          -> shared WV[7]
//   C  per line        : pair fluxes, surface flux F_EC(M,P).nJ - LFc (U+ - U), lift            -> shared A3[3][5]
//   E  per volume node : rhs = -(sum_d A3_d)/J ; rhstest ; res = a res + dt rhs ; Q += b res ;
//                        v_ufun of the NEW state (shared V[5]) + log stash                      (:212-219)
//   F  per face point  : entropy projection of the new state, u_vfun -> 8-plane traces          (:175-178)
//
// Operators never appear as matrices: the kernels hold the 1-D pair coefficients
// T[a][a'] = (w_a D[a,a'] - w_a' D[a',a]) / w_a, the end-point interpolation vectors lm, lp and the Gauss
// weights.  esdg_finalize proves (to 1e-10) that the caller's Qrh/Qsh/Qth, Ef, Ph, Lf ARE the Kronecker products
// of these before selecting this path; otherwise the generic table-driven kernels run.
//
// Trace format (8 planes per face point): rho, u/2, v/2, w/2, beta, log rho, log beta, lambda, where lambda is the
// reference's wave speed `wavespeed(rho, rhoU.n, E)` (euler_variables.jl:7-10) evaluated by the OWNER of the face
// point with its own normal -- exactly what the script's `lam[mapP]` gathers (dg3D_euler_hex.jl:186-193).
#pragma once
#include "fastmath.cuh"
#include "vecmath.cuh"

namespace esdg {

constexpr int HEX_TP = 8;     // trace planes (== ESDG_HEX_TRACE_PLANES in hex_fast.h)
constexpr int HEX_GEO = 34;   // global geometry record: G[9], J, 6 x (nxJ, nyJ, nzJ, sJ)
constexpr int HEX_SGEO = 40;  // shared-memory copy: + 1/sJ per face

struct HexParams {
    long long K, e_begin, e_count;
    // strided element ranges (pipelined host step: chunks along y of a structured slab): the e_count elements are
    // e_count / seg_len segments of seg_len consecutive elements, segment s starting at e_begin + s * seg_stride.
    // seg_len == 0: one contiguous range.  Batches never straddle a segment.
    long long seg_len, seg_stride;
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
    // curved elements (desc.affine = 0; dg3D_euler_hex.jl:69-101): rxJ.. at every hybridized point [e][9][Nh], J at every
    // volume node [e][Nq], (nxJ, nyJ, nzJ, sJ) at every face point [e][4][Nfq]; null on affine meshes
    const double *geoG, *geoJ, *geoN;
    const double* geoL;      // the same factors LINE-major: [e][(n+2)*3][NL] (node a of the line, rotated component c) -> coalesced over the lines
    double gamma, lf_scale, lg1, l2g1;   // lg1 = log(gamma-1), l2g1 = log(2(gamma-1))
    double rk_a, rk_b, dt;
    int mode;                // 0: write rhs ; 1: fused LSRK stage + traces of the new state
    int want_rhstest;
    double* partial;         // one slot per element: the batch starting at element e writes partial[e]
    int* bad;
    double T[64], lm[8], lp[8], w[8];
    double lmw[8], lpw[8];   // lm[a]/w[a], lp[a]/w[a]
};

struct Node {
    double rho, u[3], beta, lr, lb;
};

// ---- asynchronous copies ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* q) { asm volatile("prefetch.global.L2 [%0];" ::"l"(q)); }

// TMA bulk copy (cp.async.bulk, completion on an mbarrier): one instruction moves a whole plane
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
// the mbarrier also waits for every cp.async this thread has issued so far (counts as this thread's arrival)
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// batches of EB elements over the (possibly strided) element range of a launch
__device__ __forceinline__ long long range_batches(const HexParams& p, int EB) {
    if (p.seg_len <= 0) return (p.e_count + EB - 1) / EB;
    return (p.e_count / p.seg_len) * ((p.seg_len + EB - 1) / EB);
}
__device__ __forceinline__ void range_batch(const HexParams& p, int EB, long long bi, long long& e0, int& nb) {
    if (p.seg_len <= 0) {
        e0 = p.e_begin + bi * EB;
        nb = (int)max(0LL, min((long long)EB, p.e_begin + p.e_count - e0));
        return;
    }
    const long long bps = (p.seg_len + EB - 1) / EB, seg = bi / bps, w = bi - seg * bps;
    e0 = p.e_begin + seg * p.seg_stride + w * EB;
    nb = seg * p.seg_len < p.e_count ? (int)max(0LL, min((long long)EB, p.seg_len - w * EB)) : 0;
}

template <int n>
struct HexGeom {
    static constexpr int Nq = n * n * n, nfp = n * n, Nfq = 6 * nfp, Nh = Nq + Nfq, NL = 3 * nfp;
};
// plane stride of the shared volume-node array WV[7][.] of a batch of EB elements: even, with room for the one-double
// shift that aligns a batch starting at an odd global node index with 16-byte bulk copies (kernels_hex_ws.cuh)
template <int n, int EB>
constexpr int wv_stride() { return (EB * HexGeom<n>::Nq + 3) / 2 * 2; }

// face point (face = 2*d + side, m) -> first volume node of its line and the line stride
template <int n>
__device__ __forceinline__ void face_line(int d, int m, int& base, int& stride) {
    const int ip = m % n, jp = m / n;
    if (d == 0) { base = n * jp + n * n * ip; stride = 1; }
    else if (d == 1) { base = jp + n * n * ip; stride = n; }
    else { base = jp + n * ip; stride = n * n; }
}

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

// node `idx` of a 7-plane array with plane stride S, velocity components in the rotated order (pl0, pl1, pl2)
__device__ __forceinline__ Node ld_node_rot(const double* W, int S, int idx, int pl0, int pl1, int pl2) {
    Node a;
    a.rho = W[idx];
    a.u[0] = W[pl0 * S + idx];
    a.u[1] = W[pl1 * S + idx];
    a.u[2] = W[pl2 * S + idx];
    a.beta = W[4 * S + idx];
    a.lr = W[5 * S + idx];
    a.lb = W[6 * S + idx];
    return a;
}

// the same, re-read at every use (volatile: the compiler must not keep the node in registers between pairs)
__device__ __forceinline__ Node ld_node_rot_again(const double* W, int S, int idx, int pl0, int pl1, int pl2) {
    const volatile double* V = W;
    Node a;
    a.rho = V[idx];
    a.u[0] = V[pl0 * S + idx];
    a.u[1] = V[pl1 * S + idx];
    a.u[2] = V[pl2 * S + idx];
    a.beta = V[4 * S + idx];
    a.lr = V[5 * S + idx];
    a.lb = V[6 * S + idx];
    return a;
}

// Entropy-conservative two-point flux (euler_fluxes.jl:51-89 with logmean.jl:14-28) on W independent pairs,
// contracted with g; one reciprocal per pair serves rholog, 1/betalog and pa = rhoavg/(betaL+betaR).
// Node.u holds HALF velocities.  All within an ulp of the reference formulas:
//  * sums instead of means: with s = aL + aR the Taylor argument is f/2 = da/s, f^2 = 4 (da/s)^2, and the
//    branch test |da| < 1e-4 (s/2) is the same predicate written as |da| < 5e-5 s (exact scaling by 2);
//  * in the Taylor branch |f| < 1e-4, so the f^4 terms of logmean.jl:18-27 are < 5.2e-18, a twentieth of an ulp of
//    the leading 1: 1 - .2 f^2 is the reference polynomial to double precision;
//  * DIAG (axis-aligned elements): every flux component is proportional to the single metric term g[0]; it is
//    factored out here (the result is F / g[0]) and applied once to accumulated sums by the caller.
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
    uavg[0] = A.u[0] + B.u[0];
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

// log / exp of fastmath.cuh on W lanes (fdlibm reductions, MUFU-seeded reciprocal)
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

// volume-node primitives (rho, u/2, beta) and rhoe from the conservative state (betafun, euler_variables.jl:30-39)
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

// v_ufun (euler_variables.jl:79-86) with s = log((g-1)rhoe/rho^g) = (1-g) log rho - log beta - log 2
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

// u_vfun (euler_variables.jl:95-113) followed by betafun / velocity / logs / wave speed, all from the projected
// entropy variables: rhoeV = exp((log(g-1) - g log(-vE) - s)/(g-1)), rho = rhoeV(-vE), u = vU/(-vE),
// beta = -vE/(2(g-1)) (identically rho/(2p)), log beta = log(-vE) - log(2(g-1)).  log rho is taken of the
// ROUNDED rho (one more log): the pair flux divides by differences of these logs, and for rho ~ 1 only
// log(fl(rho)) is consistent with fl(rho) to better than an ulp of 1 (the reference computes it that way).
// lambda = sqrt(|u.nhat|) + sqrt(g p_n / rho) with p_n = (g-1)(E - .5 rho (u.nhat)^2), euler_variables.jl:7-10
// (sic: square root of |u_n|), i.e. g p_n / rho = g(g-1) [ 1/(-vE) + .5 (|u|^2 - (u.nhat)^2) ].
// nh = nJ * (2 / sJ) (the 2 undoes the half velocities).
template <int W>
__device__ __forceinline__ void trace_from_entropy_v(const VD<W>* V, const VD<W>* nh, double gamma, double lg1, double l2g1, VD<W>* Tv) {
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
    Tv[5] = vlog_pos(Tv[0]);
    Tv[6] = lv - l2g1;
    const VD<W> un = vfma(Tv[3], nh[2], vfma(Tv[2], nh[1], Tv[1] * nh[0]));              // u . nhat
    const VD<W> hq = vfma(Tv[1], Tv[1], vfma(Tv[2], Tv[2], Tv[3] * Tv[3]));              // |u|^2 / 4
    const VD<W> c2 = (gamma * (gamma - 1.0)) * (inv + vfma(un * un, -.5, 2.0 * hq));
    Tv[7] = vsqrt_fast(vabs(un)) + vsqrt_fast(c2);
}

// ---- the line problem ----------------------------------------------------------------------------------
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

#ifndef ESDG_LINE_WC
#define ESDG_LINE_WC 2   // pairs evaluated together in the line problem (one-shot kernel: 168 registers)
#endif
#ifndef ESDG_LINE_WC_WS
#define ESDG_LINE_WC_WS 3   // ... in the persistent kernel, whose line warps own 184 registers (measured: 2 -> 3 = -6 %, 4 = -3 %, 5 = +29 %)
#endif

// Line l of a batch of EB elements (EN = EB*Nq, EF = EB*Nfq are the plane strides of the shared arrays):
// all pairs, the surface flux at the line's two face points (dg3D_euler_hex.jl:182-197), the face rows
// H = flux* -/+ sum_a l[a] F(a,face) and their lift onto the line's own nodes.  Result: acc[a][f] in the ROTATED
// field order (rho, momentum d, c1, c2, E); line_store writes it back in physical order.
// OWN_GATHER: the exterior traces of the line's two face points were fetched by THIS thread with cp.async
// (line_gather below); they are awaited right before their first use.
template <int n, int EB, bool DIAG, bool OWN_GATHER = false, int WC = ESDG_LINE_WC>
__device__ __forceinline__ void line_compute(int l, const double* sWV, const double* sWF, const double* sP, const double* sGeo,
                                             const HexParams& p, double c_igm, double (&acc)[n][5]) {
    using G = HexGeom<n>;
    constexpr int Nq = G::Nq, nfp = G::nfp, Nfq = G::Nfq, NL = G::NL, EN = EB * Nq, EF = EB * Nfq;
    const int eb = l / NL, ll = l - eb * NL;
    const int d = ll / nfp, m = ll - d * nfp;
    int base, stride;
    face_line<n>(d, m, base, stride);
    base += eb * Nq;
    const int c1 = d == 2 ? 0 : d + 1, c2 = c1 == 2 ? 0 : c1 + 1;   // physical components in rotated order d,c1,c2
    const int pl0 = 1 + d, pl1 = 1 + c1, pl2 = 1 + c2;
    const double* Gm = sGeo + eb * HEX_SGEO;
    const double g0 = Gm[3 * d + d], g1 = Gm[3 * c1 + d], g2 = Gm[3 * c2 + d];   // (r_x, r_y, r_z) J of direction d, rotated
    Node nd[n];
#pragma unroll
    for (int a = 0; a < n; ++a) nd[a] = ld_node_rot(sWV, wv_stride<n, EB>(), base + a * stride, pl0, pl1, pl2);
#pragma unroll
    for (int a = 0; a < n; ++a)
#pragma unroll
        for (int f = 0; f < 5; ++f) acc[a][f] = 0.0;
    // volume-volume pairs: acc_a += T[a][b] F(a,b), acc_b += T[b][a] F(a,b)
    line_vv_all<n, DIAG, 0, WC>(nd, acc, g0, g1, g2, c_igm, p);
    // volume-face pairs: node a against the two face points its line ends in
    Node fn[2];
    int fidx[2];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        fidx[side] = eb * Nfq + (2 * d + side) * nfp + m;
        fn[side] = ld_node_rot(sWF, EF, fidx[side], pl0, pl1, pl2);
    }
    double h[2][5];
#pragma unroll
    for (int f = 0; f < 5; ++f) h[0][f] = h[1][f] = 0.0;
    line_vf_all<n, DIAG, 0, WC>(nd, fn, acc, h, g0, g1, g2, c_igm, p);
    if (DIAG) {   // pair_flux_v<.., true> returns F / g0: the single metric term, once per accumulated sum
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            h[0][f] *= g0;
            h[1][f] *= g0;
#pragma unroll
            for (int a = 0; a < n; ++a) acc[a][f] *= g0;
        }
    }
    // surface flux at the two face points (lane = side): F_EC(M,P).nJ - LFc (U+ - U)
    VD<2> FS[5];
    if (OWN_GATHER) cp_async_wait_all();
    {
        NodeV<2> M, P;
        VD<2> nJ[3], sJ, lamM, lamP;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            set_lane(M, side, fn[side]);
            set_lane(P, side, ld_node_rot(sP, EF, fidx[side], pl0, pl1, pl2));
            const double* gf = Gm + 10 + (2 * d + side) * 4;
            nJ[0][side] = gf[d]; nJ[1][side] = gf[c1]; nJ[2][side] = gf[c2]; sJ[side] = gf[3];
            lamM[side] = sWF[7 * EF + fidx[side]];
            lamP[side] = sP[7 * EF + fidx[side]];
        }
        pair_flux_v<2, DIAG>(M, P, nJ, c_igm, FS);
        if (DIAG) {
#pragma unroll
            for (int f = 0; f < 5; ++f) FS[f] = FS[f] * nJ[0];
        }
        if (p.lf_scale != 0.0) {
            // conservative states from the primitive traces: E = rho (1/(2(g-1) beta) + .5 |u|^2), u = 2 (u/2)
            const VD<2> LFc = p.lf_scale * vmax(lamM, lamP) * sJ;
            const VD<2> ibM = vrcp(M.beta), ibP = vrcp(P.beta);
            const VD<2> qM = vfma(M.u[0], M.u[0], vfma(M.u[1], M.u[1], M.u[2] * M.u[2]));   // |u|^2 / 4
            const VD<2> qP = vfma(P.u[0], P.u[0], vfma(P.u[1], P.u[1], P.u[2] * P.u[2]));
            const VD<2> tM = 2.0 * M.rho, tP = 2.0 * P.rho;
            VD<2> jump[5];
            jump[0] = P.rho - M.rho;
            jump[1] = vfma(tP, P.u[0], -(tM * M.u[0]));
            jump[2] = vfma(tP, P.u[1], -(tM * M.u[1]));
            jump[3] = vfma(tP, P.u[2], -(tM * M.u[2]));
            jump[4] = vfma(P.rho, vfma(ibP, c_igm, 2.0 * qP), -(M.rho * vfma(ibM, c_igm, 2.0 * qM)));
#pragma unroll
            for (int f = 0; f < 5; ++f) FS[f] = vfma(-LFc, jump[f], FS[f]);
        }
    }
    // face rows and their lift: only the nodes of THIS line lift from these two face points (Lf as a 1-D sweep)
#pragma unroll
    for (int f = 0; f < 5; ++f) {
        const double Hm = FS[f][0] + h[0][f], Hp = FS[f][1] - h[1][f];
#pragma unroll
        for (int a = 0; a < n; ++a) acc[a][f] = fma(p.lpw[a], Hp, fma(p.lmw[a], Hm, acc[a][f]));
    }
}

// this thread's (n + 2) x 3 geometric factors: geoL (line-major, coalesced over the lines) -> its column of sGL, asynchronously
template <int n, int EB>
__device__ __forceinline__ void line_geometry_load(int l, long long e0, const HexParams& p, double* sGL, int GS) {
    constexpr int NL = HexGeom<n>::NL, NG = 3 * (n + 2);
    const int eb = l / NL, ll = l - eb * NL;
    const double* GL = p.geoL + (e0 + eb) * NG * NL + ll;
#pragma unroll
    for (int k = 0; k < NG; ++k) cp_async8(sGL + k * GS, GL + k * NL);
}

// The line problem on a CURVED element (dg3D_euler_hex.jl:145-151): every pair (i, j) of hybridized points is contracted with
// the AVERAGE of the two points' geometric factors, .5 (rxJ[i] + rxJ[j]) .., the surface flux takes the face point's own
// normal.  The (n + 2) x 3 factors of a line's direction sit in the thread's own column of shared memory (registers are
// full: at N = 4 keeping them there cost 190 bytes of spills inside the pair loop).
template <int n, int EB, int WC = 2>
__device__ __forceinline__ void line_compute_curved(int l, long long e0, const double* sWV, const double* sWF, const double* sP,
                                                    const double* sGL, int GS, const HexParams& p, double c_igm, double (&acc)[n][5]) {
    using G = HexGeom<n>;
    constexpr int Nq = G::Nq, nfp = G::nfp, Nfq = G::Nfq, NL = G::NL, EF = EB * Nfq;
    constexpr PairTab<n> PT{};
    const int eb = l / NL, ll = l - eb * NL;
    const int d = ll / nfp, m = ll - d * nfp;
    int base, stride;
    face_line<n>(d, m, base, stride);
    const int c1 = d == 2 ? 0 : d + 1, c2 = c1 == 2 ? 0 : c1 + 1;
    const int pl0 = 1 + d, pl1 = 1 + c1, pl2 = 1 + c2;
    // (d_x, d_y, d_z) J of direction d at the line's n volume nodes and 2 face points (rotated components) were copied
    // by THIS thread from the line-major array geoL into its own column of sGL (stride GS): gn(a, c) = sGL[(3 a + c) GS]
    auto gn = [&](int a, int c) { return sGL[(a * 3 + c) * GS]; };
    int fidx[2];
#pragma unroll
    for (int side = 0; side < 2; ++side) fidx[side] = eb * Nfq + (2 * d + side) * nfp + m;
    // registers are full (accumulators, two face nodes, the pair in flight): the line's volume nodes are re-read from shared
    // memory for every pair instead of being held (7 loads per node and pair against ~100 FP64 issue slots per pair)
    auto nd = [&](int a) { return ld_node_rot_again(sWV, wv_stride<n, EB>(), eb * Nq + base + a * stride, pl0, pl1, pl2); };
#pragma unroll
    for (int a = 0; a < n; ++a)
#pragma unroll
        for (int f = 0; f < 5; ++f) acc[a][f] = 0.0;
    // volume-volume pairs
#pragma unroll
    for (int k0 = 0; k0 < PT.NP; k0 += WC) {
        NodeV<WC> A, B;
        VD<WC> g[3], F[5];
#pragma unroll
        for (int q = 0; q < WC; ++q) {
            const int k = k0 + q < PT.NP ? k0 + q : k0;
            set_lane(A, q, nd(PT.a[k]));
            set_lane(B, q, nd(PT.b[k]));
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c][q] = .5 * (gn(PT.a[k], c) + gn(PT.b[k], c));
        }
        pair_flux_v<WC, false>(A, B, g, c_igm, F);
#pragma unroll
        for (int q = 0; q < WC; ++q) {
            if (k0 + q < PT.NP) {
                const int a = PT.a[k0 + q], b = PT.b[k0 + q];
#pragma unroll
                for (int f = 0; f < 5; ++f) {
                    acc[a][f] = fma(p.T[a * n + b], F[f].v[q], acc[a][f]);
                    acc[b][f] = fma(p.T[b * n + a], F[f].v[q], acc[b][f]);
                }
            }
        }
    }
    // volume-face pairs (lane = side)
    Node fn[2];
#pragma unroll
    for (int side = 0; side < 2; ++side) fn[side] = ld_node_rot(sWF, EF, fidx[side], pl0, pl1, pl2);
    double h[2][5];
#pragma unroll
    for (int f = 0; f < 5; ++f) h[0][f] = h[1][f] = 0.0;
#pragma unroll
    for (int a = 0; a < n; ++a) {
        NodeV<2> A, B;
        VD<2> g[3], F[5];
#pragma unroll
        const Node na = nd(a);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            set_lane(A, side, na);
            set_lane(B, side, fn[side]);
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c][side] = .5 * (gn(a, c) + gn(n + side, c));
        }
        pair_flux_v<2, false>(A, B, g, c_igm, F);
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            acc[a][f] = fma(-p.lmw[a], F[f][0], acc[a][f]);
            acc[a][f] = fma(p.lpw[a], F[f][1], acc[a][f]);
            h[0][f] = fma(p.lm[a], F[f][0], h[0][f]);
            h[1][f] = fma(p.lp[a], F[f][1], h[1][f]);
        }
    }
    // surface flux at the two face points with their own normals
    VD<2> FS[5];
    {
        NodeV<2> M, P;
        VD<2> nJ[3], sJ, lamM, lamP;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            set_lane(M, side, fn[side]);
            set_lane(P, side, ld_node_rot(sP, EF, fidx[side], pl0, pl1, pl2));
            const double* gf = p.geoN + (e0 + eb) * 4 * Nfq + (2 * d + side) * nfp + m;
            nJ[0][side] = gf[d * Nfq]; nJ[1][side] = gf[c1 * Nfq]; nJ[2][side] = gf[c2 * Nfq]; sJ[side] = gf[3 * Nfq];
            lamM[side] = sWF[7 * EF + fidx[side]];
            lamP[side] = sP[7 * EF + fidx[side]];
        }
        pair_flux_v<2, false>(M, P, nJ, c_igm, FS);
        if (p.lf_scale != 0.0) {
            const VD<2> LFc = p.lf_scale * vmax(lamM, lamP) * sJ;
            const VD<2> ibM = vrcp(M.beta), ibP = vrcp(P.beta);
            const VD<2> qM = vfma(M.u[0], M.u[0], vfma(M.u[1], M.u[1], M.u[2] * M.u[2]));
            const VD<2> qP = vfma(P.u[0], P.u[0], vfma(P.u[1], P.u[1], P.u[2] * P.u[2]));
            const VD<2> tM = 2.0 * M.rho, tP = 2.0 * P.rho;
            VD<2> jump[5];
            jump[0] = P.rho - M.rho;
            jump[1] = vfma(tP, P.u[0], -(tM * M.u[0]));
            jump[2] = vfma(tP, P.u[1], -(tM * M.u[1]));
            jump[3] = vfma(tP, P.u[2], -(tM * M.u[2]));
            jump[4] = vfma(P.rho, vfma(ibP, c_igm, 2.0 * qP), -(M.rho * vfma(ibM, c_igm, 2.0 * qM)));
#pragma unroll
            for (int f = 0; f < 5; ++f) FS[f] = vfma(-LFc, jump[f], FS[f]);
        }
    }
#pragma unroll
    for (int f = 0; f < 5; ++f) {
        const double Hm = FS[f][0] + h[0][f], Hp = FS[f][1] - h[1][f];
#pragma unroll
        for (int a = 0; a < n; ++a) acc[a][f] = fma(p.lpw[a], Hp, fma(p.lmw[a], Hm, acc[a][f]));
    }
}

// exterior-trace indices of the two face points of line l in the batch starting at element e0 ...
template <int n, int EB>
__device__ __forceinline__ void line_mapP(int l, long long e0, const HexParams& p, int (&ip)[2]) {
    using G = HexGeom<n>;
    const int eb = l / G::NL, ll = l - eb * G::NL;
    const int d = ll / G::nfp, m = ll - d * G::nfp;
    const int* mp = p.mapP + (e0 + eb) * G::Nfq + 2 * d * G::nfp + m;
    ip[0] = mp[0];
    ip[1] = mp[G::nfp];
}
// ... and the asynchronous copy of their 8 trace planes into this thread's own slots of sP
template <int n, int EB>
__device__ __forceinline__ void line_gather(int l, const int (&ip)[2], const HexParams& p, double* sP) {
    using G = HexGeom<n>;
    constexpr int EF = EB * G::Nfq;
    const int eb = l / G::NL, ll = l - eb * G::NL;
    const int d = ll / G::nfp, m = ll - d * G::nfp;
    double* dst = sP + eb * G::Nfq + 2 * d * G::nfp + m;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
#pragma unroll
        for (int q = 0; q < HEX_TP; ++q) cp_async8(dst + q * EF + side * G::nfp, p.Tcur + q * p.t_stride + ip[side]);
    }
}

template <int n, int EB>
__device__ __forceinline__ void line_store(int l, const double (&acc)[n][5], double* sA3) {
    using G = HexGeom<n>;
    constexpr int Nq = G::Nq, nfp = G::nfp, NL = G::NL, EN = EB * Nq;
    const int eb = l / NL, ll = l - eb * NL;
    const int d = ll / nfp, m = ll - d * nfp;
    int base, stride;
    face_line<n>(d, m, base, stride);
    base += eb * Nq;
    const int c1 = d == 2 ? 0 : d + 1, c2 = c1 == 2 ? 0 : c1 + 1;
    double* A = sA3 + d * 5 * EN;
#pragma unroll
    for (int a = 0; a < n; ++a) {
        const int idx = base + a * stride;
        A[idx] = acc[a][0];
        A[(1 + d) * EN + idx] = acc[a][1];
        A[(1 + c1) * EN + idx] = acc[a][2];
        A[(1 + c2) * EN + idx] = acc[a][3];
        A[4 * EN + idx] = acc[a][4];
    }
}

// ---- per-node / per-face-point items ----------------------------------------------------------------------
// geometry records of the batch -> shared (34 doubles per element + 1/sJ per face)
template <int EB>
__device__ __forceinline__ void stage_geo(int t, int NT, int nb, long long e0, const HexParams& p, double* sGeo) {
    for (int it = t; it < nb * HEX_SGEO; it += NT) {
        const int eb = it / HEX_SGEO, k = it - eb * HEX_SGEO;
        const double* g = p.geo + (e0 + eb) * p.geo_stride;
        sGeo[it] = k < HEX_GEO ? g[k] : fast_rcp(g[10 + (k - HEX_GEO) * 4 + 3]);
    }
}

// the same in two steps for the pipelined kernel: raw records by cp.async, reciprocals once they have landed
template <int EB>
__device__ __forceinline__ void stage_geo_async(int t, int NT, int nb, long long e0, const HexParams& p, double* sGeo) {
    for (int it = t; it < nb * HEX_GEO; it += NT) {
        const int eb = it / HEX_GEO, k = it - eb * HEX_GEO;
        cp_async8(sGeo + eb * HEX_SGEO + k, p.geo + (e0 + eb) * p.geo_stride + k);
    }
}
template <int EB>
__device__ __forceinline__ void stage_geo_finish(int t, int NT, int nb, double* sGeo) {   // after this thread's copies landed
    for (int it = t; it < nb * HEX_GEO; it += NT) {
        const int eb = it / HEX_GEO, k = it - eb * HEX_GEO;
        if (k >= 10 && ((k - 10) & 3) == 3) sGeo[eb * HEX_SGEO + HEX_GEO + (k - 10) / 4] = fast_rcp(sGeo[eb * HEX_SGEO + k]);
    }
}

// A: volume node `it` of the batch: primitives from Q, logs from the stash -> sWV[7][EN]
template <int n, int EB>
__device__ __forceinline__ void node_prims(int it, long long g, const HexParams& p, double* sWV, bool& bad) {
    constexpr int EN = wv_stride<n, EB>();
    VD<1> U[5], rho, hu[3], beta, rhoe;
#pragma unroll
    for (int f = 0; f < 5; ++f) U[f][0] = p.Q[f * p.q_stride + g];
    const double lr = p.Lcur[g], lb = p.Lcur[p.q_stride + g];
    vol_prim_v<1>(U, p.gamma, rho, hu, beta, rhoe);
    bad = bad || !(U[0][0] > 0.0) || !(rhoe[0] > 0.0);
    sWV[it] = rho[0];
    sWV[EN + it] = hu[0][0];
    sWV[2 * EN + it] = hu[1][0];
    sWV[3 * EN + it] = hu[2][0];
    sWV[4 * EN + it] = beta[0];
    sWV[5 * EN + it] = lr;
    sWV[6 * EN + it] = lb;
}

// A, in place: the five WV planes hold Q (copied there asynchronously), planes 5, 6 already hold the stashed logs
template <int n, int EB>
__device__ __forceinline__ void node_prims_inplace(int it, const HexParams& p, double* sWV, bool& bad) {
    constexpr int EN = wv_stride<n, EB>();
    VD<1> U[5], rho, hu[3], beta, rhoe;
#pragma unroll
    for (int f = 0; f < 5; ++f) U[f][0] = sWV[f * EN + it];
    vol_prim_v<1>(U, p.gamma, rho, hu, beta, rhoe);
    bad = bad || !(U[0][0] > 0.0) || !(rhoe[0] > 0.0);
    sWV[EN + it] = hu[0][0];
    sWV[2 * EN + it] = hu[1][0];
    sWV[3 * EN + it] = hu[2][0];
    sWV[4 * EN + it] = beta[0];
}

// E: volume node `it`: rhs, rhstest, LSRK stage, entropy variables of the new state -> sV[5][EN], log stash
struct NodeIn {   // what the update of one node reads from global memory (loaded ahead of its use)
    double U[5], R[5];
};
__device__ __forceinline__ void node_load(long long g, const HexParams& p, NodeIn& in) {
#pragma unroll
    for (int f = 0; f < 5; ++f) in.U[f] = p.Q[f * p.q_stride + g];
    if (p.mode == 1) {
#pragma unroll
        for (int f = 0; f < 5; ++f) in.R[f] = p.res[f * p.q_stride + g];
    }
}

template <int n, int EB, bool CURVED = false>
__device__ __forceinline__ void node_update(int it, long long g, const HexParams& p, const NodeIn& in, const double* sA3, double* sV,
                                            const double* sGeo, const double* s_w, double& tsum, bool& bad) {
    using G = HexGeom<n>;
    constexpr int Nq = G::Nq, EN = EB * Nq;
    VD<1> U[5], rr[5];
    double R[5];
#pragma unroll
    for (int f = 0; f < 5; ++f) { U[f][0] = in.U[f]; R[f] = in.R[f]; }
    const int eb = it / Nq, i = it - eb * Nq;
    const double J = CURVED ? p.geoJ[g] : sGeo[eb * HEX_SGEO + 9];   // curved: `rhsQ ./ J` node by node (dg3D_euler_hex.jl:212)
    VD<1> Jv;
    Jv[0] = J;
    const VD<1> mJ = -vrcp(Jv);
#pragma unroll
    for (int f = 0; f < 5; ++f) rr[f][0] = ((sA3[f * EN + it] + sA3[(5 + f) * EN + it]) + sA3[(10 + f) * EN + it]) * mJ[0];
    if (p.want_rhstest) {
        const int xa = i % n, xb = (i / n) % n, xc = i / (n * n);
        const double wJ = s_w[xa] * s_w[xb] * s_w[xc] * J;
        VD<1> V[5], lr, lb;
        lr[0] = p.Lcur[g];
        lb[0] = p.Lcur[p.q_stride + g];
        const VD<1> hir = .5 * vrcp(U[0]);
        const VD<1> ke = hir * vfma(U[1], U[1], vfma(U[2], U[2], U[3] * U[3]));
        entropy_vars_v<1>(U, U[4] - ke, lr, lb, p.gamma, V);
#pragma unroll
        for (int f = 0; f < 5; ++f) tsum = fma(wJ * V[f][0], rr[f][0], tsum);
    }
    if (p.mode == 0) {
#pragma unroll
        for (int f = 0; f < 5; ++f) p.rhs[f * p.q_stride + g] = rr[f][0];
    } else {
        VD<1> Un[5];
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            const double rs = fma(p.rk_a, R[f], p.dt * rr[f][0]);
            Un[f][0] = fma(p.rk_b, rs, U[f][0]);
            p.res[f * p.q_stride + g] = rs;
            p.Qout[f * p.q_stride + g] = Un[f][0];
        }
        VD<1> rho, hu[3], beta, rhoen, V[5];
        vol_prim_v<1>(Un, p.gamma, rho, hu, beta, rhoen);
        const VD<1> lr = vlog_pos<1>(rho), lb = vlog_pos<1>(beta);
        entropy_vars_v<1>(Un, rhoen, lr, lb, p.gamma, V);
        bad = bad || !(Un[0][0] > 0.0) || !(rhoen[0] > 0.0);
        p.Lnext[g] = lr[0];
        p.Lnext[p.q_stride + g] = lb[0];
#pragma unroll
        for (int f = 0; f < 5; ++f) sV[f * EN + it] = V[f][0];
    }
}

// entropy variables + log stash of a given state (no update): used by the traces-only kernel
template <int n, int EB>
__device__ __forceinline__ void node_entropy(int it, long long g, const HexParams& p, double* sV, bool& bad) {
    constexpr int EN = EB * HexGeom<n>::Nq;
    VD<1> U[5], rho, hu[3], beta, rhoe, V[5];
#pragma unroll
    for (int f = 0; f < 5; ++f) U[f][0] = p.Q[f * p.q_stride + g];
    vol_prim_v<1>(U, p.gamma, rho, hu, beta, rhoe);
    bad = bad || !(U[0][0] > 0.0) || !(rhoe[0] > 0.0);
    const VD<1> ok_rho = vset<1>(U[0][0] > 0.0 ? rho[0] : 1.0), ok_beta = vset<1>(rhoe[0] > 0.0 && U[0][0] > 0.0 ? beta[0] : 1.0);
    const VD<1> lr = vlog_pos<1>(ok_rho), lb = vlog_pos<1>(ok_beta);
    entropy_vars_v<1>(U, rhoe, lr, lb, p.gamma, V);
    p.Lnext[g] = lr[0];
    p.Lnext[p.q_stride + g] = lb[0];
#pragma unroll
    for (int f = 0; f < 5; ++f) sV[f * EN + it] = V[f][0];
}

// F: line l of the batch: project the entropy variables of the new state to the line's TWO face points (lane = side;
// the n x 5 shared-memory reads serve both), 8-plane traces -> Tnext
#ifndef ESDG_F_SEQ
#define ESDG_F_SEQ 0   // 1: the two ends of a line one after the other (W = 1, fewer registers) instead of interleaved (W = 2)
#endif
template <int n, int EB, bool CURVED = false>
__device__ __forceinline__ void line_traces(int l, long long e0, const HexParams& p, const double* sV, const double* sGeo,
                                            const double* s_lm, const double* s_lp) {
    using G = HexGeom<n>;
    constexpr int Nq = G::Nq, nfp = G::nfp, Nfq = G::Nfq, NL = G::NL, EN = EB * Nq;
    const int eb = l / NL, ll = l - eb * NL;
    const int d = ll / nfp, m = ll - d * nfp;
    int base, stride;
    face_line<n>(d, m, base, stride);
    const double* v = sV + eb * Nq + base;
    VD<2> V[5], nh[3], Tv[HEX_TP];
#pragma unroll
    for (int f = 0; f < 5; ++f) V[f][0] = V[f][1] = 0.0;
#pragma unroll
    for (int a = 0; a < n; ++a) {
        const double la = s_lm[a], lb = s_lp[a];
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            const double x = v[f * EN + a * stride];
            V[f][0] = fma(la, x, V[f][0]);
            V[f][1] = fma(lb, x, V[f][1]);
        }
    }
    const double* ge = sGeo + eb * HEX_SGEO;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        if (CURVED) {   // the face point's own normal (dg3D_euler_hex.jl:80-84)
            const double* gn = p.geoN + (e0 + eb) * 4 * Nfq + (2 * d + side) * nfp + m;
            const double s2 = 2.0 * fast_rcp(gn[3 * Nfq]);
            nh[0][side] = gn[0] * s2; nh[1][side] = gn[Nfq] * s2; nh[2][side] = gn[2 * Nfq] * s2;
        } else {
            const double* gf = ge + 10 + (2 * d + side) * 4;
            const double s2 = 2.0 * ge[HEX_GEO + 2 * d + side];
            nh[0][side] = gf[0] * s2; nh[1][side] = gf[1] * s2; nh[2][side] = gf[2] * s2;
        }
    }
    double* T0 = p.Tnext + (e0 + eb) * Nfq + 2 * d * nfp + m;
#if ESDG_F_SEQ
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        VD<1> V1[5], nh1[3], Tv1[HEX_TP];
#pragma unroll
        for (int f = 0; f < 5; ++f) V1[f][0] = side ? V[f][1] : V[f][0];
#pragma unroll
        for (int c = 0; c < 3; ++c) nh1[c][0] = side ? nh[c][1] : nh[c][0];
        trace_from_entropy_v<1>(V1, nh1, p.gamma, p.lg1, p.l2g1, Tv1);
#pragma unroll
        for (int q = 0; q < HEX_TP; ++q) T0[q * p.t_stride + side * nfp] = Tv1[q][0];
    }
#else
    trace_from_entropy_v<2>(V, nh, p.gamma, p.lg1, p.l2g1, Tv);
#pragma unroll
    for (int q = 0; q < HEX_TP; ++q) {
        T0[q * p.t_stride] = Tv[q][0];
        T0[q * p.t_stride + nfp] = Tv[q][1];
    }
#endif
}

}  // namespace esdg
