// Pointwise Euler physics in FP64 for sm_100a (generic and CNS kernels).
//
// Device versions of the functions the reference's rhs calls per point:
//   logmean        examples/EntropyStableEuler/logmean.jl:14-28
//   euler_fluxes   examples/EntropyStableEuler/euler_fluxes.jl:23-48 (2D), :51-89 (3D)
//   v_ufun/u_vfun  examples/EntropyStableEuler/euler_variables.jl:79-86 / :107-113
//   betafun        :30-39 ; wavespeed :7-10
// Arithmetic is arranged for the FP64 pipe: no pow() (s = log((g-1) rhoe / rho^g) is formed from two
// logs; rhoe(V) = exp((log(g-1) - g log(-vE) - s)/(g-1)) from one log and one exp), MUFU-seeded
// reciprocals instead of divisions, one reciprocal per two-point flux, no divergent branches.
// Results agree with the reference formulas to a few ulp.
#pragma once
#include <cuda_runtime.h>

#include "fastmath.cuh"

namespace esdg {

template <int DIM>
struct Prim {  // (rho, u, v[, w], beta) + logs, the per-point operand of the two-point flux
    double rho, vel[DIM], beta, lrho, lbeta;
};

// Entropy-conservative two-point flux contracted with a direction vector c:
//   out[f] = sum_d c[d] * F_d[f](L,R)
// Contracting first is exact algebra on the reference's Fx,Fy,Fz:
//   F_d = (rholog*uavg_d, rholog*uavg_d*uavg_k + pa*delta_dk, Ep*uavg_d)
//   => sum_d c_d F_d = (m, m*uavg_k + pa*c_k, Ep*un)  with un = c.uavg, m = rholog*un.
// logmean(aL,aR,logL,logR) (logmean.jl:14-28): |f| < 1e-4 (f = da/aavg) -> aavg*(1 + v*(-.2 - v*(.0512 -
// v*0.026038857142857))), v = f^2, else -da/(logL-logR).  The series constants are the reference's (they
// are not the expansion of the log mean: the branches differ by 1e-9 at the switch, so the test is exact).
// ONE reciprocal R = 1/(X Y sb) yields 1/X, 1/Y and 1/sb, i.e. rholog = da/X, 1/betalog and
// pa = rhoavg/(betaL+betaR), where X = aavg (Taylor branch) or logR-logL, Y likewise for beta.
template <int DIM>
__device__ __forceinline__ void ec_flux_dot(const Prim<DIM>& L, const Prim<DIM>& R, const double* c,
                                            double igm1x2, double* out) {
    const double dar = R.rho - L.rho, hr = .5 * (L.rho + R.rho);
    const double dab = R.beta - L.beta, sb = L.beta + R.beta, hb = .5 * sb;
    const bool tr = fabs(dar) < 1e-4 * hr, tb = fabs(dab) < 1e-4 * hb;
    const double X = tr ? hr : (R.lrho - L.lrho);
    const double Y = tb ? hb : dab;
    const double XY = X * Y;
    const double Rc = fast_rcp(XY * sb);
    const double iX = Rc * (Y * sb), iY = Rc * (X * sb), iZ = Rc * XY;
    const double qr = dar * iX;                       // f (Taylor branch) or the log mean itself
    const double vr = qr * qr;
    const double rholog = tr ? hr * (1.0 + vr * (-.2 - vr * (.0512 - vr * 0.026038857142857))) : qr;
    const double fb = dab * iY, vb = fb * fb;
    const double xb = vb * (-.2 - vb * (.0512 - vb * 0.026038857142857));
    const double ibl = iY * (tb ? (1.0 - xb) : (R.lbeta - L.lbeta));   // 1/betalog ((1+x)^-1 = 1-x to 1e-17)
    const double pa = hr * iZ;                                          // rhoavg/(betaL+betaR)
    double uavg[DIM];
    double unorm = 0.0, un = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        uavg[d] = .5 * (L.vel[d] + R.vel[d]);
        unorm = fma(L.vel[d], R.vel[d], unorm);
        un = fma(c[d], uavg[d], un);
    }
    // Ep = rholog/(2(gamma-1)betalog) + pa + .5*rholog*unorm ; igm1x2 = 1/(2(gamma-1))
    const double Ep = fma(rholog, fma(.5, unorm, ibl * igm1x2), pa);
    const double m = rholog * un;
    out[0] = m;
#pragma unroll
    for (int d = 0; d < DIM; ++d) out[1 + d] = fma(m, uavg[d], pa * c[d]);
    out[1 + DIM] = Ep * un;
}

// conservative (rho, rhoU, E) -> entropy variables;  s = log((g-1) rhoe / rho^g)
template <int DIM>
__device__ __forceinline__ void v_of_u(const double* U, double gamma, double lg1, double* V, bool& bad) {   // lg1 = log(gamma-1)
    const double rho = U[0];
    double m2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) m2 = fma(U[1 + d], U[1 + d], m2);
    const double E = U[1 + DIM];
    const double rhoe = E - .5 * m2 * fast_rcp(rho);
    const bool ok = (rho > 0.0) && (rhoe > 0.0);
    bad = bad || !ok;
    const double s = lg1 + log_pos(ok ? rhoe : 1.0) - gamma * log_pos(ok ? rho : 1.0);
    const double ir = fast_rcp(rhoe);
    V[0] = (-E + rhoe * (gamma + 1.0 - s)) * ir;
#pragma unroll
    for (int d = 0; d < DIM; ++d) V[1 + d] = U[1 + d] * ir;
    V[1 + DIM] = -rho * ir;
}

// entropy variables -> conservative
template <int DIM>
__device__ __forceinline__ void u_of_v(const double* V, double gamma, double lg1, double* U) {
    double v2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) v2 = fma(V[1 + d], V[1 + d], v2);
    const double nvE = -V[1 + DIM];
    const double h = -.5 * v2 * fast_rcp(nvE);          // |vU|^2 / (2 vE)
    const double s = gamma - V[0] + h;
    const double ig = 1.0 / (gamma - 1.0);
    // ((g-1)/(-vE)^g)^(1/(g-1)) * exp(-s/(g-1))
    const double rhoeV = exp_mod((lg1 - gamma * log_pos(nvE > 0.0 ? nvE : 1.0) - s) * ig);
    U[0] = rhoeV * nvE;
#pragma unroll
    for (int d = 0; d < DIM; ++d) U[1 + d] = rhoeV * V[1 + d];
    U[1 + DIM] = rhoeV * (1.0 - h);
}

// conservative -> (rho, u, beta) + logs.  betafun: beta = rho/(2p), p=(gamma-1)(E-.5|rhoU|^2/rho)
template <int DIM>
__device__ __forceinline__ void prim_of_u(const double* U, double gamma, Prim<DIM>& P) {
    const double rho = U[0];
    const double ir = fast_rcp(rho);
    double m2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        m2 = fma(U[1 + d], U[1 + d], m2);
        P.vel[d] = U[1 + d] * ir;
    }
    const double p = (gamma - 1.0) * (U[1 + DIM] - .5 * m2 * ir);
    P.rho = rho;
    P.beta = .5 * rho * fast_rcp(p);
    const bool ok = (rho > 0.0) && (p > 0.0);
    P.lrho = log_pos(ok ? rho : 1.0);
    P.lbeta = log_pos(ok ? P.beta : 1.0);
}

// |wavespeed(rho, rhoU_n, E)|, euler_variables.jl:7-10 (sic: sqrt of |u_n|)
__device__ __forceinline__ double wavespeed_1d(double rho, double rhoun, double E, double gamma) {
    const double ir = fast_rcp(rho);
    const double p = (gamma - 1.0) * (E - .5 * rhoun * rhoun * ir);
    return fabs(fast_sqrt(fabs(rhoun * ir)) + fast_sqrt(gamma * p * ir));
}

}  // namespace esdg
