// Pointwise Euler physics in FP64 for sm_100a.
//
// Device versions of the functions the reference's rhs calls per point:
//   logmean        examples/EntropyStableEuler/logmean.jl:14-28
//   euler_fluxes   examples/EntropyStableEuler/euler_fluxes.jl:23-48 (2D), :51-89 (3D)
//   v_ufun/u_vfun  examples/EntropyStableEuler/euler_variables.jl:79-86 / :107-113
//   betafun        :30-39 ; wavespeed :7-10
// Arithmetic is arranged for the FP64 pipe (few divisions, no divergent
// branches); results agree with the reference formulas to a few ulp.
#pragma once
#include <cuda_runtime.h>

namespace esdg {

template <int DIM>
struct Prim {  // (rho, u, v[, w], beta) + logs, the per-point operand of the two-point flux
    double rho, vel[DIM], beta, lrho, lbeta;
};

// logmean(aL,aR,logL,logR) written as one division:
//   |f| < 1e-4 (f = da/aavg):  aavg*(1 + v*(-.2 - v*(.0512 - v*0.026038857142857))), v = f^2
//   else                    :  -da/(logL-logR) = da/(logR-logL)
// The quotient q = da/den with den = taylor ? aavg : (logR-logL) is f in the first
// case and the log mean itself in the second.
__device__ __forceinline__ double logmean(double aL, double aR, double logL, double logR) {
    const double da = aR - aL;
    const double aavg = .5 * (aR + aL);
    const bool taylor = fabs(da) < 1e-4 * aavg;
    const double den = taylor ? aavg : (logR - logL);
    const double q = da / den;
    const double v = q * q;
    const double t = aavg * (1.0 + v * (-.2 - v * (.0512 - v * 0.026038857142857)));
    return taylor ? t : q;
}

// Entropy-conservative two-point flux contracted with a direction vector c:
//   out[f] = sum_d c[d] * F_d[f](L,R)
// Contracting first is exact algebra on the reference's Fx,Fy,Fz:
//   F_d = (rholog*uavg_d, rholog*uavg_d*uavg_k + pa*delta_dk, Ep*uavg_d)
//   => sum_d c_d F_d = (m, m*uavg_k + pa*c_k, Ep*un)  with un = c.uavg, m = rholog*un.
template <int DIM>
__device__ __forceinline__ void ec_flux_dot(const Prim<DIM>& L, const Prim<DIM>& R, const double* c,
                                            double gm1x2, double* out) {
    const double rholog = logmean(L.rho, R.rho, L.lrho, R.lrho);
    const double betalog = logmean(L.beta, R.beta, L.lbeta, R.lbeta);
    const double rhoavg = .5 * (L.rho + R.rho);
    double uavg[DIM];
    double unorm = 0.0, un = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        uavg[d] = .5 * (L.vel[d] + R.vel[d]);
        unorm = fma(L.vel[d], R.vel[d], unorm);
        un = fma(c[d], uavg[d], un);
    }
    // pa = rhoavg/(betaL+betaR) ; e1 = rholog/(2(gamma-1)betalog): one reciprocal for both
    const double bs = L.beta + R.beta;
    const double bl = gm1x2 * betalog;
    const double rinv = 1.0 / (bs * bl);
    const double pa = rhoavg * bl * rinv;
    const double Ep = rholog * bs * rinv + pa + .5 * rholog * unorm;
    const double m = rholog * un;
    out[0] = m;
#pragma unroll
    for (int d = 0; d < DIM; ++d) out[1 + d] = fma(m, uavg[d], pa * c[d]);
    out[1 + DIM] = Ep * un;
}

// conservative (rho, rhoU, E) -> entropy variables
template <int DIM>
__device__ __forceinline__ void v_of_u(const double* U, double gamma, double* V, bool& bad) {
    const double rho = U[0];
    double m2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) m2 = fma(U[1 + d], U[1 + d], m2);
    const double E = U[1 + DIM];
    const double rhoe = E - .5 * m2 / rho;
    bad = bad || !(rho > 0.0) || !(rhoe > 0.0);
    const double s = log((gamma - 1.0) * rhoe / pow(rho, gamma));
    const double ir = 1.0 / rhoe;
    V[0] = (-E + rhoe * (gamma + 1.0 - s)) * ir;
#pragma unroll
    for (int d = 0; d < DIM; ++d) V[1 + d] = U[1 + d] * ir;
    V[1 + DIM] = -rho * ir;
}

// entropy variables -> conservative
template <int DIM>
__device__ __forceinline__ void u_of_v(const double* V, double gamma, double* U) {
    double v2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) v2 = fma(V[1 + d], V[1 + d], v2);
    const double vE = V[1 + DIM];
    const double h = v2 / (2.0 * vE);
    const double s = gamma - V[0] + h;
    const double ig = 1.0 / (gamma - 1.0);
    const double rhoeV = pow((gamma - 1.0) / pow(-vE, gamma), ig) * exp(-s * ig);
    U[0] = rhoeV * (-vE);
#pragma unroll
    for (int d = 0; d < DIM; ++d) U[1 + d] = rhoeV * V[1 + d];
    U[1 + DIM] = rhoeV * (1.0 - h);
}

// conservative -> (rho, u, beta) + logs.  betafun: beta = rho/(2p), p=(gamma-1)(E-.5|rhoU|^2/rho)
template <int DIM>
__device__ __forceinline__ void prim_of_u(const double* U, double gamma, Prim<DIM>& P) {
    const double rho = U[0];
    const double ir = 1.0 / rho;
    double m2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        m2 = fma(U[1 + d], U[1 + d], m2);
        P.vel[d] = U[1 + d] * ir;
    }
    const double p = (gamma - 1.0) * (U[1 + DIM] - .5 * m2 * ir);
    P.rho = rho;
    P.beta = rho / (2.0 * p);
    P.lrho = log(rho);
    P.lbeta = log(P.beta);
}

// |wavespeed(rho, rhoU_n, E)|, euler_variables.jl:7-10 (sic: sqrt of |u_n|)
__device__ __forceinline__ double wavespeed_1d(double rho, double rhoun, double E, double gamma) {
    const double ir = 1.0 / rho;
    const double p = (gamma - 1.0) * (E - .5 * rhoun * rhoun * ir);
    return fabs(sqrt(fabs(rhoun * ir)) + sqrt(gamma * p * ir));
}

}  // namespace esdg
