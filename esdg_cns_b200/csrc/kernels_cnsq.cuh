// Compressible Navier-Stokes on tensor-product quads (BASELINE config 3) in the GAUSS-NODAL basis.
//
// The CNS scripts (examples/CompressibleNS/dg2D_CNS_cavity_optimized.jl) keep modal / GLL-nodal coefficients and move to
// the (N+1)^2 Gauss quadrature points with Vq, back with Pq.  On a quad that quadrature has as many points as the basis
// has functions, so Vq is square and invertible and Pq = Vq^-1: the change of variables Q~ = Vq Q turns every operator
// of rhsRK! (:955-972) into a Kronecker product of 1-D Gauss-node tables --
//   Vq Pq = I,  Vf Pq = Ef (end-point interpolation along a grid line),  Vq Dr Pq = D (x) I,  Vq LIFT = W^-1 Ef' Wf,
//   Vq Ph = W^-1 [I; Ef]'
// -- i.e. the scheme IS the collocated one of dg2D_euler_quad.jl plus line-local viscous terms.  esdg_finalize proves
// these identities for the caller's operators (quad_fast.cu:cnsq_try_create), the state is converted once at upload
// (k_basis_change with Vq) and back at download (Pq), and one rhs evaluation is TWO launches without a single GEMM:
//   k_cnsq_sigma          theta = BR1 gradient of the entropy variables (:548-569, BCs :178-216), sigma = K(v) theta
//                         (:613-645, :786-800), sigma at the nodes and at the face points, rhstest_visc volume term
//   k_quad_stage_line<n,true>   inviscid flux differencing with the wall mirror state, BR1 divergence, stress BCs,
//                         penalisation, rhs or fused LSRK stage, traces of the new state   (kernels_quad_line.cuh)
// The entropy-variable traces the viscous terms need (v2..v4 at the face points) follow from the primitive traces the
// inviscid part already exchanges: v4 = -2 (gamma-1) beta, v2 = -v4 u, v3 = -v4 v.
#pragma once
#include "kernels_quad_line.cuh"

namespace esdg {

template <int n>
struct CnsqDims {
    using Q = QuadDims<n>;
    static constexpr int THREADS = Q::THREADS, EB = Q::EB, EN = Q::EN, EF = Q::EF;
    // V[3][EN] A[2][6][EN] (sigma[6][EN] aliases block 0) P[3][EF] geo[EB][17] lm lp w
    static constexpr int SMEM_DOUBLES = 3 * EN + 12 * EN + 3 * EF + EB * QUAD_GEO + 3 * n;
};

template <int n>
__global__ void __launch_bounds__(CnsqDims<n>::THREADS, 512 / CnsqDims<n>::THREADS) k_cnsq_sigma(const __grid_constant__ QuadParams p) {
    using D = QuadDims<n>;
    constexpr int Nq = D::Nq, Nfq = D::Nfq, NL = D::NL, EB = D::EB, NT = D::THREADS, EN = D::EN, EF = D::EF;
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[NT / 32];
    double* sV = sm;                        // [3][EN] v2, v3, v4 at the nodes
    double* sA = sV + 3 * EN;               // [2][6][EN] theta numerators by direction; block 0 later holds sigma
    double* sP = sA + 12 * EN;              // [3][EF] exterior (u/2, v/2, beta)
    double* sGeo = sP + 3 * EF;
    double* s_lm = sGeo + EB * QUAD_GEO;
    double* s_lp = s_lm + n;
    double* s_w = s_lp + n;
    const int tid = threadIdx.x;
    const long long e0 = p.e_begin + (long long)blockIdx.x * EB;
    const int nb = (int)min((long long)EB, p.e_begin + p.e_count - e0);
    const long long gq0 = e0 * Nq, gf0 = e0 * Nfq;
    // every independent request first (mapP indices, Q of this thread's nodes), their uses after
    constexpr int FIT = (EF + NT - 1) / NT, AIT = (EN + NT - 1) / NT;
    int ipv[FIT];
#pragma unroll
    for (int k = 0; k < FIT; ++k) {
        const int it = tid + k * NT;
        ipv[k] = it < nb * Nfq ? p.mapP[gf0 + it] : 0;
    }
    double Ua[AIT][4];
#pragma unroll
    for (int k = 0; k < AIT; ++k) {
        const int it = tid + k * NT;
#pragma unroll
        for (int f = 0; f < 4; ++f) Ua[k][f] = it < nb * Nq ? p.Q[f * p.q_stride + gq0 + it] : 1.0;
    }
#pragma unroll
    for (int k = 0; k < FIT; ++k) {
        const int it = tid + k * NT;
        if (it < nb * Nfq) {
#pragma unroll
            for (int q = 0; q < 3; ++q) cp_async8(sP + q * EF + it, p.Tcur + (1 + q) * p.t_stride + ipv[k]);
        }
    }
    for (int it = tid; it < nb * QUAD_GEO; it += NT) cp_async8(sGeo + it, p.geo + e0 * QUAD_GEO + it);
    cp_async_commit();
    if (tid < n) { s_lm[tid] = p.lm[tid]; s_lp[tid] = p.lp[tid]; s_w[tid] = p.w[tid]; }
    bool bad = false;
    // ---- nodes: v2..v4 = (rho u, rho v, -rho) / rho e                                  (euler_variables.jl:79-92)
#pragma unroll
    for (int k = 0; k < AIT; ++k) {
        const int it = tid + k * NT;
        if (it < nb * Nq) {
            VD<1> U[4], hu[2], beta, rhoe;
#pragma unroll
            for (int f = 0; f < 4; ++f) U[f][0] = Ua[k][f];
            vol_prim2_v<1>(U, p.gamma, hu, beta, rhoe);
            bad = bad || !(U[0][0] > 0.0) || !(rhoe[0] > 0.0);
            const VD<1> ire = vrcp(rhoe);
            sV[it] = U[1][0] * ire[0];
            sV[EN + it] = U[2][0] * ire[0];
            sV[2 * EN + it] = -(U[0][0] * ire[0]);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- lines: G D v + lift(.5 (v+ - v) nJ) of this direction                         (dg_grad!, :548-569)
    if (tid < nb * NL) {
        const int eb = tid / NL, ll = tid - eb * NL;
        int d, base, stride, f0, f1;
        quad_line<n>(ll, d, base, stride, f0, f1);
        base += eb * Nq;
        const double* Gm = sGeo + eb * QUAD_GEO;
        const double gx = Gm[d], gy = Gm[2 + d];
        double V[n][3], dV[n][3];
#pragma unroll
        for (int a = 0; a < n; ++a)
#pragma unroll
            for (int f = 0; f < 3; ++f) { V[a][f] = sV[f * EN + base + a * stride]; dV[a][f] = 0.0; }
#pragma unroll
        for (int b = 0; b < n; ++b)
#pragma unroll
            for (int a = 0; a < n; ++a)
#pragma unroll
                for (int f = 0; f < 3; ++f) dV[a][f] = fma(p.D[a * n + b], V[b][f], dV[a][f]);
        double jx[2][3], jy[2][3];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const double* l = side ? s_lp : s_lm;
            double Vf[4] = {0, 0, 0, 0}, VP[4];
#pragma unroll
            for (int a = 0; a < n; ++a)
#pragma unroll
                for (int f = 0; f < 3; ++f) Vf[1 + f] = fma(l[a], V[a][f], Vf[1 + f]);
            const int fi = eb * Nfq + (side ? f1 : f0);
            v234_from_prim(sP[fi], sP[EF + fi], sP[2 * EF + fi], p.gamma, VP);
            const double* gf = Gm + 5 + ((side ? f1 : f0) / n) * 3;
            const int kd = p.bc[gf0 + fi] & 7;
            if (kd == 1 || kd == 2) {
                const double isJ = fast_rcp(gf[2]);
                bc_entropy_vars_q(p.bctype, kd, kd == 2 ? p.vlid[gf0 + fi] : 0.0, p.inv_theta2, gf[0] * isJ, gf[1] * isJ, Vf, VP);
            }
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const double h = .5 * (VP[1 + f] - Vf[1 + f]);
                jx[side][f] = h * gf[0];
                jy[side][f] = h * gf[1];
            }
        }
        double* A = sA + d * 6 * EN;
#pragma unroll
        for (int a = 0; a < n; ++a)
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                A[f * EN + base + a * stride] = fma(gx, dV[a][f], fma(p.lmw[a], jx[0][f], p.lpw[a] * jx[1][f]));
                A[(3 + f) * EN + base + a * stride] = fma(gy, dV[a][f], fma(p.lmw[a], jy[0][f], p.lpw[a] * jy[1][f]));
            }
    }
    __syncthreads();
    // ---- nodes: theta = (.)/J, sigma = K(v) theta (:786-800), rhstest_visc volume term (:802-806)
    double tsum = 0.0;
    for (int it = tid; it < nb * Nq; it += NT) {
        const long long g = gq0 + it;
        const int eb = it / Nq, i = it - eb * Nq;
        const double J = sGeo[eb * QUAD_GEO + 4], iJ = fast_rcp(J);
        double tx[3], ty[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            tx[f] = (sA[f * EN + it] + sA[(6 + f) * EN + it]) * iJ;
            ty[f] = (sA[(3 + f) * EN + it] + sA[(9 + f) * EN + it]) * iJ;
        }
        const double v2 = sV[it], v3 = sV[EN + it], v4 = sV[2 * EN + it];
        const double lam = p.lam, mu = p.mu, l2m = lam + 2.0 * mu;
        const double ic = fast_rcp(v4 * v4 * v4);
        const double gmp = p.gamma * mu / p.Pr;
        const double v44 = v4 * v4;
        // Kxx, Kxy, Kyy rows / columns 2..4 (indices 0..2 here), :613-645
        const double xx00 = ic * -l2m * v44, xx02 = ic * l2m * v2 * v4, xx11 = ic * -mu * v44, xx12 = ic * mu * v3 * v4;
        const double xx22 = ic * -(l2m * v2 * v2 + mu * v3 * v3 - gmp * v4);
        const double xy01 = ic * -lam * v44, xy02 = ic * lam * v3 * v4, xy10 = ic * -mu * v44, xy12 = ic * mu * v2 * v4;
        const double xy20 = ic * mu * v3 * v4, xy21 = ic * lam * v2 * v4, xy22 = ic * (lam + mu) * (-v2 * v3);
        const double yy00 = ic * -mu * v44, yy02 = ic * mu * v2 * v4, yy11 = ic * -l2m * v44, yy12 = ic * l2m * v3 * v4;
        const double yy22 = ic * -(l2m * v3 * v3 + mu * v2 * v2 - gmp * v4);
        double sx[3], sy[3];
        sx[0] = xx00 * tx[0] + xy01 * ty[1] + (xx02 * tx[2] + xy02 * ty[2]);
        sx[1] = xy10 * ty[0] + xx11 * tx[1] + (xx12 * tx[2] + xy12 * ty[2]);
        sx[2] = (xx02 * tx[0] + xy20 * ty[0]) + (xx12 * tx[1] + xy21 * ty[1]) + (xx22 * tx[2] + xy22 * ty[2]);
        sy[0] = yy00 * ty[0] + xy10 * tx[1] + (xy20 * tx[2] + yy02 * ty[2]);
        sy[1] = xy01 * tx[0] + yy11 * ty[1] + (xy21 * tx[2] + yy12 * ty[2]);
        sy[2] = (xy02 * tx[0] + yy02 * ty[0]) + (xy12 * tx[1] + yy12 * ty[1]) + (xy22 * tx[2] + yy22 * ty[2]);
        if (p.want_rhstest) {
            const double wJ = s_w[i % n] * s_w[i / n] * J;
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                tsum += wJ * tx[f] * sx[f];
                tsum += wJ * ty[f] * sy[f];
            }
        }
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            p.sig[f * p.q_stride + g] = sx[f];
            p.sig[(3 + f) * p.q_stride + g] = sy[f];
            sA[f * EN + it] = sx[f];          // same thread read these slots above
            sA[(3 + f) * EN + it] = sy[f];
        }
    }
    __syncthreads();
    // ---- lines: sigma at the two face points of the line -> sigma traces                (:811-812)
    if (tid < nb * NL) {
        const int eb = tid / NL, ll = tid - eb * NL;
        int d, base, stride, f0, f1;
        quad_line<n>(ll, d, base, stride, f0, f1);
        base += eb * Nq;
        double* T0 = p.sig_tr + (e0 + eb) * Nfq;
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int a = 0; a < n; ++a) {
                const double x = sA[f * EN + base + a * stride];
                s0 = fma(s_lm[a], x, s0);
                s1 = fma(s_lp[a], x, s1);
            }
            T0[f * p.t_stride + f0] = s0;
            T0[f * p.t_stride + f1] = s1;
        }
    }
    if (p.want_rhstest) {
        for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(0xffffffffu, tsum, o);
        if ((tid & 31) == 0) red[tid >> 5] = tsum;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int wv = 0; wv < NT / 32; ++wv) t += red[wv];
            p.partial_visc[blockIdx.x] = t;
        }
    }
    if (bad) atomicOr(p.bad, 1);
}

// dst[f][e][r] = sum_q M[r][q] src[f][e][q]  (M row-major, m x m): the change of basis between the caller's coefficients
// and the Gauss-nodal values (Vq at upload, Pq at download / for rhs output)
__global__ void k_basis_change(const double* __restrict__ M, int m, const double* __restrict__ src, double* __restrict__ dst, long long nvec) {
    extern __shared__ double sMat[];
    for (int i = threadIdx.x; i < m * m; i += blockDim.x) sMat[i] = M[i];
    __syncthreads();
    const long long total = nvec * m;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long v = idx / m;
        const int r = (int)(idx - v * m);
        const double* x = src + v * m;
        double acc = 0.0;
        for (int q = 0; q < m; ++q) acc = fma(sMat[r * m + q], x[q], acc);
        dst[idx] = acc;
    }
}

// DOPRI error norm in the caller's basis: partial sums of (|Pq sum_s E[s] K_s| / (errTol (1 + |Pq Q~|)))^2
// (dg2D_CNS_cavity_optimized.jl:1015-1023); one partial per CTA, fixed order
struct DopriErrArgs {
    const double* K[7];
    double c[7];
};
__global__ void k_dopri_err_basis(const double* __restrict__ M, int m, const double* __restrict__ Q, DopriErrArgs a, double err_tol,
                                  long long nvec, double* __restrict__ partial) {
    extern __shared__ double sMat[];
    __shared__ double sh[256];
    for (int i = threadIdx.x; i < m * m; i += blockDim.x) sMat[i] = M[i];
    __syncthreads();
    double t = 0.0;
    const long long total = nvec * m;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const long long v = idx / m;
        const int r = (int)(idx - v * m);
        double e = 0.0, qm = 0.0;
        for (int q = 0; q < m; ++q) {
            const long long i = v * m + q;
            double ev = 0.0;
            for (int s = 0; s < 7; ++s) ev = ev + a.c[s] * a.K[s][i];
            e = fma(sMat[r * m + q], ev, e);
            qm = fma(sMat[r * m + q], Q[i], qm);
        }
        const double sc = fabs(e) / (err_tol * (1.0 + fabs(qm)));
        t += sc * sc;
    }
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

}  // namespace esdg
