// 1-D compressible Navier-Stokes, modal ESDG + SSPRK(3,3): examples/CompressibleNS/dg1D_CNS_modalESDG.jl (SURVEY 8f
// rank 4).  The script's problem is tiny (N = 4, K = 128 intervals), so the mapping is one thread per element and the
// operators (all dense, <= 12 x 12) live in constant-sized device tables; what matters here is that the path exists
// behind the same C ABI and agrees with the reference formulas.  `rhs_ESDG` (:338-353) has the two inter-element
// dependences of every BR1 scheme, hence three kernels per evaluation:
//   k_traces : v = v_ufun(Vq U), entropy-projected face state (rho, u, beta), face conservative state, wave speed,
//              and the modal entropy variables at the faces                                         (:225-236, :273-279)
//   k_sigma  : theta = (Dr v + LIFT(.5 [[v]] n))/J, sigma = K(v) theta, Pq, face values            (:288-318)
//   k_rhs    : flux differencing with the dense hybridized operator, Lax-Friedrichs surface flux with the exact states
//              on the two domain ends, viscous divergence, and the SSPRK combination                (:204-267, :320-335, :379-391)
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/esdg_b200.h"
#include "physics.cuh"

namespace {
constexpr int MAXP = 8, MAXQ = 10, MAXH = 12;

struct Ops1D {
    int Np, Nq, Nh;
    double Vq[MAXQ * MAXP], Pq[MAXP * MAXQ], Vf[2 * MAXP], LIFT[MAXP * 2], Dr[MAXP * MAXP];
    double VhP[MAXH * MAXQ], Ph[MAXP * MAXH], Qsk[MAXH * MAXH];   // row-major copies
    double J, rxJ, gamma, c22 /* 2 mu - lambda */, kcv /* kappa / cv */;
};

struct Bc1D {   // exact-solution data at the two ends for one rhs evaluation
    double rhoL, uL, betaL, rhoR, uR, betaR, vL[3], vR[3];
};
}  // namespace

struct esdg1d_ctx {
    int device = 0;
    long long K = 0;
    Ops1D h{};
    Ops1D* ops = nullptr;
    double *U = nullptr, *W = nullptr, *rhs = nullptr;      // [3][Np*K]
    double *trU = nullptr;     // [7][2K]: rho, u, beta (entropy projected), rho, rhou, E (face conservative), lambda
    double *trV = nullptr;     // [3][2K]: modal entropy variables at the faces
    double *trS = nullptr;     // [3][2K]: sigma at the faces
    double *sig = nullptr;     // [3][Np*K]: modal sigma
    std::string err;
    long long launches = 0;
};

namespace {
using namespace esdg;

__device__ inline void v_ufun1(double rho, double rhou, double E, double gamma, double* v) {   // euler_variables / entropy_variables.jl
    const double rhoe = E - .5 * rhou * rhou / rho;
    const double s = log((gamma - 1.0) * rhoe / pow(rho, gamma));
    v[0] = (-E + rhoe * (gamma + 1.0 - s)) / rhoe;
    v[1] = rhou / rhoe;
    v[2] = -rho / rhoe;
}
__device__ inline void u_vfun1(const double* v, double gamma, double* u) {
    const double s = gamma - v[0] + v[1] * v[1] / (2.0 * v[2]);
    const double rhoe = pow((gamma - 1.0) / pow(-v[2], gamma), 1.0 / (gamma - 1.0)) * exp(-s / (gamma - 1.0));
    u[0] = rhoe * (-v[2]);
    u[1] = rhoe * v[1];
    u[2] = rhoe * (1.0 - v[1] * v[1] / (2.0 * v[2]));
}
__device__ inline double logmean1(double aL, double aR, double logL, double logR) {   // logmean.jl:14-28
    const double da = aR - aL, aavg = .5 * (aR + aL), f = da / aavg, v = f * f;
    if (fabs(f) < 1e-4) return aavg * (1.0 + v * (-.2 - v * (.0512 - v * 0.026038857142857)));
    return -da / (logL - logR);
}
// euler_fluxes_1D (:39-60) on (rho, u, beta) with the logs supplied
__device__ inline void flux1(const double* L, const double* R, double lrL, double lbL, double lrR, double lbR, double gamma, double* F) {
    const double rholog = logmean1(L[0], R[0], lrL, lrR), betalog = logmean1(L[2], R[2], lbL, lbR);
    const double rhoavg = .5 * (L[0] + R[0]), uavg = .5 * (L[1] + R[1]);
    const double pa = rhoavg / (L[2] + R[2]);
    const double f4 = rholog / (2.0 * (gamma - 1.0) * betalog) + pa + .5 * rholog * (L[1] * R[1]);
    F[0] = rholog * uavg;
    F[1] = F[0] * uavg + pa;
    F[2] = f4 * uavg;
}

// entropy projection of element e: Uh (conservative, Nh points) and the modal entropy variables VU (Np)
__device__ inline void project(const Ops1D& o, const double* U, long long stride, long long e, double (*Uh)[MAXH], double (*VUm)[MAXP]) {
    double vq[3][MAXQ];
    for (int q = 0; q < o.Nq; ++q) {
        double u[3] = {0, 0, 0};
        for (int c = 0; c < 3; ++c)
            for (int j = 0; j < o.Np; ++j) u[c] += o.Vq[q * o.Np + j] * U[c * stride + e * o.Np + j];
        double v[3];
        v_ufun1(u[0], u[1], u[2], o.gamma, v);
        for (int c = 0; c < 3; ++c) vq[c][q] = v[c];
    }
    for (int i = 0; i < o.Nh; ++i) {
        double v[3] = {0, 0, 0}, u[3];
        for (int c = 0; c < 3; ++c)
            for (int q = 0; q < o.Nq; ++q) v[c] += o.VhP[i * o.Nq + q] * vq[c][q];
        u_vfun1(v, o.gamma, u);
        for (int c = 0; c < 3; ++c) Uh[c][i] = u[c];
    }
    for (int c = 0; c < 3; ++c)
        for (int j = 0; j < o.Np; ++j) {
            double a = 0;
            for (int q = 0; q < o.Nq; ++q) a += o.Pq[j * o.Nq + q] * vq[c][q];
            VUm[c][j] = a;
        }
}

__global__ void k1d_traces(const Ops1D* __restrict__ op, const double* __restrict__ U, long long K, double* trU, double* trV) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K) return;
    const Ops1D& o = *op;
    double Uh[3][MAXH], VUm[3][MAXP];
    project(o, U, o.Np * K, e, Uh, VUm);
    for (int f = 0; f < 2; ++f) {
        const int i = o.Nq + f;
        const double rho = Uh[0][i], rhou = Uh[1][i], E = Uh[2][i];
        const double p = (o.gamma - 1.0) * (E - .5 * rhou * rhou / rho);
        const long long t = 2 * e + f;
        trU[t] = rho;
        trU[2 * K + t] = rhou / rho;
        trU[4 * K + t] = rho / (2.0 * p);
        trU[6 * K + t] = rho;
        trU[8 * K + t] = rhou;
        trU[10 * K + t] = E;
        const double rhoun = rhou * (f ? 1.0 : -1.0);                      // wavespeed_1D(rhoM, rhouM*nxJ, EM), :253-255
        trU[12 * K + t] = fabs(fabs(rhoun / rho) + sqrt(o.gamma * p / rho));
        for (int c = 0; c < 3; ++c) {
            double a = 0;
            for (int j = 0; j < o.Np; ++j) a += o.Vf[f * o.Np + j] * VUm[c][j];
            trV[c * 2 * K + t] = a;
        }
    }
}

__global__ void k1d_sigma(const Ops1D* __restrict__ op, const double* __restrict__ U, long long K, const double* __restrict__ trV, Bc1D bc,
                          double* sig, double* trS) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K) return;
    const Ops1D& o = *op;
    double Uh[3][MAXH], VUm[3][MAXP];
    project(o, U, o.Np * K, e, Uh, VUm);
    double th[3][MAXP];
    for (int c = 0; c < 3; ++c) {
        double jump[2];
        for (int f = 0; f < 2; ++f) {
            const long long t = 2 * e + f;
            double vP;
            if (f == 0) vP = e == 0 ? bc.vL[c] : trV[c * 2 * K + t - 1];
            else vP = e == K - 1 ? bc.vR[c] : trV[c * 2 * K + t + 1];
            jump[f] = .5 * (vP - trV[c * 2 * K + t]) * (f ? 1.0 : -1.0);
        }
        for (int j = 0; j < o.Np; ++j) {
            double a = 0;
            for (int m = 0; m < o.Np; ++m) a += o.Dr[j * o.Np + m] * VUm[c][m];
            a += o.LIFT[j * 2] * jump[0] + o.LIFT[j * 2 + 1] * jump[1];
            th[c][j] = a / o.J;
        }
    }
    double sq[3][MAXQ];
    for (int q = 0; q < o.Nq; ++q) {
        double v[3] = {0, 0, 0}, tx[3] = {0, 0, 0};
        for (int c = 0; c < 3; ++c)
            for (int j = 0; j < o.Np; ++j) {
                v[c] += o.Vq[q * o.Np + j] * VUm[c][j];
                tx[c] += o.Vq[q * o.Np + j] * th[c][j];
            }
        const double v2 = v[1], v4 = v[2];
        const double K22 = -o.c22 / v4, K23 = o.c22 * v2 / (v4 * v4);
        const double K33 = -o.c22 * v2 * v2 / (v4 * v4 * v4) + o.kcv / (v4 * v4);
        sq[0][q] = 0.0;
        sq[1][q] = K22 * tx[1] + K23 * tx[2];
        sq[2][q] = K23 * tx[1] + K33 * tx[2];
    }
    for (int c = 0; c < 3; ++c) {
        double sm[MAXP];
        for (int j = 0; j < o.Np; ++j) {
            double a = 0;
            for (int q = 0; q < o.Nq; ++q) a += o.Pq[j * o.Nq + q] * sq[c][q];
            sm[j] = a;
            sig[c * o.Np * K + e * o.Np + j] = a;
        }
        for (int f = 0; f < 2; ++f) {
            double a = 0;
            for (int j = 0; j < o.Np; ++j) a += o.Vf[f * o.Np + j] * sm[j];
            trS[c * 2 * K + 2 * e + f] = a;
        }
    }
}

// out = alpha * U0 + beta * (S + dt * rhs(S))   (rhs_only: out = rhs)
__global__ void k1d_rhs(const Ops1D* __restrict__ op, const double* __restrict__ S, const double* __restrict__ U0, long long K,
                        const double* __restrict__ trU, const double* __restrict__ trS, const double* __restrict__ sig, Bc1D bc, double alpha,
                        double beta, double dt, int rhs_only, double* out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K) return;
    const Ops1D& o = *op;
    const long long stride = o.Np * K;
    double Uh[3][MAXH], VUm[3][MAXP];
    project(o, S, stride, e, Uh, VUm);
    double Ub[3][MAXH], lr[MAXH], lb[MAXH];
    for (int i = 0; i < o.Nh; ++i) {
        const double rho = Uh[0][i], rhou = Uh[1][i], E = Uh[2][i];
        const double p = (o.gamma - 1.0) * (E - .5 * rhou * rhou / rho);
        Ub[0][i] = rho;
        Ub[1][i] = rhou / rho;
        Ub[2][i] = rho / (2.0 * p);
        lr[i] = log(Ub[0][i]);
        lb[i] = log(Ub[2][i]);
    }
    double UF[3][MAXH];
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < o.Nh; ++i) UF[c][i] = 0.0;
    for (int j = 0; j < o.Nh; ++j)               // flux_differencing!, :204-218
        for (int i = j; i < o.Nh; ++i)
            if (i < o.Nq || j < o.Nq) {
                const double L[3] = {Ub[0][i], Ub[1][i], Ub[2][i]}, R[3] = {Ub[0][j], Ub[1][j], Ub[2][j]};
                double F[3];
                flux1(L, R, lr[i], lb[i], lr[j], lb[j], o.gamma, F);
                for (int c = 0; c < 3; ++c) {
                    const double upd = 2.0 * o.rxJ * o.Qsk[i * o.Nh + j] * F[c];
                    UF[c][i] += upd;
                    UF[c][j] -= upd;
                }
            }
    double flux[3][2], sjump[3][2];
    for (int f = 0; f < 2; ++f) {
        const long long t = 2 * e + f, tp = f == 0 ? (e == 0 ? t : t - 1) : (e == K - 1 ? t : t + 1);   // mapP (self at the ends)
        const double nx = f ? 1.0 : -1.0;
        double M[3] = {trU[t], trU[2 * K + t], trU[4 * K + t]}, P[3] = {trU[tp], trU[2 * K + tp], trU[4 * K + tp]};
        if (f == 0 && e == 0) { P[0] = bc.rhoL; P[1] = bc.uL; P[2] = bc.betaL; }
        if (f == 1 && e == K - 1) { P[0] = bc.rhoR; P[1] = bc.uR; P[2] = bc.betaR; }
        double F[3];
        flux1(M, P, log(M[0]), log(M[2]), log(P[0]), log(P[2]), o.gamma, F);
        const double LFc = .25 * fmax(trU[12 * K + t], trU[12 * K + tp]);
        for (int c = 0; c < 3; ++c) {
            flux[c][f] = F[c] * nx - LFc * (trU[(6 + 2 * c) * K + tp] - trU[(6 + 2 * c) * K + t]);
            sjump[c][f] = .5 * (trS[c * 2 * K + tp] - trS[c * 2 * K + t]) * nx;     // sxP = sxf at the ends: zero jump
        }
    }
    for (int c = 0; c < 3; ++c)
        for (int j = 0; j < o.Np; ++j) {
            double a = o.LIFT[j * 2] * flux[c][0] + o.LIFT[j * 2 + 1] * flux[c][1];
            double ph = 0;
            for (int i = 0; i < o.Nh; ++i) ph += o.Ph[j * o.Nh + i] * UF[c][i];
            const double rI = -(ph + a) / o.J;
            double dv = 0;
            for (int m = 0; m < o.Np; ++m) dv += o.Dr[j * o.Np + m] * sig[c * stride + e * o.Np + m];
            dv += o.LIFT[j * 2] * sjump[c][0] + o.LIFT[j * 2 + 1] * sjump[c][1];
            const double r = rI + dv / o.J;
            const long long g = c * stride + e * o.Np + j;
            out[g] = rhs_only ? r : alpha * U0[g] + beta * (S[g] + dt * r);
        }
}

int fail1(esdg1d_ctx* c, int code, const std::string& m) {
    if (c) c->err = m;
    return code;
}
#define CK1(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return fail1(ctx, ESDG_ECUDA, std::string(#call ": ") + cudaGetErrorString(e_)); \
    } while (0)

void to_rowmajor(const double* colmajor, int rows, int cols, double* out) {
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) out[i * cols + j] = colmajor[(size_t)j * rows + i];
}

Bc1D make_bc(const double* b) {
    Bc1D bc;
    bc.rhoL = b[0]; bc.uL = b[1]; bc.betaL = b[2]; bc.rhoR = b[3]; bc.uR = b[4]; bc.betaR = b[5];
    for (int c = 0; c < 3; ++c) { bc.vL[c] = b[6 + c]; bc.vR[c] = b[9 + c]; }
    return bc;
}

// one evaluation of rhs_ESDG on the state S; result combined into `out`
int eval(esdg1d_ctx* ctx, const double* S, const double* U0, const Bc1D& bc, double alpha, double beta, double dt, int rhs_only, double* out) {
    const int T = 64;
    const unsigned B = (unsigned)((ctx->K + T - 1) / T);
    k1d_traces<<<B, T>>>(ctx->ops, S, ctx->K, ctx->trU, ctx->trV);
    k1d_sigma<<<B, T>>>(ctx->ops, S, ctx->K, ctx->trV, bc, ctx->sig, ctx->trS);
    k1d_rhs<<<B, T>>>(ctx->ops, S, U0, ctx->K, ctx->trU, ctx->trS, ctx->sig, bc, alpha, beta, dt, rhs_only, out);
    ctx->launches += 3;
    CK1(cudaGetLastError());
    return ESDG_OK;
}
}  // namespace

extern "C" {

int esdg1d_create(int device, int N, int64_t K, int Nq, const double* Vq, const double* Pq, const double* Vf, const double* LIFT,
                  const double* Dr, const double* VhP, const double* Ph, const double* Qrh_skew, double J, double rxJ, double gamma,
                  double mu, double lambda, double kappa_over_cv, esdg1d_ctx** out) {
    if (!out) return ESDG_EINVAL;
    esdg1d_ctx* ctx = new esdg1d_ctx();
    *out = ctx;
    const int Np = N + 1, Nh = Nq + 2;
    if (N < 1 || Np > MAXP || Nq < 1 || Nq > MAXQ || Nh > MAXH || K < 1 || !Vq || !Pq || !Vf || !LIFT || !Dr || !VhP || !Ph || !Qrh_skew || !(J > 0))
        return fail1(ctx, ESDG_EINVAL, "esdg1d_create: bad argument (N <= 7, Nq <= 10)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail1(ctx, ESDG_ECUDA, "esdg1d_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail1(ctx, ESDG_EINVAL, "esdg1d_create: bad device ordinal");
    CK1(cudaSetDevice(device));
    ctx->device = device;
    ctx->K = K;
    Ops1D& o = ctx->h;
    o.Np = Np; o.Nq = Nq; o.Nh = Nh;
    to_rowmajor(Vq, Nq, Np, o.Vq);
    to_rowmajor(Pq, Np, Nq, o.Pq);
    to_rowmajor(Vf, 2, Np, o.Vf);
    to_rowmajor(LIFT, Np, 2, o.LIFT);
    to_rowmajor(Dr, Np, Np, o.Dr);
    to_rowmajor(VhP, Nh, Nq, o.VhP);
    to_rowmajor(Ph, Np, Nh, o.Ph);
    to_rowmajor(Qrh_skew, Nh, Nh, o.Qsk);
    o.J = J; o.rxJ = rxJ; o.gamma = gamma; o.c22 = 2.0 * mu - lambda; o.kcv = kappa_over_cv;
    const size_t ns = (size_t)3 * Np * K, nt = (size_t)2 * K;
    CK1(cudaMalloc(&ctx->ops, sizeof(Ops1D)));
    CK1(cudaMemcpy(ctx->ops, &o, sizeof(Ops1D), cudaMemcpyHostToDevice));
    CK1(cudaMalloc(&ctx->U, ns * 8));
    CK1(cudaMalloc(&ctx->W, ns * 8));
    CK1(cudaMalloc(&ctx->rhs, ns * 8));
    CK1(cudaMalloc(&ctx->sig, ns * 8));
    CK1(cudaMalloc(&ctx->trU, 7 * nt * 8));
    CK1(cudaMalloc(&ctx->trV, 3 * nt * 8));
    CK1(cudaMalloc(&ctx->trS, 3 * nt * 8));
    CK1(cudaMemset(ctx->U, 0, ns * 8));
    return ESDG_OK;
}

int esdg1d_destroy(esdg1d_ctx* ctx) {
    if (!ctx) return ESDG_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (void* p : {(void*)ctx->ops, (void*)ctx->U, (void*)ctx->W, (void*)ctx->rhs, (void*)ctx->sig, (void*)ctx->trU, (void*)ctx->trV,
                    (void*)ctx->trS})
        if (p) cudaFree(p);
    delete ctx;
    return ESDG_OK;
}

const char* esdg1d_last_error(const esdg1d_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int esdg1d_upload_state(esdg1d_ctx* ctx, const double* const* U) {
    if (!ctx || !U) return ESDG_EINVAL;
    CK1(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->h.Np * ctx->K;
    for (int c = 0; c < 3; ++c) CK1(cudaMemcpy(ctx->U + c * n, U[c], n * 8, cudaMemcpyHostToDevice));
    return ESDG_OK;
}

int esdg1d_download_state(esdg1d_ctx* ctx, double* const* U) {
    if (!ctx || !U) return ESDG_EINVAL;
    CK1(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->h.Np * ctx->K;
    for (int c = 0; c < 3; ++c) CK1(cudaMemcpy(U[c], ctx->U + c * n, n * 8, cudaMemcpyDeviceToHost));
    return ESDG_OK;
}

int esdg1d_rhs(esdg1d_ctx* ctx, const double* bc12, double* const* rhs_out) {
    if (!ctx || !bc12 || !rhs_out) return ESDG_EINVAL;
    CK1(cudaSetDevice(ctx->device));
    int rc = eval(ctx, ctx->U, ctx->U, make_bc(bc12), 0.0, 0.0, 0.0, 1, ctx->rhs);
    if (rc) return rc;
    const size_t n = (size_t)ctx->h.Np * ctx->K;
    for (int c = 0; c < 3; ++c) CK1(cudaMemcpy(rhs_out[c], ctx->rhs + c * n, n * 8, cudaMemcpyDeviceToHost));
    return ESDG_OK;
}

int esdg1d_ssprk3_step(esdg1d_ctx* ctx, double dt, const double* bc12) {
    if (!ctx || !bc12) return ESDG_EINVAL;
    CK1(cudaSetDevice(ctx->device));
    const Bc1D bc = make_bc(bc12);
    int rc;
    // W = U + dt r(U);  W = 3/4 U + 1/4 (W + dt r(W));  U = 1/3 U + 2/3 (W + dt r(W))      (:383-391)
    if ((rc = eval(ctx, ctx->U, ctx->U, bc, 0.0, 1.0, dt, 0, ctx->W))) return rc;
    if ((rc = eval(ctx, ctx->W, ctx->U, bc, .75, .25, dt, 0, ctx->rhs))) return rc;
    if ((rc = eval(ctx, ctx->rhs, ctx->U, bc, 1.0 / 3.0, 2.0 / 3.0, dt, 0, ctx->W))) return rc;
    std::swap(ctx->U, ctx->W);
    CK1(cudaDeviceSynchronize());
    return ESDG_OK;
}

int64_t esdg1d_kernel_launches(const esdg1d_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
