/* esdg_oracle.c -- scalar C restatement of the reference rhs + LSRK45 (collocated quad / hex).
 *
 * TEST INFRASTRUCTURE ONLY (checker and CPU baseline; never linked into the product).
 * PARITY UNPINNED: the reference is Julia and ships no golden vectors; this file is
 * cross-checked against the independent numpy restatement (oracle/solvers.py).
 *
 * It follows the reference AS WRITTEN, loop for loop:
 *   rhs                  examples/dg2D_euler_quad.jl:141-194, examples/dg3D_euler_hex.jl:167-222
 *   sparse_hadamard_sum  examples/dg2D_euler_quad.jl:102-138, examples/dg3D_euler_hex.jl:122-164
 *                        (every (i,j) AND (j,i) of the union sparsity is evaluated, logs
 *                        hoisted per element, geofacs averaged per pair in 3D)
 *   logmean / fluxes     examples/EntropyStableEuler/logmean.jl:14-28, euler_fluxes.jl:23-89
 *   v_ufun / u_vfun ...  examples/EntropyStableEuler/euler_variables.jl:7-120
 *   LSRK45               examples/dg2D_euler_quad.jl:198-207, src/CommonUtils.jl:29-49
 * The only liberty is `#pragma omp parallel for` over elements / points (the reference is
 * single-threaded Julia); the thread count is reported by esdg_oracle_threads().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GAMMA 1.4
#define MAXF 5

typedef struct {
    int dim, nf, Nq, Nfq, Nh;
    int64_t K;
    /* CSR operators (0-based) */
    const int *ef_ptr, *ef_col; const double *ef_val;       /* Ef  Nfq x Nq */
    const int *ph_ptr, *ph_col; const double *ph_val;       /* Ph  Nq x Nh  */
    const int *lf_ptr, *lf_col; const double *lf_val;       /* Lf  Nq x Nfq */
    const int *q_ptr, *q_col; const double *q_val;          /* union pattern, q_val[3*k+d] */
    /* geometry: per element G[9] (rxJ,sxJ,txJ,ryJ,syJ,tyJ,rzJ,szJ,tzJ), J; per face point normals */
    const double *G, *J;
    const double *nxJ, *nyJ, *nzJ, *sJ;                       /* Nfq x K */
    const double *wq;
    const int64_t *mapP;                                      /* Nfq x K, 0-based */
    double lf_scale, volume_scale;
} oracle_desc;

/* bench.py --impl reference under torchrun inherits OMP_NUM_THREADS=1: let the caller ask for all host threads */
void esdg_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int esdg_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline double logmean(double aL, double aR, double logL, double logR) {
    double da = aR - aL;
    double aavg = .5 * (aR + aL);
    double f = da / aavg;
    double v = f * f;
    if (fabs(f) < 1e-4) return aavg * (1 + v * (-.2 - v * (.0512 - v * 0.026038857142857)));
    return -da / (logL - logR);
}

/* U = (rho,u,v[,w],beta); F[d][f] */
static inline void euler_fluxes(int dim, const double* UL, const double* UR, double lrL, double lbL, double lrR,
                                double lbR, double F[3][MAXF]) {
    double rhoL = UL[0], rhoR = UR[0], betaL = UL[dim + 1], betaR = UR[dim + 1];
    double rholog = logmean(rhoL, rhoR, lrL, lrR);
    double betalog = logmean(betaL, betaR, lbL, lbR);
    double rhoavg = .5 * (rhoL + rhoR);
    double uavg[3], unorm = 0;
    for (int d = 0; d < dim; ++d) {
        uavg[d] = .5 * (UL[1 + d] + UR[1 + d]);
        unorm += UL[1 + d] * UR[1 + d];
    }
    double pa = rhoavg / (betaL + betaR);
    double Ep = rholog / (2 * (GAMMA - 1) * betalog) + pa + .5 * rholog * unorm;
    for (int d = 0; d < dim; ++d) {
        double Fd1 = rholog * uavg[d];
        F[d][0] = Fd1;
        for (int k = 0; k < dim; ++k) F[d][1 + k] = (rholog * uavg[k < d ? k : d]) * uavg[k < d ? d : k] + (k == d ? pa : 0.0);
        F[d][dim + 1] = Ep * uavg[d];
    }
}

static inline void v_ufun(int dim, const double* U, double* V) {
    double rho = U[0], E = U[dim + 1], m2 = 0;
    for (int d = 0; d < dim; ++d) m2 += U[1 + d] * U[1 + d];
    double rhoe = E - .5 * m2 / rho;
    double s = log((GAMMA - 1) * rhoe / pow(rho, GAMMA));
    V[0] = (-E + rhoe * (GAMMA + 1 - s)) / rhoe;
    for (int d = 0; d < dim; ++d) V[1 + d] = U[1 + d] / rhoe;
    V[dim + 1] = (-rho) / rhoe;
}

static inline void u_vfun(int dim, const double* V, double* U) {
    double vE = V[dim + 1], v2 = 0;
    for (int d = 0; d < dim; ++d) v2 += V[1 + d] * V[1 + d];
    double s = GAMMA - V[0] + v2 / (2 * vE);
    double rhoeV = pow((GAMMA - 1) / pow(-vE, GAMMA), 1 / (GAMMA - 1)) * exp(-s / (GAMMA - 1));
    U[0] = rhoeV * (-vE);
    for (int d = 0; d < dim; ++d) U[1 + d] = rhoeV * V[1 + d];
    U[dim + 1] = rhoeV * (1 - v2 / (2 * vE));
}

static inline double betafun(int dim, const double* U) {
    double m2 = 0;
    for (int d = 0; d < dim; ++d) m2 += U[1 + d] * U[1 + d];
    double p = (GAMMA - 1) * (U[dim + 1] - .5 * (m2 / U[0]));
    return U[0] / (2 * p);
}

static inline double wavespeed(double rho, double rhou, double E) {
    double p = (GAMMA - 1) * (E - .5 * (rhou * rhou / rho));
    return sqrt(fabs(rhou / rho)) + sqrt(GAMMA * p / rho);
}

/* Q, rhs: nf arrays of Nq*K (field-major, contiguous: Q + f*Nq*K).  work: caller-provided scratch of
 * esdg_oracle_work_doubles() doubles.  Returns rhstest (0 if !compute_rhstest). */
int64_t esdg_oracle_work_doubles(const oracle_desc* d) {
    return (int64_t)d->K * ((int64_t)d->nf * d->Nq        /* VU */
                            + (int64_t)d->nf * d->Nfq     /* Uf */
                            + (int64_t)(d->nf) * d->Nh    /* Qh */
                            + (int64_t)d->Nfq             /* lam */
                            + (int64_t)d->nf * d->Nfq);   /* flux */
}

double esdg_oracle_rhs(const oracle_desc* d, const double* Q, double* rhs, double* work, int compute_rhstest) {
    const int dim = d->dim, nf = d->nf, Nq = d->Nq, Nfq = d->Nfq, Nh = d->Nh;
    const int64_t K = d->K;
    const int64_t sQ = (int64_t)Nq * K, sF = (int64_t)Nfq * K, sH = (int64_t)Nh * K;
    double* VU = work;
    double* Uf = VU + nf * sQ;
    double* Qh = Uf + nf * sF;
    double* lam = Qh + nf * sH;
    double* flux = lam + sF;
    /* VU = v_ufun(Q...) ; Uf = u_vfun(Ef*VU...) ; Qh = (rho, rhoU./rho, beta) at [Q;Uf]   (:148-155) */
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < K; ++e) {
        for (int i = 0; i < Nq; ++i) {
            double U[MAXF], V[MAXF];
            for (int f = 0; f < nf; ++f) U[f] = Q[f * sQ + e * Nq + i];
            v_ufun(dim, U, V);
            for (int f = 0; f < nf; ++f) VU[f * sQ + e * Nq + i] = V[f];
            double b = betafun(dim, U);
            Qh[0 * sH + e * Nh + i] = U[0];
            for (int k = 0; k < dim; ++k) Qh[(1 + k) * sH + e * Nh + i] = U[1 + k] / U[0];
            Qh[(dim + 1) * sH + e * Nh + i] = b;
        }
        for (int r = 0; r < Nfq; ++r) {
            double V[MAXF] = {0, 0, 0, 0, 0}, U[MAXF];
            for (int k = d->ef_ptr[r]; k < d->ef_ptr[r + 1]; ++k)
                for (int f = 0; f < nf; ++f) V[f] += d->ef_val[k] * VU[f * sQ + e * Nq + d->ef_col[k]];
            u_vfun(dim, V, U);
            for (int f = 0; f < nf; ++f) Uf[f * sF + e * Nfq + r] = U[f];
            double b = betafun(dim, U);
            Qh[0 * sH + e * Nh + Nq + r] = U[0];
            for (int k = 0; k < dim; ++k) Qh[(1 + k) * sH + e * Nh + Nq + r] = U[1 + k] / U[0];
            Qh[(dim + 1) * sH + e * Nh + Nq + r] = b;
            /* lam = abs(wavespeed(rhoM, rhoUM_n, EM))   (:162-164) */
            int64_t g = e * Nfq + r;
            double un = U[1] * d->nxJ[g] + U[2] * d->nyJ[g];
            if (dim == 3) un += U[3] * d->nzJ[g];
            un /= d->sJ[g];
            lam[g] = fabs(wavespeed(U[0], un, U[dim + 1]));
        }
    }
    /* surface flux (:158-170) */
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < K; ++e) {
        for (int r = 0; r < Nfq; ++r) {
            int64_t g = e * Nfq + r;
            int64_t gp = d->mapP[g];
            int64_t ep = gp / Nfq;
            int rp = (int)(gp - ep * Nfq);
            double QM[MAXF], QP[MAXF];
            for (int f = 0; f < nf; ++f) {
                QM[f] = Qh[f * sH + e * Nh + Nq + r];
                QP[f] = Qh[f * sH + ep * Nh + Nq + rp];
            }
            double F[3][MAXF];
            /* flux_fun(QM,QP) recomputes the logs (euler_fluxes.jl:9-15) */
            euler_fluxes(dim, QM, QP, log(QM[0]), log(QM[dim + 1]), log(QP[0]), log(QP[dim + 1]), F);
            double LFc = d->lf_scale * fmax(lam[g], lam[gp]) * d->sJ[g];
            for (int f = 0; f < nf; ++f) {
                double fn = F[0][f] * d->nxJ[g] + F[1][f] * d->nyJ[g];
                if (dim == 3) fn += F[2][f] * d->nzJ[g];
                flux[f * sF + g] = fn - LFc * (Uf[f * sF + gp] - Uf[f * sF + g]);
            }
        }
    }
    double rhstest = 0.0;
    /* rhsQ = Lf*flux ; volume Hadamard sums ; += vs*Ph*QF ; -./J   (:170-184) */
#pragma omp parallel for schedule(static) reduction(+ : rhstest)
    for (int64_t e = 0; e < K; ++e) {
        double Qe[MAXF][512], lr[512], lb[512], QF[MAXF][512];
        for (int i = 0; i < Nh; ++i) {
            for (int f = 0; f < nf; ++f) Qe[f][i] = Qh[f * sH + e * Nh + i];
            lr[i] = log(Qe[0][i]);          /* Qlog = (log.(rho), log.(beta))  (:111) */
            lb[i] = log(Qe[dim + 1][i]);
        }
        const double* G = d->G + 9 * e;
        for (int i = 0; i < Nh; ++i) {
            double rhsi[MAXF] = {0, 0, 0, 0, 0};
            double Ui[MAXF];
            for (int f = 0; f < nf; ++f) Ui[f] = Qe[f][i];
            for (int k = d->q_ptr[i]; k < d->q_ptr[i + 1]; ++k) {
                int j = d->q_col[k];
                double Uj[MAXF];
                for (int f = 0; f < nf; ++f) Uj[f] = Qe[f][j];
                double F[3][MAXF];
                euler_fluxes(dim, Ui, Uj, lr[i], lb[i], lr[j], lb[j], F);
                const double qr = d->q_val[3 * k], qs = d->q_val[3 * k + 1], qt = d->q_val[3 * k + 2];
                if (dim == 2) {
                    for (int f = 0; f < nf; ++f) {
                        double Fr = G[0] * F[0][f] + G[3] * F[1][f];   /* rxJ*Fx + ryJ*Fy */
                        double Fs = G[1] * F[0][f] + G[4] * F[1][f];   /* sxJ*Fx + syJ*Fy */
                        rhsi[f] += qr * Fr + qs * Fs;
                    }
                } else {
                    double g[9];
                    for (int q = 0; q < 9; ++q) g[q] = .5 * (G[q] + G[q]);   /* avg(vgeo_i,vgeo_j): affine => equal (:145-147) */
                    for (int f = 0; f < nf; ++f) {
                        double Fr = g[0] * F[0][f] + g[3] * F[1][f] + g[6] * F[2][f];
                        double Fs = g[1] * F[0][f] + g[4] * F[1][f] + g[7] * F[2][f];
                        double Ft = g[2] * F[0][f] + g[5] * F[1][f] + g[8] * F[2][f];
                        rhsi[f] += qr * Fr + qs * Fs + qt * Ft;
                    }
                }
            }
            for (int f = 0; f < nf; ++f) QF[f][i] = rhsi[f];
        }
        for (int i = 0; i < Nq; ++i) {
            for (int f = 0; f < nf; ++f) {
                double a = 0, b = 0;
                for (int k = d->lf_ptr[i]; k < d->lf_ptr[i + 1]; ++k) a += d->lf_val[k] * flux[f * sF + e * Nfq + d->lf_col[k]];
                for (int k = d->ph_ptr[i]; k < d->ph_ptr[i + 1]; ++k) b += d->ph_val[k] * QF[f][d->ph_col[k]];
                double r = -(a + d->volume_scale * b) / d->J[e];
                rhs[f * sQ + e * Nq + i] = r;
                if (compute_rhstest) rhstest += d->wq[i] * d->J[e] * VU[f * sQ + e * Nq + i] * r;
            }
        }
    }
    return rhstest;
}

static const double RK4A[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                               -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
static const double RK4B[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                               1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                               2277821191437.0 / 14882151754819.0};

/* nsteps LSRK45 steps in place (dg2D_euler_quad.jl:198-207); returns the last stage's rhstest */
double esdg_oracle_lsrk45(const oracle_desc* d, double* Q, double* resQ, double* rhs, double* work, double dt, int nsteps) {
    const int64_t n = (int64_t)d->nf * d->Nq * d->K;
    double rhstest = 0;
    for (int s = 0; s < nsteps; ++s)
        for (int k = 0; k < 5; ++k) {
            rhstest = esdg_oracle_rhs(d, Q, rhs, work, k == 4);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) {
                resQ[i] = RK4A[k] * resQ[i] + dt * rhs[i];
                Q[i] += RK4B[k] * resQ[i];
            }
        }
    return rhstest;
}
