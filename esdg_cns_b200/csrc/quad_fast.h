// Host side of the tensor-product quad path (kernels_quad_line.cuh).
#pragma once
#include "esdg_ctx.h"

constexpr int ESDG_QUAD_TRACE_PLANES = 7;   // == esdg::QUAD_TP

struct QuadFast {
    int n = 0;                               // N + 1
    double T[64], lm[8], lp[8], w[8];
    double* trace[2] = {nullptr, nullptr};   // 7 planes x trace_stride each
    double* logs[2] = {nullptr, nullptr};    // (log rho, log beta) at the volume nodes
    int prod = 0, cons = 0;                  // buffer produced last / consumed by the running stage
    size_t smem_stage = 0, smem_traces = 0;
    int threads = 0, EB = 1;
    // compressible Navier-Stokes in the Gauss-nodal basis (kernels_cnsq.cuh)
    bool cns = false;
    double D[64];                            // 1-D differentiation matrix at the Gauss nodes
    double *Vq_dev = nullptr, *Pq_dev = nullptr;   // Nq x Nq, row-major: caller's coefficients <-> nodal values
    size_t smem_sigma = 0, smem_rhs = 0;
};

// Decide whether the caller's operators are the Kronecker products the quad kernels assume; on success ctx->quad is set.
// A mismatch is not an error: ctx->quad stays null and the generic kernels are used.
int quadfast_try_create(esdg_ctx* ctx);
void quadfast_destroy(esdg_ctx* ctx);
int quadfast_traces(esdg_ctx* ctx, const double* Q);
// one rhs evaluation (mode 0) or fused LSRK stage (mode 1) over a part of the elements (ESDG_PART_*, hex_fast.h)
int quadfast_stage(esdg_ctx* ctx, const double* Q, int mode, double a, double b, double dt, int want_rhstest, int part);

// CNS contexts (ctx->cns.on, after finalize_cns): prove that in the Gauss-nodal basis the caller's operators are the
// Kronecker products of kernels_cnsq.cuh; on success ctx->quad is set with cns = true and the state lives in that basis.
int cnsq_try_create(esdg_ctx* ctx);
int cnsq_blocks(const esdg_ctx* ctx);                 // CTAs (= rhstest partials) of the sigma / rhs launches
int cnsq_sigma(esdg_ctx* ctx, const double* Q, int want_rhstest);
int cnsq_rhs(esdg_ctx* ctx, const double* Q, int mode, double a, double b, double dt, int want_rhstest);
// dst = Vq src (to_nodal) or Pq src, all fields of a state-sized array
int cnsq_basis(esdg_ctx* ctx, bool to_nodal, const double* src, double* dst);
// DOPRI error partials in the caller's basis (k_dopri_err_basis)
int cnsq_dopri_err(esdg_ctx* ctx, const double* const* K7, const double* c7, double err_tol, int nblocks, double* partial);
