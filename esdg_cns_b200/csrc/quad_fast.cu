// Host side of the tensor-product quad path: structure detection, buffers, launches.
#include "quad_fast.h"

#include "ctx_util.h"
#include "hex_fast.h"
#include "kernels_quad_line.cuh"
#include "tp_tables.h"

using namespace esdg;
using namespace esdg_host;

namespace {

template <int n>
void fill_quad_params(esdg_ctx* ctx, QuadParams& p) {
    QuadFast* h = ctx->quad;
    const esdg_desc& d = ctx->d;
    p.K = d.K;
    p.e_begin = 0;
    p.e_count = d.K;
    p.q_stride = (long long)d.Nq * d.K;
    p.t_stride = ctx->trace_stride;
    p.Q = ctx->Q;
    p.Qout = ctx->Q;
    p.res = ctx->res;
    p.rhs = ctx->rhs_target ? ctx->rhs_target : ctx->rhs;
    p.Lcur = h->logs[h->cons];
    p.Lnext = h->logs[h->prod];
    p.Tcur = h->trace[h->cons];
    p.Tnext = h->trace[h->prod];
    p.mapP = ctx->mapP;
    p.geo = ctx->geo;
    p.gamma = d.gamma;
    p.lf_scale = d.lf_scale;
    p.lg1 = std::log(d.gamma - 1.0);
    p.l2g1 = std::log(2.0 * (d.gamma - 1.0));
    p.rk_a = p.rk_b = p.dt = 0.0;
    p.mode = 0;
    p.want_rhstest = 0;
    p.partial = ctx->partials;
    p.bad = ctx->bad_flag;
    std::memcpy(p.T, h->T, sizeof p.T);
    std::memcpy(p.lm, h->lm, sizeof p.lm);
    std::memcpy(p.lp, h->lp, sizeof p.lp);
    std::memcpy(p.w, h->w, sizeof p.w);
    for (int a = 0; a < 8; ++a) {
        p.lmw[a] = a < n ? h->lm[a] / h->w[a] : 0.0;
        p.lpw[a] = a < n ? h->lp[a] / h->w[a] : 0.0;
    }
}

template <int n>
int set_attrs(esdg_ctx* ctx) {
    using D = QuadDims<n>;
    QuadFast* h = ctx->quad;
    h->EB = D::EB;
    h->threads = D::THREADS;
    h->smem_stage = (size_t)D::SMEM_DOUBLES * sizeof(double);
    h->smem_traces = (size_t)(4 * D::EN + D::EB * QUAD_GEO + 2 * n) * sizeof(double);
    CK(cudaFuncSetAttribute(k_quad_stage_line<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_stage));
    CK(cudaFuncSetAttribute(k_quad_traces<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_traces));
    return ESDG_OK;
}

template <int n>
int launch_traces_n(esdg_ctx* ctx, const double* Q) {
    QuadFast* h = ctx->quad;
    QuadParams p;
    fill_quad_params<n>(ctx, p);
    p.Q = Q;
    const int blocks = (int)((p.e_count + h->EB - 1) / h->EB);
    tick(ctx, "quad_traces");
    k_quad_traces<n><<<blocks, h->threads, h->smem_traces, ctx->stream>>>(p);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

template <int n>
int launch_stage_n(esdg_ctx* ctx, QuadParams& p, const double* Q, int mode, double a, double b, double dt, int want, long long e_begin,
                   long long e_count) {
    QuadFast* h = ctx->quad;
    fill_quad_params<n>(ctx, p);
    p.Q = Q;
    p.e_begin = e_begin;
    p.e_count = e_count;
    p.mode = mode;
    p.rk_a = a; p.rk_b = b; p.dt = dt;
    p.want_rhstest = want;
    const int blocks = (int)((e_count + h->EB - 1) / h->EB);
    tick(ctx, "quad_stage");
    k_quad_stage_line<n><<<blocks, h->threads, h->smem_stage, ctx->stream>>>(p);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

}  // namespace

int quadfast_try_create(esdg_ctx* ctx) {
    const esdg_desc& d = ctx->d;
    if (d.elem != ESDG_QUAD || d.dim != 2 || !d.collocated || !d.affine) return ESDG_OK;
    if (getenv("ESDG_FORCE_GENERIC")) return ESDG_OK;
    int n = (int)std::lround(std::sqrt((double)d.Nq));
    if (n * n != d.Nq || n < 2 || n > 5 || d.Nfaces != 4 || d.Nfq != 4 * n) return ESDG_OK;
    const int Nq = d.Nq;
    double x[8], w[8], D[64], lm[8], lp[8];
    gauss_legendre(n, x, w);
    lagrange_tables(n, x, D, lm, lp);
    const HostMat* Ef = find(ctx->ops, "Ef");
    const HostMat* Ph = find(ctx->ops, "Ph");
    const HostMat* Lf = find(ctx->ops, "Lf");
    const HostMat* wq = find(ctx->ops, "wq");
    const HostMat* Qm[2] = {find(ctx->ops, "Qrh"), find(ctx->ops, "Qsh")};
    if (!Ef || !Ph || !Lf || !wq || !Qm[0] || !Qm[1]) return ESDG_OK;
    // canonical generators: node k = i + n j, s = x[i] (fastest), r = x[j]; faces s=-1, r=+1, s=+1 (reversed), r=-1 (reversed)
    auto wq_c = [&](int k) { return w[k % n] * w[k / n]; };
    auto face_of = [&](int fp, int& dd, int& side, int& m) {   // direction (0: r, 1: s), side, index of the line it ends
        const int face = fp / n, q = fp % n;
        static const int DD[4] = {1, 0, 1, 0}, SD[4] = {0, 1, 1, 0};
        dd = DD[face];
        side = SD[face];
        m = face >= 2 ? n - 1 - q : q;
    };
    auto ef_c = [&](int fp, int k) -> double {
        int dd, side, m;
        face_of(fp, dd, side, m);
        const int i = k % n, j = k / n;
        const int along = dd == 0 ? j : i, across = dd == 0 ? i : j;
        if (across != m) return 0.0;
        return side ? lp[along] : lm[along];
    };
    auto wf_c = [&](int fp) { int dd, side, m; face_of(fp, dd, side, m); return w[m]; };
    auto q_c = [&](int dir, int a, int b) -> double {
        if (a < Nq && b < Nq) {
            const int ia = a % n, ja = a / n, ib = b % n, jb = b / n;
            if (dir == 0) {   // d/dr acts on j
                if (ia != ib) return 0.0;
                return w[ia] * .5 * (w[ja] * D[ja * n + jb] - w[jb] * D[jb * n + ja]);
            }
            if (ja != jb) return 0.0;
            return w[ja] * .5 * (w[ia] * D[ia * n + ib] - w[ib] * D[ib * n + ia]);
        }
        if (a < Nq || b < Nq) {   // +-.5 Ef' B
            const bool upper = a < Nq;
            const int fp = (upper ? b : a) - Nq, k = upper ? a : b;
            int dd, side, m;
            face_of(fp, dd, side, m);
            if (dd != dir) return 0.0;
            const double v = .5 * ef_c(fp, k) * wf_c(fp) * (side ? 1.0 : -1.0);
            return upper ? v : -v;
        }
        return 0.0;
    };
    const double TOLM = 1e-10;
    bool ok = rel_mismatch(*wq, [&](int i, int) { return wq_c(i); }) < TOLM;
    ok = ok && rel_mismatch(*Ef, ef_c) < TOLM;
    for (int dir = 0; dir < 2 && ok; ++dir) ok = rel_mismatch(*Qm[dir], [&](int i, int j) { return q_c(dir, i, j); }) < TOLM;
    ok = ok && rel_mismatch(*Ph, [&](int i, int j) {
             const double v = j < Nq ? (i == j ? 1.0 : 0.0) : ef_c(j - Nq, i);
             return 2.0 * v / wq_c(i) / d.volume_scale;
         }) < TOLM;
    ok = ok && rel_mismatch(*Lf, [&](int i, int k) { return ef_c(k, i) * wf_c(k) / wq_c(i); }) < TOLM;
    if (!ok) return ESDG_OK;

    QuadFast* h = new QuadFast();
    ctx->quad = h;
    h->n = n;
    for (int a = 0; a < n; ++a) {
        h->lm[a] = lm[a];
        h->lp[a] = lp[a];
        h->w[a] = w[a];
        for (int b = 0; b < n; ++b) h->T[a * n + b] = (w[a] * D[a * n + b] - w[b] * D[b * n + a]) / w[a];
    }
    int rc;
    switch (n) {
        case 2: rc = set_attrs<2>(ctx); break;
        case 3: rc = set_attrs<3>(ctx); break;
        case 4: rc = set_attrs<4>(ctx); break;
        default: rc = set_attrs<5>(ctx); break;
    }
    if (rc) return rc;
    for (int q = 0; q < 2; ++q) {
        if ((rc = dev_alloc(ctx, &h->trace[q], (size_t)QUAD_TP * ctx->trace_stride))) return rc;
        if ((rc = dev_alloc(ctx, &h->logs[q], (size_t)2 * d.Nq * d.K))) return rc;
        CK(cudaMemsetAsync(h->trace[q], 0, (size_t)QUAD_TP * ctx->trace_stride * sizeof(double), ctx->stream));
    }
    return ESDG_OK;
}

void quadfast_destroy(esdg_ctx* ctx) {
    delete ctx->quad;
    ctx->quad = nullptr;
}

int quadfast_traces(esdg_ctx* ctx, const double* Q) {
    QuadFast* h = ctx->quad;
    h->prod ^= 1;
    h->cons = h->prod ^ 1;
    switch (h->n) {
        case 2: return launch_traces_n<2>(ctx, Q);
        case 3: return launch_traces_n<3>(ctx, Q);
        case 4: return launch_traces_n<4>(ctx, Q);
        default: return launch_traces_n<5>(ctx, Q);
    }
}

int quadfast_stage(esdg_ctx* ctx, const double* Q, int mode, double a, double b, double dt, int want_rhstest, int part) {
    QuadFast* h = ctx->quad;
    const esdg_desc& d = ctx->d;
    if (part != ESDG_PART_INTERIOR) {
        h->cons = h->prod;
        if (mode == 1) h->prod = h->cons ^ 1;
    }
    long long begin[2] = {0, 0}, count[2] = {d.K, 0};
    int nr = 1;
    const long long lo = ctx->bnd_lo, hi = ctx->bnd_hi;
    if (part == ESDG_PART_BOUNDARY) {
        begin[0] = 0; count[0] = lo;
        begin[1] = d.K - hi; count[1] = hi;
        nr = 2;
    } else if (part == ESDG_PART_INTERIOR) {
        begin[0] = lo; count[0] = d.K - lo - hi;
    }
    if (want_rhstest && part != ESDG_PART_INTERIOR) CK(cudaMemsetAsync(ctx->partials, 0, (size_t)d.K * sizeof(double), ctx->stream));
    for (int r = 0; r < nr; ++r) {
        if (count[r] <= 0) continue;
        QuadParams p;
        int rc;
        switch (h->n) {
            case 2: rc = launch_stage_n<2>(ctx, p, Q, mode, a, b, dt, want_rhstest, begin[r], count[r]); break;
            case 3: rc = launch_stage_n<3>(ctx, p, Q, mode, a, b, dt, want_rhstest, begin[r], count[r]); break;
            case 4: rc = launch_stage_n<4>(ctx, p, Q, mode, a, b, dt, want_rhstest, begin[r], count[r]); break;
            default: rc = launch_stage_n<5>(ctx, p, Q, mode, a, b, dt, want_rhstest, begin[r], count[r]); break;
        }
        if (rc) return rc;
    }
    return ESDG_OK;
}
