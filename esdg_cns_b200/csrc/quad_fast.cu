// Host side of the tensor-product quad path: structure detection, buffers, launches.
#include "quad_fast.h"

#include "ctx_util.h"
#include "hex_fast.h"
#include "kernels_cnsq.cuh"
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
    p.sig = nullptr; p.sig_tr = nullptr; p.bc = nullptr; p.vlid = nullptr;
    p.partial_visc = p.partial_vrhs = nullptr;
    p.bctype = 1; p.viscous_dissp = 0;
    p.lam = p.mu = p.Pr = p.Re = p.inv_theta2 = 0.0;
    if (h->cns) {
        const CnsState& c = ctx->cns;
        p.sig = c.sig;
        p.sig_tr = c.sig_tr;
        p.bc = c.bc;
        p.vlid = c.vlid_dev;
        p.bctype = c.bctype;
        p.viscous_dissp = c.viscous_dissp;
        p.lam = -c.lambda;                                  // `let lambda=-lambda` (dg2D_CNS_cavity_optimized.jl:614)
        p.mu = c.mu; p.Pr = c.Pr; p.Re = c.Re;
        const double theta = 1.0 / 0.3 / 0.3 / 1.4 / 0.4;   // :189
        p.inv_theta2 = 2.0 / theta;
        std::memcpy(p.D, h->D, sizeof p.D);
        p.partial_visc = c.partial_visc;
        p.partial_vrhs = c.partial_vrhs;
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
    h->smem_rhs = (size_t)D::SMEM_DOUBLES_CNS * sizeof(double);
    h->smem_sigma = (size_t)CnsqDims<n>::SMEM_DOUBLES * sizeof(double);
    CK(cudaFuncSetAttribute(k_quad_stage_line<n, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_rhs));
    CK(cudaFuncSetAttribute(k_cnsq_sigma<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_sigma));
    return ESDG_OK;
}

template <int n>
int launch_traces_n(esdg_ctx* ctx, const double* Q) {
    QuadFast* h = ctx->quad;
    QuadParams p;
    fill_quad_params<n>(ctx, p);
    p.Q = Q;
    const int blocks = (int)((p.e_count + h->EB - 1) / h->EB);
    tick(ctx, h->cns ? "cns_traces" : "quad_traces");
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
    tick(ctx, h->cns ? "cns_rhs" : "quad_stage");
    if (h->cns)
        k_quad_stage_line<n, true><<<blocks, h->threads, h->smem_rhs, ctx->stream>>>(p);
    else
        k_quad_stage_line<n><<<blocks, h->threads, h->smem_stage, ctx->stream>>>(p);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

template <int n>
int launch_sigma_n(esdg_ctx* ctx, const double* Q, int want) {
    QuadFast* h = ctx->quad;
    QuadParams p;
    fill_quad_params<n>(ctx, p);
    p.Q = Q;
    p.Tcur = h->trace[h->prod];          // traces of the state the sigma pass is evaluated at (halo included)
    p.want_rhstest = want;
    const int blocks = (int)((p.e_count + h->EB - 1) / h->EB);
    tick(ctx, "cns_sigma");
    k_cnsq_sigma<n><<<blocks, h->threads, h->smem_sigma, ctx->stream>>>(p);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

// C = A B for small host matrices (column-major HostMat)
HostMat matmul(const HostMat& A, const HostMat& B) {
    HostMat C;
    C.rows = A.rows;
    C.cols = B.cols;
    C.a.assign((size_t)(A.rows * B.cols), 0.0);
    for (int64_t j = 0; j < B.cols; ++j)
        for (int64_t k = 0; k < A.cols; ++k) {
            const double b = B.at(k, j);
            for (int64_t i = 0; i < A.rows; ++i) C.a[(size_t)(i + j * A.rows)] += A.at(i, k) * b;
        }
    return C;
}

// the canonical tensor-product tables of a Gauss-collocated quad with N + 1 = n points per direction
struct QuadCanon {
    int n, Nq;
    double x[8], w[8], D[64], lm[8], lp[8];
    explicit QuadCanon(int n_) : n(n_), Nq(n_ * n_) {
        gauss_legendre(n, x, w);
        lagrange_tables(n, x, D, lm, lp);
    }
    // node k = i + n j, s = x[i] (fastest), r = x[j]; faces s=-1, r=+1, s=+1 (reversed), r=-1 (reversed)
    double wq(int k) const { return w[k % n] * w[k / n]; }
    void face_of(int fp, int& dd, int& side, int& m) const {   // direction (0: r, 1: s), side, index of the line it ends
        const int face = fp / n, q = fp % n;
        static const int DD[4] = {1, 0, 1, 0}, SD[4] = {0, 1, 1, 0};
        dd = DD[face];
        side = SD[face];
        m = face >= 2 ? n - 1 - q : q;
    }
    double ef(int fp, int k) const {
        int dd, side, m;
        face_of(fp, dd, side, m);
        const int i = k % n, j = k / n;
        const int along = dd == 0 ? j : i, across = dd == 0 ? i : j;
        if (across != m) return 0.0;
        return side ? lp[along] : lm[along];
    }
    double wf(int fp) const { int dd, side, m; face_of(fp, dd, side, m); return w[m]; }
    double deriv(int dir, int a, int b) const {     // (D (x) I) for r (acts on j), (I (x) D) for s
        const int ia = a % n, ja = a / n, ib = b % n, jb = b / n;
        if (dir == 0) return ia == ib ? D[ja * n + jb] : 0.0;
        return ja == jb ? D[ia * n + ib] : 0.0;
    }
    double qskew(int dir, int a, int b) const {
        if (a < Nq && b < Nq) {
            const int ia = a % n, ja = a / n, ib = b % n, jb = b / n;
            if (dir == 0) {
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
            const double v = .5 * ef(fp, k) * wf(fp) * (side ? 1.0 : -1.0);
            return upper ? v : -v;
        }
        return 0.0;
    }
};

int create_common(esdg_ctx* ctx, const QuadCanon& C, bool cns) {
    const esdg_desc& d = ctx->d;
    const int n = C.n;
    QuadFast* h = new QuadFast();
    ctx->quad = h;
    h->n = n;
    h->cns = cns;
    for (int a = 0; a < n; ++a) {
        h->lm[a] = C.lm[a];
        h->lp[a] = C.lp[a];
        h->w[a] = C.w[a];
        for (int b = 0; b < n; ++b) {
            h->T[a * n + b] = (C.w[a] * C.D[a * n + b] - C.w[b] * C.D[b * n + a]) / C.w[a];
            h->D[a * n + b] = C.D[a * n + b];
        }
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

}  // namespace

int cnsq_try_create(esdg_ctx* ctx) {
    const esdg_desc& d = ctx->d;
    const CnsState& c = ctx->cns;
    if (!c.on || d.elem != ESDG_QUAD || d.dim != 2 || !d.affine) return ESDG_OK;
    if (getenv("ESDG_FORCE_GENERIC") || getenv("ESDG_CNS_MODAL")) return ESDG_OK;
    if (!c.inflow.empty() || !c.outflow.empty()) return ESDG_OK;      // shock-tube BCs: table-driven kernels
    int n = (int)std::lround(std::sqrt((double)d.Nq));
    if (n * n != d.Nq || d.Ns != d.Nq || n < 2 || n > 5 || d.Nfaces != 4 || d.Nfq != 4 * n) return ESDG_OK;
    const int Nq = d.Nq;
    const HostMat *Vq = find(ctx->ops, "Vq"), *Pq = find(ctx->ops, "Pq"), *Vf = find(ctx->ops, "Vf"), *Dr = find(ctx->ops, "Dr"),
                  *Ds = find(ctx->ops, "Ds"), *Lf = find(ctx->ops, "Lf"), *Ph = find(ctx->ops, "Ph"), *wq = find(ctx->ops, "wq"),
                  *VhP = find(ctx->ops, "VhP");
    const HostMat* Qm[2] = {find(ctx->ops, "Qrh"), find(ctx->ops, "Qsh")};
    if (!Vq || !Pq || !Vf || !Dr || !Ds || !Lf || !Ph || !wq || !VhP || !Qm[0] || !Qm[1]) return ESDG_OK;
    if (Vq->rows != Nq || Vq->cols != Nq || Pq->rows != Nq || Pq->cols != Nq) return ESDG_OK;
    const QuadCanon C(n);
    const double TOLM = 1e-10;
    bool ok = rel_mismatch(*wq, [&](int i, int) { return C.wq(i); }) < TOLM;
    ok = ok && rel_mismatch(matmul(*Vq, *Pq), [&](int i, int j) { return i == j ? 1.0 : 0.0; }) < TOLM;
    ok = ok && rel_mismatch(matmul(*Pq, *Vq), [&](int i, int j) { return i == j ? 1.0 : 0.0; }) < TOLM;
    ok = ok && rel_mismatch(matmul(*Vf, *Pq), [&](int k, int i) { return C.ef(k, i); }) < TOLM;
    ok = ok && rel_mismatch(*VhP, [&](int r, int i) { return r < Nq ? (r == i ? 1.0 : 0.0) : C.ef(r - Nq, i); }) < TOLM;
    ok = ok && rel_mismatch(matmul(matmul(*Vq, *Dr), *Pq), [&](int a, int b) { return C.deriv(0, a, b); }) < TOLM;
    ok = ok && rel_mismatch(matmul(matmul(*Vq, *Ds), *Pq), [&](int a, int b) { return C.deriv(1, a, b); }) < TOLM;
    ok = ok && rel_mismatch(matmul(*Vq, *Lf), [&](int i, int k) { return C.ef(k, i) * C.wf(k) / C.wq(i); }) < TOLM;
    ok = ok && rel_mismatch(matmul(*Vq, *Ph), [&](int i, int j) {
             const double v = j < Nq ? (i == j ? 1.0 : 0.0) : C.ef(j - Nq, i);
             return 2.0 * v / C.wq(i) / d.volume_scale;
         }) < TOLM;
    for (int dir = 0; dir < 2 && ok; ++dir) ok = rel_mismatch(*Qm[dir], [&](int i, int j) { return C.qskew(dir, i, j); }) < TOLM;
    if (!ok) return ESDG_OK;
    int rc;
    if ((rc = create_common(ctx, C, true))) return rc;
    QuadFast* h = ctx->quad;
    std::vector<double> vq((size_t)Nq * Nq), pq((size_t)Nq * Nq);
    for (int r = 0; r < Nq; ++r)
        for (int q = 0; q < Nq; ++q) {
            vq[(size_t)r * Nq + q] = Vq->at(r, q);
            pq[(size_t)r * Nq + q] = Pq->at(r, q);
        }
    if ((rc = dev_upload(ctx, &h->Vq_dev, vq))) return rc;
    if ((rc = dev_upload(ctx, &h->Pq_dev, pq))) return rc;
    return ESDG_OK;
}

int cnsq_blocks(const esdg_ctx* ctx) { return (int)((ctx->d.K + ctx->quad->EB - 1) / ctx->quad->EB); }

int cnsq_sigma(esdg_ctx* ctx, const double* Q, int want) {
    switch (ctx->quad->n) {
        case 2: return launch_sigma_n<2>(ctx, Q, want);
        case 3: return launch_sigma_n<3>(ctx, Q, want);
        case 4: return launch_sigma_n<4>(ctx, Q, want);
        default: return launch_sigma_n<5>(ctx, Q, want);
    }
}

int cnsq_rhs(esdg_ctx* ctx, const double* Q, int mode, double a, double b, double dt, int want) {
    return quadfast_stage(ctx, Q, mode, a, b, dt, want, ESDG_PART_ALL);
}

int cnsq_basis(esdg_ctx* ctx, bool to_nodal, const double* src, double* dst) {
    QuadFast* h = ctx->quad;
    const int m = ctx->d.Nq;
    const long long nvec = (long long)ctx->NF * ctx->d.K;
    const int blocks = (int)std::min<long long>((nvec * m + 255) / 256, 148 * 16);
    tick(ctx, "basis_change");
    k_basis_change<<<blocks, 256, (size_t)m * m * sizeof(double), ctx->stream>>>(to_nodal ? h->Vq_dev : h->Pq_dev, m, src, dst, nvec);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

int cnsq_dopri_err(esdg_ctx* ctx, const double* const* K7, const double* c7, double err_tol, int nblocks, double* partial) {
    QuadFast* h = ctx->quad;
    const int m = ctx->d.Nq;
    DopriErrArgs a;
    for (int s = 0; s < 7; ++s) { a.K[s] = K7[s]; a.c[s] = c7[s]; }
    tick(ctx, "dopri_err");
    k_dopri_err_basis<<<nblocks, 256, (size_t)m * m * sizeof(double), ctx->stream>>>(h->Pq_dev, m, ctx->Q, a, err_tol,
                                                                                    (long long)ctx->NF * ctx->d.K, partial);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

int quadfast_try_create(esdg_ctx* ctx) {
    const esdg_desc& d = ctx->d;
    if (d.elem != ESDG_QUAD || d.dim != 2 || !d.collocated || !d.affine) return ESDG_OK;
    if (getenv("ESDG_FORCE_GENERIC")) return ESDG_OK;
    int n = (int)std::lround(std::sqrt((double)d.Nq));
    if (n * n != d.Nq || n < 2 || n > 5 || d.Nfaces != 4 || d.Nfq != 4 * n) return ESDG_OK;
    const int Nq = d.Nq;
    const HostMat* Ef = find(ctx->ops, "Ef");
    const HostMat* Ph = find(ctx->ops, "Ph");
    const HostMat* Lf = find(ctx->ops, "Lf");
    const HostMat* wq = find(ctx->ops, "wq");
    const HostMat* Qm[2] = {find(ctx->ops, "Qrh"), find(ctx->ops, "Qsh")};
    if (!Ef || !Ph || !Lf || !wq || !Qm[0] || !Qm[1]) return ESDG_OK;
    const QuadCanon C(n);
    const double TOLM = 1e-10;
    bool ok = rel_mismatch(*wq, [&](int i, int) { return C.wq(i); }) < TOLM;
    ok = ok && rel_mismatch(*Ef, [&](int k, int i) { return C.ef(k, i); }) < TOLM;
    for (int dir = 0; dir < 2 && ok; ++dir) ok = rel_mismatch(*Qm[dir], [&](int i, int j) { return C.qskew(dir, i, j); }) < TOLM;
    ok = ok && rel_mismatch(*Ph, [&](int i, int j) {
             const double v = j < Nq ? (i == j ? 1.0 : 0.0) : C.ef(j - Nq, i);
             return 2.0 * v / C.wq(i) / d.volume_scale;
         }) < TOLM;
    ok = ok && rel_mismatch(*Lf, [&](int i, int k) { return C.ef(k, i) * C.wf(k) / C.wq(i); }) < TOLM;
    if (!ok) return ESDG_OK;
    return create_common(ctx, C, false);
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
