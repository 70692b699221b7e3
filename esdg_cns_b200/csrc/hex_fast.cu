// Host side of the tensor-product hex path: structure detection, buffers, launches.
#include "hex_fast.h"

#include "ctx_util.h"
#include "tp_tables.h"
#include "kernels_hex_ws.cuh"

using namespace esdg;
using namespace esdg_host;

namespace {

template <int n>
void fill_hex_params(esdg_ctx* ctx, HexParams& p) {
    HexFast* h = ctx->hex;
    const esdg_desc& d = ctx->d;
    p.K = d.K;
    p.e_begin = 0;
    p.e_count = d.K;
    p.seg_len = 0;
    p.seg_stride = 0;
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
    p.geo_stride = ctx->geo_stride;
    p.geoG = ctx->geoG; p.geoJ = ctx->geoJ; p.geoN = ctx->geoN;
    p.geoL = h->geoL;
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
    using TD = TraceDims<n>;
    using LD = LineDims<n>;
    using WD = WsDims<n>;
    HexFast* h = ctx->hex;
    // ESDG_HEX_KERNEL = "line": the one-shot kernel everywhere; "ws": the warp-specialised persistent kernel everywhere;
    // default: persistent kernel for element ranges that fill the GPU, one-shot kernel for small ones
    const char* kv = getenv("ESDG_HEX_KERNEL");
    h->kernel_choice = kv && !strcmp(kv, "line") ? 1 : (kv && !strcmp(kv, "ws") ? 2 : 0);
    h->EB_tr = TD::EB;
    h->threads_tr = TD::THREADS;
    h->smem_traces = (size_t)TD::SMEM_DOUBLES * sizeof(double);
    h->EB = LD::EB;
    h->threads = LD::THREADS;
    h->smem_stage = (size_t)LD::SMEM_DOUBLES * sizeof(double);
    h->EB_ws = WD::EB;
    h->threads_ws = WD::THREADS;
    h->smem_ws = (size_t)WD::SMEM_DOUBLES * sizeof(double);
    CK(cudaFuncSetAttribute(k_hex_stage_line<n, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_stage));
    CK(cudaFuncSetAttribute(k_hex_stage_line<n, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_stage));
    CK(cudaFuncSetAttribute(k_hex_stage_ws<n, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_ws));
    CK(cudaFuncSetAttribute(k_hex_stage_ws<n, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_ws));
    CK(cudaFuncSetAttribute(k_hex_traces<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_traces));
    h->smem_curved = (size_t)LD::SMEM_DOUBLES_CURVED * sizeof(double);
    CK(cudaFuncSetAttribute(k_hex_stage_line_curved<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_curved));
    CK(cudaFuncSetAttribute(k_hex_traces<n, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_traces));
    CK(cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (const char* gv = getenv("ESDG_WS_GRID")) h->sms = std::max(1, atoi(gv));   // experiments
    h->comm_sms = 2;
    if (const char* cv = getenv("ESDG_COMM_SMS")) h->comm_sms = std::max(0, atoi(cv));
    return ESDG_OK;
}

long long host_batches(const HexParams& p, int EB) {   // == range_batches (hex_common.cuh)
    if (p.seg_len <= 0) return (p.e_count + EB - 1) / EB;
    return (p.e_count / p.seg_len) * ((p.seg_len + EB - 1) / EB);
}

template <int n>
int launch_traces_n(esdg_ctx* ctx, const double* Q, long long e_begin = 0, long long e_count = -1, long long seg_len = 0,
                    long long seg_stride = 0) {
    HexFast* h = ctx->hex;
    HexParams p;
    fill_hex_params<n>(ctx, p);
    p.Q = Q;
    if (e_count >= 0) { p.e_begin = e_begin; p.e_count = e_count; p.seg_len = seg_len; p.seg_stride = seg_stride; }
    const int blocks = (int)host_batches(p, h->EB_tr);
    tick(ctx, "hex_traces");
    if (h->curved)
        k_hex_traces<n, true><<<blocks, h->threads_tr, h->smem_traces, ctx->stream>>>(p);
    else
        k_hex_traces<n><<<blocks, h->threads_tr, h->smem_traces, ctx->stream>>>(p);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

template <int n>
int launch_stage_n(esdg_ctx* ctx, const HexParams& p, int reserve_sms = 0) {
    HexFast* h = ctx->hex;
    tick(ctx, "hex_stage");
    // the persistent kernel wants at least two batches per SM (its pipeline overlaps batch i with i +- 1)
    const long long nbatch_ws = host_batches(p, h->EB_ws);
    const bool ws = !h->curved && (h->kernel_choice == 2 || (h->kernel_choice == 0 && nbatch_ws >= 2LL * h->sms));
    if (h->curved) {
        const int blocks = (int)host_batches(p, h->EB);
        k_hex_stage_line_curved<n><<<blocks, h->threads, h->smem_curved, ctx->stream>>>(p);
    } else if (ws) {
        // the persistent CTAs take every register of their SM: while a halo exchange is in flight a few SMs stay free
        // for the NCCL send / receive kernels, which could otherwise not start before this launch has drained
        const int grid = (int)std::min<long long>(nbatch_ws, std::max(1, h->sms - reserve_sms));
        if (h->diag)
            k_hex_stage_ws<n, true><<<grid, h->threads_ws, h->smem_ws, ctx->stream>>>(p);
        else
            k_hex_stage_ws<n, false><<<grid, h->threads_ws, h->smem_ws, ctx->stream>>>(p);
    } else {
        const int blocks = (int)host_batches(p, h->EB);
        if (h->diag)
            k_hex_stage_line<n, true><<<blocks, h->threads, h->smem_stage, ctx->stream>>>(p);
        else
            k_hex_stage_line<n, false><<<blocks, h->threads, h->smem_stage, ctx->stream>>>(p);
    }
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

// curved meshes: the geometric factors each line needs -- direction d of line (d, m): G[3 c + d] for the rotated components
// c = d, d+1, d+2 at its n volume nodes and its two face points -- gathered line-major, [e][(n+2)*3][NL]
int build_line_geometry(esdg_ctx* ctx) {
    HexFast* h = ctx->hex;
    const esdg_desc& d = ctx->d;
    const int n = h->n, nfp = n * n, NL = 3 * nfp, Nq = d.Nq, NA = n + 2;
    static const char* g3[9] = {"rxJ", "sxJ", "txJ", "ryJ", "syJ", "tyJ", "rzJ", "szJ", "tzJ"};
    const HostMat* G[9];
    for (int k = 0; k < 9; ++k) {
        G[k] = find(ctx->geom, g3[k]);
        if (!G[k] || G[k]->rows != Nq + d.Nfq) return fail(ctx, ESDG_EINVAL, "esdg_finalize: curved mesh needs rxJ.. at every hybridized point");
    }
    std::vector<double> gl((size_t)d.K * 3 * NA * NL);
    for (int l = 0; l < NL; ++l) {
        const int dir = l / nfp, m = l % nfp, ip = m % n, jp = m / n;
        int base, stride;   // == face_line (hex_common.cuh)
        if (dir == 0) { base = n * jp + n * n * ip; stride = 1; }
        else if (dir == 1) { base = jp + n * n * ip; stride = n; }
        else { base = jp + n * ip; stride = n * n; }
        for (int a = 0; a < NA; ++a) {
            const int node = a < n ? base + a * stride : Nq + (2 * dir + (a - n)) * nfp + m;
            for (int c = 0; c < 3; ++c) {
                const int comp = (dir + c) % 3;
                const HostMat* g = G[3 * comp + dir];
                for (int64_t e = 0; e < d.K; ++e) gl[((size_t)e * 3 * NA + (a * 3 + c)) * NL + l] = g->at(node, e);
            }
        }
    }
    return dev_upload(ctx, &h->geoL, gl);
}

int set_attrs_dispatch(esdg_ctx* ctx) {
    switch (ctx->hex->n) {
        case 2: return set_attrs<2>(ctx);
        case 3: return set_attrs<3>(ctx);
        case 4: return set_attrs<4>(ctx);
        case 5: return set_attrs<5>(ctx);
    }
    return fail(ctx, ESDG_EINVAL, "hex path: unsupported degree");
}

}  // namespace

int hexfast_try_create(esdg_ctx* ctx) {
    const esdg_desc& d = ctx->d;
    if (d.elem != ESDG_HEX || d.dim != 3 || !d.collocated) return ESDG_OK;
    if (getenv("ESDG_FORCE_GENERIC")) return ESDG_OK;
    if (!d.affine && getenv("ESDG_CURVED_GENERIC")) return ESDG_OK;   // curved elements on the table-driven kernels (cross-check)
    int n = (int)std::lround(std::cbrt((double)d.Nq));
    if (n * n * n != d.Nq || n < 2 || n > 5 || d.Nfaces != 6 || d.Nfq != 6 * n * n) return ESDG_OK;
    const int Nq = d.Nq, nfp = n * n, Nfq = d.Nfq, Nh = Nq + Nfq;
    double x[8], w[8], D[64], lm[8], lp[8];
    gauss_legendre(n, x, w);
    lagrange_tables(n, x, D, lm, lp);
    const HostMat* Ef = find(ctx->ops, "Ef");
    const HostMat* Ph = find(ctx->ops, "Ph");
    const HostMat* Lf = find(ctx->ops, "Lf");
    const HostMat* wq = find(ctx->ops, "wq");
    const HostMat* Qm[3] = {find(ctx->ops, "Qrh"), find(ctx->ops, "Qsh"), find(ctx->ops, "Qth")};
    if (!Ef || !Ph || !Lf || !wq || !Qm[0] || !Qm[1] || !Qm[2]) return ESDG_OK;
    // canonical generators -------------------------------------------------------------
    auto coords = [&](int i, int* c) { c[0] = i % n; c[1] = (i / n) % n; c[2] = i / (n * n); };
    auto wq_c = [&](int i) { int c[3]; coords(i, c); return w[c[0]] * w[c[1]] * w[c[2]]; };
    // face point k -> (direction, side, line base coords)
    auto ef_c = [&](int k, int i) -> double {
        const int face = k / nfp, m = k % nfp, dd = face / 2, side = face % 2;
        const int ip = m % n, jp = m / n;
        int c[3];
        coords(i, c);
        int o1, o2;   // the two on-face coordinates of node i in the face's (jp, ip) convention
        if (dd == 0) { o1 = c[1]; o2 = c[2]; }
        else if (dd == 1) { o1 = c[0]; o2 = c[2]; }
        else { o1 = c[0]; o2 = c[1]; }
        if (o1 != jp || o2 != ip) return 0.0;
        return side ? lp[c[dd]] : lm[c[dd]];
    };
    auto wf_c = [&](int k) { const int m = k % nfp; return w[m % n] * w[m / n]; };
    auto q_c = [&](int dd, int i, int j) -> double {
        if (i < Nq && j < Nq) {
            int a[3], b[3];
            coords(i, a);
            coords(j, b);
            for (int q = 0; q < 3; ++q)
                if (q != dd && a[q] != b[q]) return 0.0;
            const int o1 = (dd + 1) % 3, o2 = (dd + 2) % 3;
            const double ww = w[a[o1]] * w[a[o2]];
            return ww * .5 * (w[a[dd]] * D[a[dd] * n + b[dd]] - w[b[dd]] * D[b[dd] * n + a[dd]]);
        }
        if (i < Nq && j >= Nq) {   // .5 * Ef' * B
            const int k = j - Nq, face = k / nfp;
            if (face / 2 != dd) return 0.0;
            const double nr = (face % 2) ? 1.0 : -1.0;
            return .5 * ef_c(k, i) * wf_c(k) * nr;
        }
        if (i >= Nq && j < Nq) {
            const int k = i - Nq, face = k / nfp;
            if (face / 2 != dd) return 0.0;
            const double nr = (face % 2) ? 1.0 : -1.0;
            return -.5 * ef_c(k, j) * wf_c(k) * nr;
        }
        return 0.0;
    };
    const double TOLM = 1e-10;
    bool ok = rel_mismatch(*wq, [&](int i, int) { return wq_c(i); }) < TOLM;
    ok = ok && rel_mismatch(*Ef, ef_c) < TOLM;
    for (int dd = 0; dd < 3 && ok; ++dd) ok = rel_mismatch(*Qm[dd], [&](int i, int j) { return q_c(dd, i, j); }) < TOLM;
    // volume_scale * Ph == 2 * diag(1/wq) [I; Ef]'
    ok = ok && rel_mismatch(*Ph, [&](int i, int j) {
             const double v = j < Nq ? (i == j ? 1.0 : 0.0) : ef_c(j - Nq, i);
             return 2.0 * v / wq_c(i) / d.volume_scale;
         }) < TOLM;
    ok = ok && rel_mismatch(*Lf, [&](int i, int k) { return ef_c(k, i) * wf_c(k) / wq_c(i); }) < TOLM;
    if (!ok) return ESDG_OK;   // e.g. the upstream s-fastest node ordering: generic kernels handle it

    HexFast* h = new HexFast();
    ctx->hex = h;
    h->n = n;
    h->curved = !d.affine;
    for (int a = 0; a < n; ++a) {
        h->lm[a] = lm[a];
        h->lp[a] = lp[a];
        h->w[a] = w[a];
        for (int b = 0; b < n; ++b) h->T[a * n + b] = (w[a] * D[a * n + b] - w[b] * D[b * n + a]) / w[a];
    }
    (void)Nh;
    int rc;
    if ((rc = set_attrs_dispatch(ctx))) return rc;
    if (h->curved && (rc = build_line_geometry(ctx))) return rc;
    for (int q = 0; q < 2; ++q) {
        if ((rc = dev_alloc(ctx, &h->trace[q], (size_t)HEX_TP * ctx->trace_stride))) return rc;
        if ((rc = dev_alloc(ctx, &h->logs[q], (size_t)2 * d.Nq * d.K))) return rc;
        CK(cudaMemsetAsync(h->trace[q], 0, (size_t)HEX_TP * ctx->trace_stride * sizeof(double), ctx->stream));
    }
    return ESDG_OK;
}

void hexfast_destroy(esdg_ctx* ctx) {
    delete ctx->hex;
    ctx->hex = nullptr;
}

int hexfast_traces(esdg_ctx* ctx, const double* Q) {
    HexFast* h = ctx->hex;
    h->prod ^= 1;            // write a fresh buffer; consumers switch to it at their next stage
    h->cons = h->prod ^ 1;
    switch (h->n) {
        case 2: return launch_traces_n<2>(ctx, Q);
        case 3: return launch_traces_n<3>(ctx, Q);
        case 4: return launch_traces_n<4>(ctx, Q);
        case 5: return launch_traces_n<5>(ctx, Q);
    }
    return fail(ctx, ESDG_EINVAL, "hex path: unsupported degree");
}

int hexfast_traces_range(esdg_ctx* ctx, const double* Q, long long e_begin, long long e_count, long long seg_len, long long seg_stride) {
    switch (ctx->hex->n) {
        case 2: return launch_traces_n<2>(ctx, Q, e_begin, e_count, seg_len, seg_stride);
        case 3: return launch_traces_n<3>(ctx, Q, e_begin, e_count, seg_len, seg_stride);
        case 4: return launch_traces_n<4>(ctx, Q, e_begin, e_count, seg_len, seg_stride);
        case 5: return launch_traces_n<5>(ctx, Q, e_begin, e_count, seg_len, seg_stride);
    }
    return fail(ctx, ESDG_EINVAL, "hex path: unsupported degree");
}

int hexfast_stage_range(esdg_ctx* ctx, const double* Q, int mode, double a, double b, double dt, long long e_begin, long long e_count,
                        long long seg_len, long long seg_stride) {
    HexFast* h = ctx->hex;
    HexParams p;
    switch (h->n) {
        case 2: fill_hex_params<2>(ctx, p); break;
        case 3: fill_hex_params<3>(ctx, p); break;
        case 4: fill_hex_params<4>(ctx, p); break;
        default: fill_hex_params<5>(ctx, p); break;
    }
    p.Q = Q;
    p.e_begin = e_begin;
    p.e_count = e_count;
    p.seg_len = seg_len;
    p.seg_stride = seg_stride;
    p.mode = mode;
    p.rk_a = a; p.rk_b = b; p.dt = dt;
    switch (h->n) {
        case 2: return launch_stage_n<2>(ctx, p);
        case 3: return launch_stage_n<3>(ctx, p);
        case 4: return launch_stage_n<4>(ctx, p);
        default: return launch_stage_n<5>(ctx, p);
    }
}

int hexfast_stage(esdg_ctx* ctx, const double* Q, int mode, double a, double b, double dt, int want_rhstest, int part) {
    HexFast* h = ctx->hex;
    const esdg_desc& d = ctx->d;
    if (part != ESDG_PART_INTERIOR) {   // first (or only) part of a stage: consume what was produced last
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
    if (want_rhstest && part != ESDG_PART_INTERIOR)   // one slot per element; batches write the slot of their first element
        CK(cudaMemsetAsync(ctx->partials, 0, (size_t)d.K * sizeof(double), ctx->stream));
    for (int r = 0; r < nr; ++r) {
        if (count[r] <= 0) continue;
        HexParams p;
        switch (h->n) {
            case 2: fill_hex_params<2>(ctx, p); break;
            case 3: fill_hex_params<3>(ctx, p); break;
            case 4: fill_hex_params<4>(ctx, p); break;
            default: fill_hex_params<5>(ctx, p); break;
        }
        p.Q = Q;
        p.e_begin = begin[r];
        p.e_count = count[r];
        p.mode = mode;
        p.rk_a = a; p.rk_b = b; p.dt = dt;
        p.want_rhstest = want_rhstest;
        int rc;
        const int reserve = (part == ESDG_PART_INTERIOR) ? h->comm_sms : 0;
        switch (h->n) {
            case 2: rc = launch_stage_n<2>(ctx, p, reserve); break;
            case 3: rc = launch_stage_n<3>(ctx, p, reserve); break;
            case 4: rc = launch_stage_n<4>(ctx, p, reserve); break;
            default: rc = launch_stage_n<5>(ctx, p, reserve); break;
        }
        if (rc) return rc;
    }
    return ESDG_OK;
}

// ---- diagnostic entry point: the kernels' own log / exp on caller-supplied arguments -------------
namespace {
__global__ void k_debug_log_exp(const double* x, long long n, double* lg, double* ex) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        lg[i] = x[i] > 0.0 ? log_pos(x[i]) : 0.0;
        ex[i] = exp_mod(x[i]);
    }
}
}  // namespace

extern "C" int esdg_debug_log_exp(int device, const double* x, int64_t n, double* log_out, double* exp_out) {
    if (!x || !log_out || !exp_out || n <= 0) return ESDG_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return ESDG_ECUDA;
    double *dx = nullptr, *dl = nullptr, *de = nullptr;
    if (cudaMalloc(&dx, n * 8) != cudaSuccess || cudaMalloc(&dl, n * 8) != cudaSuccess || cudaMalloc(&de, n * 8) != cudaSuccess)
        return ESDG_ENOMEM;
    cudaMemcpy(dx, x, n * 8, cudaMemcpyHostToDevice);
    k_debug_log_exp<<<(unsigned)((n + 255) / 256), 256>>>(dx, n, dl, de);
    cudaMemcpy(log_out, dl, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(exp_out, de, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dl); cudaFree(de);
    return cudaGetLastError() == cudaSuccess ? ESDG_OK : ESDG_ECUDA;
}


// ---- diagnostic entry point: the line kernel's own two-point flux (pair_flux_v) on caller-supplied pairs -----
namespace {
__global__ void k_debug_pair_flux(const double* L, const double* R, long long n, double gamma, double* F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    NodeV<1> A, B;
    const double* S[2] = {L, R};
    NodeV<1>* N2[2] = {&A, &B};
    for (int s = 0; s < 2; ++s) {
        N2[s]->rho[0] = S[s][i];
        for (int c = 0; c < 3; ++c) N2[s]->u[c][0] = .5 * S[s][(1 + c) * n + i];   // the kernels keep HALF velocities
        N2[s]->beta[0] = S[s][4 * n + i];
        N2[s]->lr[0] = S[s][5 * n + i];
        N2[s]->lb[0] = S[s][6 * n + i];
    }
    const double c_igm = 1.0 / (2.0 * (gamma - 1.0));
    VD<1> Fv[5];
    {   // axis-aligned form: flux / g0 along the first component
        const VD<1> g[3] = {vset<1>(1.0), vset<1>(0.0), vset<1>(0.0)};
        pair_flux_v<1, true>(A, B, g, c_igm, Fv);
        for (int f = 0; f < 5; ++f) F[(0 * 5 + f) * n + i] = Fv[f][0];
    }
    for (int d = 0; d < 3; ++d) {   // general form contracted with e_d
        const VD<1> g[3] = {vset<1>(d == 0 ? 1.0 : 0.0), vset<1>(d == 1 ? 1.0 : 0.0), vset<1>(d == 2 ? 1.0 : 0.0)};
        pair_flux_v<1, false>(A, B, g, c_igm, Fv);
        for (int f = 0; f < 5; ++f) F[((1 + d) * 5 + f) * n + i] = Fv[f][0];
    }
}
}  // namespace

extern "C" int esdg_debug_pair_flux(int device, const double* UL, const double* UR, int64_t n, double gamma, double* F_out) {
    if (!UL || !UR || !F_out || n <= 0) return ESDG_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return ESDG_ECUDA;
    double *dl = nullptr, *dr = nullptr, *df = nullptr;
    if (cudaMalloc(&dl, n * 56) != cudaSuccess || cudaMalloc(&dr, n * 56) != cudaSuccess || cudaMalloc(&df, n * 160) != cudaSuccess)
        return ESDG_ENOMEM;
    cudaMemcpy(dl, UL, n * 56, cudaMemcpyHostToDevice);
    cudaMemcpy(dr, UR, n * 56, cudaMemcpyHostToDevice);
    k_debug_pair_flux<<<(unsigned)((n + 127) / 128), 128>>>(dl, dr, n, gamma, df);
    cudaMemcpy(F_out, df, n * 160, cudaMemcpyDeviceToHost);
    cudaFree(dl); cudaFree(dr); cudaFree(df);
    return cudaGetLastError() == cudaSuccess ? ESDG_OK : ESDG_ECUDA;
}

#ifdef ESDG_PHASE_TIMING
// debug build: read and clear the per-activity cycle counters of k_hex_stage_ws
extern "C" int esdg_debug_phase_cycles(unsigned long long* out16) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof z) != cudaSuccess) return ESDG_ECUDA;
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z);
    return ESDG_OK;
}
#endif
