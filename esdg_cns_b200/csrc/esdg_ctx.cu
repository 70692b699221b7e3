// libesdg_b200.so -- C ABI implementation (see include/esdg_b200.h).
#include "ctx_util.h"
#include "hex_fast.h"
#include "quad_fast.h"
#include <algorithm>
#include <functional>

#include "kernels_generic.cuh"
#include "kernels_cns.cuh"

using namespace esdg;

#ifndef ESDG_SRC_HASH
#define ESDG_SRC_HASH "unknown"
#endif
// the source hash (Makefile) lets the loader refuse a stale prebuilt library: esdg_cns_b200/_lib.py recomputes it
#define ESDG_VERSION "esdg_b200 0.2 (sm_100a) src=" ESDG_SRC_HASH
#ifndef ESDG_GENERIC_COLL_REGS
#define ESDG_GENERIC_COLL_REGS 96   // register cap of the generic rhs kernel on collocated 2-D elements (kernels_generic.cuh)
#endif

using namespace esdg_host;

namespace {

void fill_params(esdg_ctx* ctx, GenericParams& p) {
    const esdg_desc& d = ctx->d;
    p.Ns = d.Ns; p.Nq = d.Nq; p.Nfq = d.Nfq; p.Nh = ctx->Nh; p.Nfaces = d.Nfaces; p.Nfp = ctx->Nfp;
    p.collocated = d.collocated;
    p.EB = ctx->EB;
    p.K = d.K;
    p.trace_stride = ctx->trace_stride;
    p.gamma = d.gamma; p.lf_scale = d.lf_scale; p.volume_scale = d.volume_scale;
    p.lg1 = std::log(d.gamma - 1.0); p.igm1x2 = 1.0 / (2.0 * (d.gamma - 1.0));
    p.Vq = ctx->Vq.view();
    p.Proj = ctx->Proj.view();
    p.Ph = ctx->Ph.view();
    p.Lf = ctx->Lf.view();
    p.Qp = PairEll{ctx->Nh, ctx->qp_width, ctx->qp_cnt, ctx->qp_col, ctx->qp_c};
    p.sym_rounds = (ctx->sym_rounds > 0 && ctx->threads >= ctx->EB * ctx->Nh) ? ctx->sym_rounds : 0;
    p.sym_dst = ctx->sym_dst;
    p.sym_c = ctx->sym_c;
    p.dVq = ctx->dVq; p.dProj = ctx->dProj; p.dProjF = ctx->dProjF; p.dPh = ctx->dPh; p.dLf = ctx->dLf;
    p.wq = ctx->wq;
    p.geo = ctx->geo;
    p.geo_stride = ctx->geo_stride;
    p.curved = ctx->d.affine == 0;
    p.geoG = ctx->geoG; p.geoJ = ctx->geoJ; p.geoN = ctx->geoN;
    p.mapP = ctx->mapP;
    p.Q = ctx->Q;
    p.q_stride = (long long)d.Ns * d.K;
    p.traces = ctx->traces;
    p.rhs = ctx->rhs_target ? ctx->rhs_target : ctx->rhs;
    p.Qout = ctx->Q;
    p.res = ctx->res;
    p.rk_a = 0; p.rk_b = 0; p.dt = 0;
    p.mode = 0;
    p.want_rhstest = 0;
    p.rhstest_partial = ctx->partials;
    p.bad_flag = ctx->bad_flag;
    p.stage_tables = ctx->stage_tables;
    p.table_off = ctx->table_off;
}

int grid_blocks(const esdg_ctx* ctx) { return (int)((ctx->d.K + ctx->EB - 1) / ctx->EB); }

int launch_traces(esdg_ctx* ctx, const double* Qsrc) {
    GenericParams p;
    fill_params(ctx, p);
    p.Q = Qsrc;
    tick(ctx, "generic_traces");
    if (ctx->d.dim == 2)
        k_generic_traces<2><<<grid_blocks(ctx), ctx->threads, ctx->smem_generic_traces, ctx->stream>>>(p);
    else
        k_generic_traces<3><<<grid_blocks(ctx), ctx->threads, ctx->smem_generic_traces, ctx->stream>>>(p);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

int launch_rhs(esdg_ctx* ctx, const double* Qsrc, int mode, double a, double b, double dt, int want_rhstest) {
    GenericParams p;
    fill_params(ctx, p);
    p.Q = Qsrc;
    p.mode = mode;
    p.rk_a = a; p.rk_b = b; p.dt = dt;
    p.want_rhstest = want_rhstest;
    tick(ctx, "generic_rhs");
    if (ctx->d.dim == 2 && !ctx->d.collocated)
        k_generic_rhs<2, 64><<<grid_blocks(ctx), ctx->threads, ctx->smem_bytes, ctx->stream>>>(p);
    else if (ctx->d.dim == 2)
        k_generic_rhs<2, ESDG_GENERIC_COLL_REGS><<<grid_blocks(ctx), ctx->threads, ctx->smem_bytes, ctx->stream>>>(p);
    else
        k_generic_rhs<3, 128><<<grid_blocks(ctx), ctx->threads, ctx->smem_bytes, ctx->stream>>>(p);
    tock(ctx);
    CK(cudaGetLastError());
    if (want_rhstest) {
        tick(ctx, "reduce_partials");
        k_reduce_partials<<<1, 256, 0, ctx->stream>>>(ctx->partials, grid_blocks(ctx), ctx->rhstest_dev);
        tock(ctx);
        CK(cudaGetLastError());
    }
    return ESDG_OK;
}

// ---- compressible Navier-Stokes path (kernels_cns.cuh) -----------------------------------------
void fill_cns(esdg_ctx* ctx, CnsParams& P, const double* Qsrc) {
    fill_params(ctx, P.g);
    P.g.Q = Qsrc;
    const CnsState& c = ctx->cns;
    P.Pq = c.Pq.view(); P.Vf = c.Vf.view(); P.Dr = c.Dr.view(); P.Ds = c.Ds.view();
    P.dPq = c.dPq; P.dVf = c.dVf; P.dDr = c.dDr; P.dDs = c.dDs;
    P.bc = c.bc;
    P.vlid = c.vlid_dev;
    {
        const double g = ctx->d.gamma, rho = c.inflow_state[0], u = c.inflow_state[1], v = c.inflow_state[2], pr = c.inflow_state[3];
        const double beta = rho / (2.0 * pr);
        P.in_prim[0] = rho; P.in_prim[1] = u; P.in_prim[2] = v; P.in_prim[3] = beta;
        P.in_prim[4] = std::log(rho); P.in_prim[5] = std::log(beta);
        // v_ufun (euler_variables.jl:79-92)
        const double rhoe = pr / (g - 1.0), E = rhoe + .5 * rho * (u * u + v * v);
        const double sU = std::log((g - 1.0) * rhoe / std::pow(rho, g));
        P.in_v[0] = (-E + rhoe * (g + 1.0 - sU)) / rhoe;
        P.in_v[1] = rho * u / rhoe;
        P.in_v[2] = rho * v / rhoe;
        P.in_v[3] = -rho / rhoe;
    }
    P.bctype = c.bctype;
    P.viscous_dissp = c.viscous_dissp;
    P.lam = -c.lambda;                                  // `let lambda=-lambda` (dg2D_CNS_cavity_optimized.jl:614)
    P.mu = c.mu; P.Pr = c.Pr; P.Re = c.Re;
    const double theta = 1.0 / 0.3 / 0.3 / 1.4 / 0.4;   // :189 (`1.0/0.3^2/1.4/0.4`, hard-coded in the reference)
    P.inv_theta2 = 2.0 / theta;
    P.sig = c.sig;
    P.sig_stride = (long long)ctx->d.Ns * ctx->d.K;
    P.sig_tr = c.sig_tr;
    P.partial_visc = c.partial_visc;
    P.partial_vrhs = c.partial_vrhs;
}

int cns_traces(esdg_ctx* ctx, const double* Qsrc) {
    if (ctx->quad) return quadfast_traces(ctx, Qsrc);      // Gauss-nodal tensor-product path (kernels_cnsq.cuh)
    CnsParams P;
    fill_cns(ctx, P, Qsrc);
    tick(ctx, "cns_traces");
    k_cns_traces<<<grid_blocks(ctx), ctx->threads, ctx->smem_cns_traces, ctx->stream>>>(P);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

// sigma pass (the V traces of Qsrc, halo included, must be valid): sigma (modal) + sigma traces of the local elements
int cns_sigma_pass(esdg_ctx* ctx, const double* Qsrc, int want) {
    if (ctx->quad) return cnsq_sigma(ctx, Qsrc, want);
    CnsParams P;
    fill_cns(ctx, P, Qsrc);
    P.g.want_rhstest = want;
    tick(ctx, "cns_sigma");
    k_cns_sigma<<<grid_blocks(ctx), ctx->threads, ctx->smem_cns_sigma, ctx->stream>>>(P);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

// rhs pass (V traces and sigma traces, halo included, must be valid)
int cns_rhs_pass(esdg_ctx* ctx, const double* Qsrc, int mode, double a, double b, double dt, int want) {
    if (ctx->quad) {
        int rc = cnsq_rhs(ctx, Qsrc, mode, a, b, dt, want);
        if (rc) return rc;
        if (want) {
            tick(ctx, "reduce_partials");
            k_reduce_partials_cns<<<1, 256, 0, ctx->stream>>>(ctx->partials, ctx->cns.partial_vrhs, ctx->cns.partial_visc, cnsq_blocks(ctx),
                                                             ctx->rhstest_dev);
            tock(ctx);
            CK(cudaGetLastError());
        }
        return ESDG_OK;
    }
    CnsParams P;
    fill_cns(ctx, P, Qsrc);
    P.g.mode = mode;
    P.g.rk_a = a; P.g.rk_b = b; P.g.dt = dt;
    P.g.want_rhstest = want;
    tick(ctx, "cns_rhs");
    k_cns_rhs<<<grid_blocks(ctx), ctx->threads, ctx->smem_cns_rhs, ctx->stream>>>(P);
    tock(ctx);
    CK(cudaGetLastError());
    if (want) {
        tick(ctx, "reduce_partials");
        k_reduce_partials_cns<<<1, 256, 0, ctx->stream>>>(ctx->partials, ctx->cns.partial_vrhs, ctx->cns.partial_visc,
                                                         grid_blocks(ctx), ctx->rhstest_dev);
        tock(ctx);
        CK(cudaGetLastError());
    }
    return ESDG_OK;
}

// sigma pass + rhs pass on one GPU
int cns_stage(esdg_ctx* ctx, const double* Qsrc, int mode, double a, double b, double dt, int want) {
    int rc = cns_sigma_pass(ctx, Qsrc, want);
    if (rc) return rc;
    return cns_rhs_pass(ctx, Qsrc, mode, a, b, dt, want);
}

// ---- path selection: tensor-product hex kernels when the operators allow, else generic ----
int total_partials(const esdg_ctx* ctx) {
    // hex path: one slot per element (a batch writes the slot of its first element, the rest stay zero), so the
    // deterministic reduction does not depend on which stage kernel / batch size / element split produced them
    return (ctx->hex || ctx->quad) ? (int)ctx->d.K : grid_blocks(ctx);
}

int reduce_rhstest(esdg_ctx* ctx, bool split) {
    // a split stage (boundary + interior launches) lays its partials out per range; an unsplit one
    // uses ceil(K/EB) slots
    (void)split;
    const int n = total_partials(ctx);
    tick(ctx, "reduce_partials");
    k_reduce_partials<<<1, 256, 0, ctx->stream>>>(ctx->partials, n, ctx->rhstest_dev);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

// traces (and whatever else the next rhs needs) of state Q
int do_traces(esdg_ctx* ctx, const double* Q) {
    if (ctx->cns.on) return cns_traces(ctx, Q);
    if (ctx->hex) return hexfast_traces(ctx, Q);
    if (ctx->quad) return quadfast_traces(ctx, Q);
    return launch_traces(ctx, Q);
}

// rhs (mode 0) or fused LSRK stage (mode 1).  After a mode-1 call the traces of the NEW state are
// valid on both paths (the hex kernel produces them in its epilogue; the generic path launches
// its trace kernel).
int do_stage(esdg_ctx* ctx, const double* Q, int mode, double a, double b, double dt, int want, int part) {
    int rc;
    if (ctx->cns.on) {
        if (part == ESDG_PART_INTERIOR) return ESDG_OK;
        if (ctx->d.n_halo != 0) {   // distributed: esdg_cns_sigma_pass + the sigma-trace exchange came first
            if (!ctx->cns.sigma_valid) return fail(ctx, ESDG_EINVAL, "esdg_stage_update: call esdg_cns_sigma_pass (and exchange the sigma traces) first");
            ctx->cns.sigma_valid = false;
            rc = cns_rhs_pass(ctx, Q, mode, a, b, dt, want);
        } else {
            rc = cns_stage(ctx, Q, mode, a, b, dt, want);
        }
        if (rc) return rc;
        if (mode == 1 && !ctx->quad) return cns_traces(ctx, ctx->Q);   // the tensor-product stage kernel wrote them
        return ESDG_OK;
    }
    if (ctx->hex) {
        if ((rc = hexfast_stage(ctx, Q, mode, a, b, dt, want, part))) return rc;
        if (want && part != ESDG_PART_BOUNDARY) return reduce_rhstest(ctx, part == ESDG_PART_INTERIOR);
        return ESDG_OK;
    }
    if (ctx->quad) {
        if ((rc = quadfast_stage(ctx, Q, mode, a, b, dt, want, part))) return rc;
        if (want && part != ESDG_PART_BOUNDARY) return reduce_rhstest(ctx, part == ESDG_PART_INTERIOR);
        return ESDG_OK;
    }
    if (part == ESDG_PART_INTERIOR) return ESDG_OK;   // generic path does not split: all work in the first call
    if ((rc = launch_rhs(ctx, Q, mode, a, b, dt, want))) return rc;
    if (mode == 1) return launch_traces(ctx, ctx->Q);
    return ESDG_OK;
}

int check_bad(esdg_ctx* ctx) {
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->bad_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        CK(cudaMemsetAsync(ctx->bad_flag, 0, sizeof(int), ctx->stream));
        return fail(ctx, ESDG_ESTATE, "non-positive density or internal energy met in rhs (Julia: DomainError)");
    }
    return ESDG_OK;
}

const double RK4A[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                        -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
const double RK4B[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                        1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                        2277821191437.0 / 14882151754819.0};  // src/CommonUtils.jl:29-49

}  // namespace

static void comm_destroy(esdg_ctx* ctx);

extern "C" {

const char* esdg_version(void) { return ESDG_VERSION; }

int esdg_create(const esdg_desc* desc, int device, esdg_ctx** out) {
    if (!desc || !out) return ESDG_EINVAL;
    *out = nullptr;
    esdg_ctx* ctx = new esdg_ctx();
    ctx->d = *desc;
    ctx->device = device;
    *out = ctx;  // returned even on failure so that esdg_last_error works
    const esdg_desc& d = ctx->d;
    if (!(d.dim == 2 || d.dim == 3) || d.Nfields != d.dim + 2 || d.K <= 0 || d.Ns <= 0 || d.Nq <= 0 ||
        d.Nfq <= 0 || d.Nfaces <= 0 || d.Nfq % d.Nfaces != 0 || d.n_halo < 0)
        return fail(ctx, ESDG_EINVAL, "esdg_create: inconsistent descriptor");
    if (d.affine != 1 && !(d.affine == 0 && d.dim == 3 && d.collocated))
        return fail(ctx, ESDG_EINVAL, "esdg_create: desc.affine = 0 (curved elements) is supported for collocated hexes only; "
                                      "the 2-D scripts read the geometric factors of the first node (`[1,e]`)");
    if (d.collocated && d.Ns != d.Nq) return fail(ctx, ESDG_EINVAL, "esdg_create: collocated needs Ns == Nq");
    if ((double)d.Nfq * (double)d.K + (double)d.n_halo > 2.0e9)
        return fail(ctx, ESDG_EINVAL, "esdg_create: Nfq*K+n_halo exceeds 32-bit trace indexing");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ctx, ESDG_ECUDA, "esdg_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(ctx, ESDG_EINVAL, "esdg_create: bad device ordinal");
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
    ctx->Nh = d.Nq + d.Nfq;
    ctx->Nfp = d.Nfq / d.Nfaces;
    ctx->NF = d.Nfields;
    ctx->trace_stride = (long long)d.Nfq * d.K + d.n_halo;
    return ESDG_OK;
}

int esdg_destroy(esdg_ctx* ctx) {
    if (!ctx) return ESDG_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& t : ctx->timers) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (cudaEvent_t e : ctx->chunk_up) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->chunk_done) cudaEventDestroy(e);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    comm_destroy(ctx);
    for (void* p : ctx->allocs) cudaFree(p);
    hexfast_destroy(ctx);
    quadfast_destroy(ctx);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ESDG_OK;
}

const char* esdg_last_error(const esdg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int esdg_set_stream(esdg_ctx* ctx, void* s) {
    if (!ctx) return ESDG_EINVAL;
    if (ctx->own_stream && ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    ctx->stream = (cudaStream_t)s;
    ctx->own_stream = false;
    return ESDG_OK;
}

int esdg_set_operator(esdg_ctx* ctx, const char* name, const double* a, int64_t rows, int64_t cols) {
    if (!ctx || !name || !a || rows <= 0 || cols <= 0) return fail(ctx, ESDG_EINVAL, "esdg_set_operator: bad argument");
    static const char* known[] = {"Ef", "Vq", "VhP", "Ph", "Lf", "Qrh", "Qsh", "Qth", "wq", "Pq", "Vf", "Dr", "Ds"};
    bool ok = false;
    for (const char* k : known) ok = ok || !strcmp(k, name);
    if (!ok) return fail(ctx, ESDG_EINVAL, std::string("esdg_set_operator: unknown operator ") + name);
    HostMat m;
    m.rows = rows; m.cols = cols;
    m.a.assign(a, a + rows * cols);
    ctx->ops[name] = std::move(m);
    if (ctx->finalized) return fail(ctx, ESDG_EINVAL, "setup calls (esdg_set_*) must precede esdg_finalize: a finalized context owns its device tables; create a new context to change them");
    return ESDG_OK;
}

int esdg_set_geometry(esdg_ctx* ctx, const char* name, const double* a, int64_t rows, int64_t cols) {
    if (!ctx || !name || !a || rows <= 0) return fail(ctx, ESDG_EINVAL, "esdg_set_geometry: bad argument");
    if (cols != ctx->d.K) return fail(ctx, ESDG_EINVAL, "esdg_set_geometry: cols must equal desc.K");
    static const char* known[] = {"rxJ", "sxJ", "txJ", "ryJ", "syJ", "tyJ", "rzJ", "szJ", "tzJ",
                                  "J", "nxJ", "nyJ", "nzJ", "sJ", "wJq"};
    bool ok = false;
    for (const char* k : known) ok = ok || !strcmp(k, name);
    if (!ok) return fail(ctx, ESDG_EINVAL, std::string("esdg_set_geometry: unknown array ") + name);
    HostMat m;
    m.rows = rows; m.cols = cols;
    m.a.assign(a, a + rows * cols);
    ctx->geom[name] = std::move(m);
    if (ctx->finalized) return fail(ctx, ESDG_EINVAL, "setup calls (esdg_set_*) must precede esdg_finalize: a finalized context owns its device tables; create a new context to change them");
    return ESDG_OK;
}

int esdg_set_maps(esdg_ctx* ctx, const int64_t* mapP, int64_t n) {
    if (!ctx || !mapP) return fail(ctx, ESDG_EINVAL, "esdg_set_maps: bad argument");
    if (n != (int64_t)ctx->d.Nfq * ctx->d.K) return fail(ctx, ESDG_EINVAL, "esdg_set_maps: mapP must hold Nfq*K entries");
    ctx->mapP_host.assign(mapP, mapP + n);
    if (ctx->finalized) return fail(ctx, ESDG_EINVAL, "setup calls (esdg_set_*) must precede esdg_finalize: a finalized context owns its device tables; create a new context to change them");
    return ESDG_OK;
}

int esdg_set_viscosity(esdg_ctx* ctx, double lambda, double mu, double Pr, double Re, int viscous_dissp) {
    if (!ctx) return ESDG_EINVAL;
    if (ctx->d.dim != 2 || ctx->d.collocated)
        return fail(ctx, ESDG_EINVAL, "esdg_set_viscosity: the CNS path is 2D with the Vq/Pq projection (desc.dim = 2, collocated = 0)");
    if (!(mu >= 0.0) || !(Pr > 0.0) || !(Re > 0.0)) return fail(ctx, ESDG_EINVAL, "esdg_set_viscosity: need mu >= 0, Pr > 0, Re > 0");
    CnsState& c = ctx->cns;
    c.on = true;
    c.lambda = lambda; c.mu = mu; c.Pr = Pr; c.Re = Re;
    c.viscous_dissp = viscous_dissp != 0;
    if (ctx->finalized) return fail(ctx, ESDG_EINVAL, "setup calls (esdg_set_*) must precede esdg_finalize: a finalized context owns its device tables; create a new context to change them");
    return ESDG_OK;
}

int esdg_set_wall_bc(esdg_ctx* ctx, int bctype, const int64_t* wall_idx, int64_t nwall, const int64_t* lid_idx,
                     const double* vlid, int64_t nlid) {
    if (!ctx) return ESDG_EINVAL;
    if (bctype < 1 || bctype > 3) return fail(ctx, ESDG_EINVAL, "esdg_set_wall_bc: BCTYPE must be 1 (adiabatic), 2 (isothermal) or 3 (slip)");
    if (nwall < 0 || nlid < 0 || (nwall > 0 && !wall_idx) || (nlid > 0 && (!lid_idx || !vlid)))
        return fail(ctx, ESDG_EINVAL, "esdg_set_wall_bc: bad argument");
    CnsState& c = ctx->cns;
    c.bctype = bctype;
    c.wall.assign(wall_idx, wall_idx + nwall);
    c.lid.assign(lid_idx, lid_idx + nlid);
    c.vlid.assign(vlid, vlid + nlid);
    if (ctx->finalized) return fail(ctx, ESDG_EINVAL, "setup calls (esdg_set_*) must precede esdg_finalize: a finalized context owns its device tables; create a new context to change them");
    return ESDG_OK;
}

int esdg_set_inflow_outflow_bc(esdg_ctx* ctx, const int64_t* inflow_idx, int64_t n_in, const double* state, const int64_t* outflow_idx,
                               int64_t n_out) {
    if (!ctx) return ESDG_EINVAL;
    if (n_in < 0 || n_out < 0 || (n_in > 0 && (!inflow_idx || !state)) || (n_out > 0 && !outflow_idx))
        return fail(ctx, ESDG_EINVAL, "esdg_set_inflow_outflow_bc: bad argument");
    if (ctx->finalized) return fail(ctx, ESDG_EINVAL, "esdg_set_inflow_outflow_bc: call before esdg_finalize");
    CnsState& c = ctx->cns;
    if (n_in > 0) {
        if (!(state[0] > 0.0) || !(state[3] > 0.0)) return fail(ctx, ESDG_EINVAL, "esdg_set_inflow_outflow_bc: inflow density and pressure must be positive");
        for (int k = 0; k < 4; ++k) c.inflow_state[k] = state[k];
    }
    c.inflow.assign(inflow_idx, inflow_idx + n_in);
    c.outflow.assign(outflow_idx, outflow_idx + n_out);
    return ESDG_OK;
}

int esdg_set_boundary_nodes(esdg_ctx* ctx, const int64_t* mapB, int64_t nB) {
    if (!ctx || nB < 0 || (nB > 0 && !mapB)) return fail(ctx, ESDG_EINVAL, "esdg_set_boundary_nodes: bad argument");
    ctx->cns.mapB.assign(mapB, mapB + nB);
    ctx->cns.have_mapB = true;
    if (ctx->finalized) return fail(ctx, ESDG_EINVAL, "setup calls (esdg_set_*) must precede esdg_finalize: a finalized context owns its device tables; create a new context to change them");
    return ESDG_OK;
}

// CNS part of esdg_finalize: extra operators, BC tables, sigma storage, launch shape.
static int finalize_cns(esdg_ctx* ctx) {
    const esdg_desc& d = ctx->d;
    CnsState& c = ctx->cns;
    const int Nh = ctx->Nh;
    struct { const char* n; int64_t r, cl; DevEll* out; Dense* dout; } need[4] = {
        {"Pq", d.Ns, d.Nq, &c.Pq, &c.dPq}, {"Vf", d.Nfq, d.Ns, &c.Vf, &c.dVf}, {"Dr", d.Ns, d.Ns, &c.Dr, &c.dDr},
        {"Ds", d.Ns, d.Ns, &c.Ds, &c.dDs}};
    int rc;
    for (auto& q : need) {
        const HostMat* m = find(ctx->ops, q.n);
        if (!m || m->rows != q.r || m->cols != q.cl)
            return fail(ctx, ESDG_EINVAL, std::string("esdg_finalize: CNS operator missing or mis-sized: ") + q.n);
        if ((rc = build_ell(ctx, *m, 0.0, q.out))) return rc;
        if ((rc = build_dense(ctx, *m, 1.0, 0, (int)q.r, q.dout))) return rc;
    }
    const int64_t ntr = (int64_t)d.Nfq * d.K;
    std::vector<unsigned char> bc((size_t)ntr, 0);
    std::vector<double> vl;
    auto mark = [&](const std::vector<int64_t>& idx, unsigned char bits, bool exclusive) -> bool {
        for (int64_t v1 : idx) {
            const int64_t v = v1 - 1;
            if (v < 0 || v >= ntr) return false;
            if (exclusive && (bc[(size_t)v] & 7)) return false;
            bc[(size_t)v] |= bits;
        }
        return true;
    };
    if (!mark(c.wall, 1, true) || !mark(c.lid, 2, true) || !mark(c.inflow, 3, true) || !mark(c.outflow, 4, true))
        return fail(ctx, ESDG_EINVAL, "esdg_finalize: wall/lid/inflow/outflow index out of range or listed twice");
    if (!c.lid.empty()) {
        vl.assign((size_t)ntr, 0.0);
        for (size_t i = 0; i < c.lid.size(); ++i) vl[(size_t)(c.lid[i] - 1)] = c.vlid[i];
    }
    if (c.have_mapB) {
        if (!mark(c.mapB, 8, false)) return fail(ctx, ESDG_EINVAL, "esdg_finalize: mapB index out of range");
    } else {
        for (auto& b : bc) if ((b & 7) == 1 || (b & 7) == 2) b |= 8;   // default: md.mapB = wall + lid (cavity, :139-147)
    }
    // a wall/lid node must be a boundary node of the mesh: mapP == mapM there (SetupDG.jl:294-298); inflow / outflow
    // nodes may be connected to anything (the shock-tube script makes the box periodic first)
    for (int64_t i = 0; i < ntr; ++i)
        if (((bc[(size_t)i] & 7) == 1 || (bc[(size_t)i] & 7) == 2) && ctx->mapP_host[(size_t)i] - 1 != i)
            return fail(ctx, ESDG_EINVAL, "esdg_finalize: wall/lid node is not a boundary node (mapP != mapM)");
    if ((rc = dev_upload(ctx, &c.bc, bc))) return rc;
    if (!vl.empty() && (rc = dev_upload(ctx, &c.vlid_dev, vl))) return rc;
    if ((rc = dev_alloc(ctx, &c.sig, (size_t)6 * d.Ns * d.K))) return rc;
    if ((rc = dev_alloc(ctx, &c.sig_tr, (size_t)6 * ctx->trace_stride))) return rc;
    // launch shape: generic modal block + CNS extras
    int maxsm = 0;
    CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    size_t tbl = ell_stage_bytes(ctx->Vq.rows, ctx->Vq.width) + ell_stage_bytes(ctx->Proj.rows, ctx->Proj.width) +
                 ell_stage_bytes(ctx->Ph.rows, ctx->Ph.width) + ell_stage_bytes(ctx->Lf.rows, ctx->Lf.width) +
                 pair_stage_bytes(Nh, ctx->qp_width) + ell_stage_bytes(c.Pq.rows, c.Pq.width) +
                 ell_stage_bytes(c.Vf.rows, c.Vf.width) + ell_stage_bytes(c.Dr.rows, c.Dr.width) +
                 ell_stage_bytes(c.Ds.rows, c.Ds.width);
    ctx->stage_tables = tbl <= 64 * 1024 && getenv("ESDG_STAGE_TABLES") != nullptr;   // see esdg_finalize
    if (!ctx->stage_tables) tbl = 0;
    auto bytes = [&](int EB) {
        return generic_smem_bytes(2, EB, d.Ns, d.Nq, d.Nfq, Nh, 0) + cns_extra_doubles(EB, d.Ns, d.Nq, d.Nfq) * sizeof(double);
    };
    // 4 elements x 4 fields = 16 GEMM columns (two 8-wide DMMA tiles, no padding columns) in 128-thread CTAs:
    // several CTAs per SM hide the barriers between the GEMM and pointwise phases (measured best of 2..10)
    int EB = std::min(4, std::max(1, ESDG_GENERIC_MAX_THREADS / Nh));
    if (const char* ev = getenv("ESDG_CNS_EB")) EB = std::max(1, atoi(ev));   // launch-shape experiments
    while (EB > 1 && bytes(EB) + tbl > (size_t)maxsm / 2) --EB;
    if (bytes(EB) + tbl > (size_t)maxsm) return fail(ctx, ESDG_EINVAL, "esdg_finalize: element too large for shared memory (CNS)");
    ctx->EB = EB;
    ctx->table_off = (unsigned)bytes(EB);
    ctx->smem_bytes = bytes(EB) + tbl;
    ctx->threads = std::min(ESDG_GENERIC_MAX_THREADS, ((EB * Nh + 31) / 32) * 32);
    if (const char* tv = getenv("ESDG_CNS_THREADS")) ctx->threads = std::max(32, (atoi(tv) / 32) * 32);   // launch-shape experiments
    ctx->smem_cns_sigma = ctx->stage_tables ? ctx->smem_bytes : cns_sigma_smem_bytes(EB, d.Ns, d.Nq, d.Nfq);
    ctx->smem_cns_traces = ctx->stage_tables ? ctx->smem_bytes : cns_traces_smem_bytes(EB, d.Ns, d.Nq, d.Nfq);
    CK(cudaFuncSetAttribute(k_cns_traces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_cns_traces));
    CK(cudaFuncSetAttribute(k_cns_sigma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_cns_sigma));
    ctx->smem_cns_rhs = ctx->stage_tables ? ctx->smem_bytes
                                          : generic_smem_bytes(2, EB, d.Ns, d.Nq, d.Nfq, Nh, 0) + cns_rhs_extra_doubles(EB, d.Ns, d.Nq, d.Nfq) * sizeof(double);
    CK(cudaFuncSetAttribute(k_cns_rhs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_cns_rhs));
    const size_t nb = (size_t)grid_blocks(ctx);
    if ((rc = dev_alloc(ctx, &c.partial_visc, nb))) return rc;
    if ((rc = dev_alloc(ctx, &c.partial_vrhs, nb))) return rc;
    CK(cudaMemsetAsync(c.partial_visc, 0, nb * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(c.partial_vrhs, 0, nb * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(c.sig, 0, (size_t)6 * d.Ns * d.K * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(c.sig_tr, 0, (size_t)6 * ctx->trace_stride * sizeof(double), ctx->stream));
    return ESDG_OK;
}

}  // extern "C"

namespace {
// GPU-side setup of a structured hex mesh: what src/UniformHexMesh.jl + connect_mesh.jl +
// node_map_functions.jl:139-213 (periodic maps) + geometric_factors.jl produce for an axis-aligned box, written straight
// into the device tables.  Element e = i + Kx j + Kx Ky (k - kz0) (x fastest, UniformHexMesh.jl:59-63); faces in the
// order r-, r+, s-, s+, t-, t+; matching face points share their on-face index m.  Points owned by the z-slab below /
// above map to halo slots Nfq K + (i + Kx j) nfp + m (+ Kx Ky nfp for the upper slab), as esdg_cns_b200/setup_dg.py does.
__global__ void k_structured_mapP(int* __restrict__ mapP, long long K, int Kx, int Ky, long long Kz, long long kz0, long long kzn,
                                  int nfp, int px, int py, int pz) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Nfq = 6 * nfp;
    if (t >= K * Nfq) return;
    const long long e = t / Nfq;
    const int kf = (int)(t - e * Nfq), f = kf / nfp, m = kf - f * nfp;
    const int i = (int)(e % Kx), j = (int)((e / Kx) % Ky);
    const long long k = e / ((long long)Kx * Ky) + kz0;
    const int d = f >> 1, sgn = (f & 1) ? 1 : -1, fn = f ^ 1;
    long long c[3] = {i, j, k};
    const long long Kd[3] = {Kx, Ky, Kz};
    const int per[3] = {px, py, pz};
    c[d] += sgn;
    bool ok = true;
    if (per[d]) c[d] = (c[d] + Kd[d]) % Kd[d];
    else ok = c[d] >= 0 && c[d] < Kd[d];
    long long tgt;
    if (!ok) tgt = t;                                                    // physical boundary: mapP = mapM
    else if (d == 2 && (c[2] < kz0 || c[2] >= kz0 + kzn))                // owned by a neighbouring slab: halo slot
        tgt = (long long)Nfq * K + ((long long)c[0] + (long long)Kx * c[1]) * nfp + m + (sgn < 0 ? 0 : (long long)Kx * Ky * nfp);
    else tgt = (long long)(fn * nfp + m) + Nfq * (c[0] + (long long)Kx * c[1] + (long long)Kx * Ky * (c[2] - kz0));
    mapP[t] = (int)tgt;
}
__global__ void k_structured_geo(double* __restrict__ geo, long long K, int stride, double hx, double hy, double hz) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K) return;
    double* g = geo + e * stride;
    for (int q = 0; q < stride; ++q) g[q] = 0.0;
    const double a[3] = {hy * hz / 4, hx * hz / 4, hx * hy / 4};          // rxJ, syJ, tzJ
    g[0] = a[0]; g[4] = a[1]; g[8] = a[2];
    g[9] = hx * hy * hz / 8;
    for (int f = 0; f < 6; ++f) {
        const int d = f >> 1;
        g[10 + 4 * f + d] = (f & 1) ? a[d] : -a[d];
        g[10 + 4 * f + 3] = a[d];
    }
}
}  // namespace

extern "C" {

int esdg_set_structured_hex(esdg_ctx* ctx, const int64_t K3[3], const double box[6], const int periodic[3], int64_t kz_begin,
                            int64_t kz_count) {
    if (!ctx || !K3 || !box || !periodic) return fail(ctx, ESDG_EINVAL, "esdg_set_structured_hex: bad argument");
    if (ctx->finalized) return fail(ctx, ESDG_EINVAL, "setup calls (esdg_set_*) must precede esdg_finalize");
    const esdg_desc& d = ctx->d;
    if (d.dim != 3 || d.elem != ESDG_HEX || d.Nfaces != 6 || !d.affine)
        return fail(ctx, ESDG_EINVAL, "esdg_set_structured_hex: affine hex contexts only");
    if (K3[0] < 1 || K3[1] < 1 || K3[2] < 1 || kz_begin < 0 || kz_count < 1 || kz_begin + kz_count > K3[2] ||
        K3[0] * K3[1] * kz_count != d.K)
        return fail(ctx, ESDG_EINVAL, "esdg_set_structured_hex: element counts do not match desc.K");
    const bool split = kz_count < K3[2];
    const long long need_halo = split ? 2 * K3[0] * K3[1] * (long long)ctx->Nfp : 0;
    if (d.n_halo != need_halo) return fail(ctx, ESDG_EINVAL, "esdg_set_structured_hex: desc.n_halo must be 2 Kx Ky Nfp for a z-slab (0 for the whole mesh)");
    if (!(box[1] > box[0] && box[3] > box[2] && box[5] > box[4])) return fail(ctx, ESDG_EINVAL, "esdg_set_structured_hex: empty box");
    ctx->structured = true;
    for (int q = 0; q < 3; ++q) { ctx->sK[q] = K3[q]; ctx->sper[q] = periodic[q] != 0; }
    for (int q = 0; q < 6; ++q) ctx->sbox[q] = box[q];
    ctx->skz0 = kz_begin;
    ctx->skzn = kz_count;
    return ESDG_OK;
}

int esdg_finalize(esdg_ctx* ctx) {
    if (!ctx) return ESDG_EINVAL;
    if (ctx->finalized) return fail(ctx, ESDG_EINVAL, "esdg_finalize: context already finalized");
    const esdg_desc& d = ctx->d;
    const int DIM = d.dim, NF = d.Nfields, Nh = ctx->Nh;
    CK(cudaSetDevice(ctx->device));
    // ---- operators -----------------------------------------------------------
    auto need = [&](const char* n, int64_t r, int64_t c) -> const HostMat* {
        const HostMat* m = find(ctx->ops, n);
        if (!m) { fail(ctx, ESDG_EINVAL, std::string("esdg_finalize: operator missing: ") + n); return nullptr; }
        if (m->rows != r || m->cols != c) {
            char buf[160];
            snprintf(buf, sizeof buf, "esdg_finalize: operator %s is %lldx%lld, expected %lldx%lld", n,
                     (long long)m->rows, (long long)m->cols, (long long)r, (long long)c);
            fail(ctx, ESDG_EINVAL, buf);
            return nullptr;
        }
        return m;
    };
    const double TOL = 1e-12;  // droptol of dg2D_euler_quad.jl:62-63,76-78
    int rc;
    if (d.collocated) {
        const HostMat* Ef = need("Ef", d.Nfq, d.Nq);
        if (!Ef) return ESDG_EINVAL;
        if ((rc = build_ell(ctx, *Ef, TOL, &ctx->Proj))) return rc;
        HostMat eye; eye.rows = eye.cols = 1; eye.a = {1.0};
        if ((rc = build_ell(ctx, eye, TOL, &ctx->Vq))) return rc;
    } else {
        const HostMat* Vq = need("Vq", d.Nq, d.Ns);
        const HostMat* VhP = need("VhP", Nh, d.Nq);
        if (!Vq || !VhP) return ESDG_EINVAL;
        if ((rc = build_ell(ctx, *Vq, 0.0, &ctx->Vq))) return rc;
        if ((rc = build_ell(ctx, *VhP, 0.0, &ctx->Proj))) return rc;
        if ((rc = build_dense(ctx, *Vq, 1.0, 0, d.Nq, &ctx->dVq))) return rc;
        if ((rc = build_dense(ctx, *VhP, 1.0, 0, Nh, &ctx->dProj))) return rc;
        if ((rc = build_dense(ctx, *VhP, 1.0, d.Nq, d.Nfq, &ctx->dProjF))) return rc;
    }
    const HostMat* Ph = need("Ph", d.Ns, Nh);
    const HostMat* Lf = need("Lf", d.Ns, d.Nfq);
    const HostMat* wq = need("wq", d.Nq, 1);
    if (!Ph || !Lf || !wq) return ESDG_EINVAL;
    if ((rc = build_ell(ctx, *Ph, d.collocated ? TOL : 0.0, &ctx->Ph))) return rc;
    if ((rc = build_ell(ctx, *Lf, d.collocated ? TOL : 0.0, &ctx->Lf))) return rc;
    if (!d.collocated) {
        if ((rc = build_dense(ctx, *Ph, d.volume_scale, 0, d.Ns, &ctx->dPh))) return rc;
        if ((rc = build_dense(ctx, *Lf, 1.0, 0, d.Ns, &ctx->dLf))) return rc;
    }
    if ((rc = dev_upload(ctx, &ctx->wq, wq->a))) return rc;
    const char* qn[3] = {"Qrh", "Qsh", "Qth"};
    const HostMat* Qm[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < DIM; ++k)
        if (!(Qm[k] = need(qn[k], Nh, Nh))) return ESDG_EINVAL;
    {
        std::vector<int> cnt(Nh, 0);
        int width = 1;
        auto nz = [&](int i, int j) {
            for (int k = 0; k < DIM; ++k)
                if (std::fabs(Qm[k]->at(i, j)) > TOL) return true;
            return false;
        };
        for (int i = 0; i < Nh; ++i) {
            for (int j = 0; j < Nh; ++j) cnt[i] += nz(i, j);
            width = std::max(width, cnt[i]);
        }
        std::vector<int> col((size_t)Nh * width, 0);
        std::vector<double> c((size_t)Nh * width * 3, 0.0);
        for (int i = 0; i < Nh; ++i) {
            int k = 0;
            for (int j = 0; j < Nh; ++j)
                if (nz(i, j)) {
                    col[(size_t)i * width + k] = j;
                    for (int q = 0; q < DIM; ++q) {
                        const double v = Qm[q]->at(i, j);
                        c[((size_t)i * width + k) * 3 + q] = std::fabs(v) > TOL ? v : 0.0;
                    }
                    ++k;
                }
        }
        ctx->qp_width = width;
        // ---- symmetric schedule: every unordered pair once (operators skew to round-off, else stay row-wise)
        {
            bool skew = true;
            double qmax = 0.0;
            for (int k = 0; k < DIM; ++k)
                for (int i = 0; i < Nh; ++i)
                    for (int j = 0; j < Nh; ++j) qmax = std::max(qmax, std::fabs(Qm[k]->at(i, j)));
            for (int k = 0; k < DIM && skew; ++k)
                for (int i = 0; i < Nh && skew; ++i)
                    for (int j = i; j < Nh; ++j)
                        if (std::fabs(Qm[k]->at(i, j) + Qm[k]->at(j, i)) > 1e-13 * qmax) { skew = false; break; }
            const char* ev = getenv("ESDG_SYM_PAIRS");
            if (ev && atoi(ev) == 0) skew = false;
            ctx->sym_rounds = 0;
            if (skew) {
                std::vector<std::pair<int, int>> edges;
                for (int i = 0; i < Nh; ++i)
                    for (int j = i + 1; j < Nh; ++j)
                        if (nz(i, j) || nz(j, i)) edges.push_back({i, j});
                // greedy: an edge goes to the earliest round in which its source has not evaluated and its
                // destination has not been written, trying both orientations; high-degree rows first
                std::vector<int> deg(Nh, 0);
                for (auto& e2 : edges) { deg[e2.first]++; deg[e2.second]++; }
                std::stable_sort(edges.begin(), edges.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
                    return std::max(deg[a.first], deg[a.second]) > std::max(deg[b.first], deg[b.second]);
                });
                std::vector<std::vector<char>> src_busy, dst_busy;
                std::vector<std::vector<int>> dst_of;
                auto ensure = [&](size_t t) {
                    while (src_busy.size() <= t) {
                        src_busy.emplace_back(Nh, 0);
                        dst_busy.emplace_back(Nh, 0);
                        dst_of.emplace_back(Nh, -1);
                    }
                };
                auto place_greedy = [&]() {
                    for (auto& e2 : edges) {
                        for (size_t t = 0;; ++t) {
                            ensure(t);
                            int a = e2.first, b = e2.second;
                            bool ok1 = !src_busy[t][a] && !dst_busy[t][b], ok2 = !src_busy[t][b] && !dst_busy[t][a];
                            if (!ok1 && !ok2) continue;
                            if (!ok1) std::swap(a, b);
                            src_busy[t][a] = 1; dst_busy[t][b] = 1; dst_of[t][a] = b;
                            break;
                        }
                    }
                };
                place_greedy();
                int rounds = (int)dst_of.size();
                // circulant schedule i -> (i + k) mod Nh: optimal for dense operators (triangles)
                const int circ = Nh / 2;
                if (circ < rounds) {
                    dst_of.assign(circ, std::vector<int>(Nh, -1));
                    for (int k = 1; k <= circ; ++k)
                        for (int i = 0; i < Nh; ++i) {
                            if (Nh % 2 == 0 && k == circ && i >= circ) continue;
                            const int j = (i + k) % Nh;
                            if (nz(i, j) || nz(j, i)) dst_of[k - 1][i] = j;
                        }
                    rounds = circ;
                }
                std::vector<int> sd((size_t)rounds * Nh, -1);
                std::vector<double> sc((size_t)rounds * Nh * 3, 0.0);
                for (int t = 0; t < rounds; ++t)
                    for (int i = 0; i < Nh; ++i) {
                        const int j = dst_of[t][i];
                        sd[(size_t)t * Nh + i] = j;
                        if (j >= 0)
                            for (int q = 0; q < DIM; ++q) {
                                const double v = Qm[q]->at(i, j);
                                sc[((size_t)t * Nh + i) * 3 + q] = std::fabs(v) > TOL ? v : 0.0;
                            }
                    }
                // a round costs a CTA barrier and idles the rows without a partner: measured on B200 the schedule wins on
                // the dense triangle operators (config 2, 15 rounds against 30 entries per row: 1.04 -> 0.98 ms) and loses
                // on the sparse quad operators (4.97 -> 5.36 ms), so it is used only when it really halves the row length
                const bool force = ev && atoi(ev) == 2;
                if (force || 2 * rounds <= width) {
                    ctx->sym_rounds = rounds;
                    if ((rc = dev_upload(ctx, &ctx->sym_dst, sd))) return rc;
                    if ((rc = dev_upload(ctx, &ctx->sym_c, sc))) return rc;
                }
            }
        }
        if ((rc = dev_upload(ctx, &ctx->qp_cnt, cnt))) return rc;
        if ((rc = dev_upload(ctx, &ctx->qp_col, col))) return rc;
        if ((rc = dev_upload(ctx, &ctx->qp_c, c))) return rc;
    }
    // ---- geometry (affine: one record per element) -----------------------------
    static const char* g2[4] = {"rxJ", "sxJ", "ryJ", "syJ"};
    static const char* g3[9] = {"rxJ", "sxJ", "txJ", "ryJ", "syJ", "tyJ", "rzJ", "szJ", "tzJ"};
    static const char* nn[4] = {"nxJ", "nyJ", "nzJ", "sJ"};
    ctx->geo_stride = DIM * DIM + 1 + d.Nfaces * (DIM + 1);
    std::vector<double> geo(ctx->structured ? 0 : (size_t)d.K * ctx->geo_stride);
    if (ctx->structured) {
        if ((rc = dev_alloc(ctx, &ctx->geo, (size_t)d.K * ctx->geo_stride))) return rc;
        const double hx = (ctx->sbox[1] - ctx->sbox[0]) / ctx->sK[0], hy = (ctx->sbox[3] - ctx->sbox[2]) / ctx->sK[1],
                     hz = (ctx->sbox[5] - ctx->sbox[4]) / ctx->sK[2];
        k_structured_geo<<<(unsigned)((d.K + 255) / 256), 256, 0, ctx->stream>>>(ctx->geo, d.K, ctx->geo_stride, hx, hy, hz);
        CK(cudaGetLastError());
    } else {
        const HostMat* G[9];
        for (int k = 0; k < DIM * DIM; ++k) {
            G[k] = find(ctx->geom, DIM == 2 ? g2[k] : g3[k]);
            if (!G[k]) return fail(ctx, ESDG_EINVAL, std::string("esdg_finalize: geometry missing: ") + (DIM == 2 ? g2[k] : g3[k]));
        }
        const HostMat* J = find(ctx->geom, "J");
        if (!J) return fail(ctx, ESDG_EINVAL, "esdg_finalize: geometry missing: J");
        const HostMat* Nn[4];
        for (int k = 0; k <= DIM; ++k) {
            const char* nm = (k == DIM) ? "sJ" : nn[k];
            Nn[k] = find(ctx->geom, nm);
            if (!Nn[k]) return fail(ctx, ESDG_EINVAL, std::string("esdg_finalize: geometry missing: ") + nm);
            if (Nn[k]->rows != d.Nfq && Nn[k]->rows != d.Nfaces)
                return fail(ctx, ESDG_EINVAL, std::string("esdg_finalize: rows of ") + nm + " must be Nfq or Nfaces");
        }
        for (int64_t e = 0; e < d.K; ++e) {
            double* g = &geo[(size_t)e * ctx->geo_stride];
            for (int k = 0; k < DIM * DIM; ++k) g[k] = G[k]->at(0, e);        // `[1,e]`, dg2D_euler_quad.jl:175
            g[DIM * DIM] = J->at(0, e);
            if (!(g[DIM * DIM] > 0.0)) return fail(ctx, ESDG_EINVAL, "esdg_finalize: non-positive Jacobian J");
            for (int f = 0; f < d.Nfaces; ++f)
                for (int k = 0; k <= DIM; ++k) {
                    const int64_t row0 = (Nn[k]->rows == d.Nfaces) ? f : (int64_t)f * ctx->Nfp;
                    const double v0 = Nn[k]->at(row0, e);
                    g[DIM * DIM + 1 + f * (DIM + 1) + k] = v0;
                    if (Nn[k]->rows == d.Nfq && d.affine) {  // affine check: constant along the face
                        for (int q = 1; q < ctx->Nfp; ++q)
                            if (std::fabs(Nn[k]->at(row0 + q, e) - v0) > 1e-10 * (std::fabs(v0) + g[DIM * DIM]))
                                return fail(ctx, ESDG_EINVAL, "esdg_finalize: face normals vary along a face (mesh not affine)");
                    }
                }
            // affine check on the volume geofacs / J when full arrays were given
            for (int k = 0; k < DIM * DIM && d.affine; ++k)
                for (int64_t r = 1; r < G[k]->rows; ++r)
                    if (std::fabs(G[k]->at(r, e) - g[k]) > 1e-10 * (std::fabs(g[k]) + 1e-300 + std::fabs(g[DIM * DIM])))
                        return fail(ctx, ESDG_EINVAL, "esdg_finalize: volume geometric factors vary inside an element (mesh not affine)");
        }
    }
    if (!ctx->structured && (rc = dev_upload(ctx, &ctx->geo, geo))) return rc;
    if (!d.affine) {
        // curved hexes: the full arrays of the script, rxJ.. = [Vq;Vf]*rxJ.. (Nh x K), J = Vq*J (Nq x K),
        // nxJ.. at the face points (Nfq x K)   (dg3D_euler_hex.jl:69-101)
        std::vector<double> gG((size_t)d.K * 9 * Nh), gJ((size_t)d.K * d.Nq), gN((size_t)d.K * 4 * d.Nfq);
        const HostMat* J = find(ctx->geom, "J");
        if (J->rows != d.Nq) return fail(ctx, ESDG_EINVAL, "esdg_finalize: curved mesh needs J at every node (Nq x K)");
        for (int k = 0; k < 9; ++k) {
            const HostMat* G = find(ctx->geom, g3[k]);
            if (G->rows != Nh) return fail(ctx, ESDG_EINVAL, std::string("esdg_finalize: curved mesh needs ") + g3[k] + " at every hybridized point (Nh x K)");
            for (int64_t e = 0; e < d.K; ++e)
                for (int r = 0; r < Nh; ++r) gG[((size_t)e * 9 + k) * Nh + r] = G->at(r, e);
        }
        for (int64_t e = 0; e < d.K; ++e)
            for (int r = 0; r < d.Nq; ++r) {
                const double v = J->at(r, e);
                if (!(v > 0.0)) return fail(ctx, ESDG_EINVAL, "esdg_finalize: non-positive Jacobian J");
                gJ[(size_t)e * d.Nq + r] = v;
            }
        for (int k = 0; k < 4; ++k) {
            const HostMat* Nk = find(ctx->geom, nn[k]);
            if (Nk->rows != d.Nfq) return fail(ctx, ESDG_EINVAL, std::string("esdg_finalize: curved mesh needs ") + nn[k] + " at every face point (Nfq x K)");
            for (int64_t e = 0; e < d.K; ++e)
                for (int r = 0; r < d.Nfq; ++r) gN[((size_t)e * 4 + k) * d.Nfq + r] = Nk->at(r, e);
        }
        if ((rc = dev_upload(ctx, &ctx->geoG, gG))) return rc;
        if ((rc = dev_upload(ctx, &ctx->geoJ, gJ))) return rc;
        if ((rc = dev_upload(ctx, &ctx->geoN, gN))) return rc;
    }
    // ---- maps ------------------------------------------------------------------
    if (ctx->structured) {
        if ((rc = dev_alloc(ctx, &ctx->mapP, (size_t)d.Nfq * d.K))) return rc;
        const long long nt = (long long)d.Nfq * d.K;
        k_structured_mapP<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(ctx->mapP, d.K, (int)ctx->sK[0], (int)ctx->sK[1], ctx->sK[2],
                                                                                ctx->skz0, ctx->skzn, ctx->Nfp, ctx->sper[0], ctx->sper[1],
                                                                                ctx->sper[2]);
        CK(cudaGetLastError());
    } else if ((int64_t)ctx->mapP_host.size() != (int64_t)d.Nfq * d.K) {
        return fail(ctx, ESDG_EINVAL, "esdg_finalize: mapP missing (esdg_set_maps or esdg_set_structured_hex)");
    } else {
        std::vector<int> mp(ctx->mapP_host.size());
        for (size_t i = 0; i < mp.size(); ++i) {
            const int64_t v = ctx->mapP_host[i] - 1;  // Julia 1-based
            if (v < 0 || v >= ctx->trace_stride) return fail(ctx, ESDG_EINVAL, "esdg_finalize: mapP entry out of range");
            mp[i] = (int)v;
        }
        if ((rc = dev_upload(ctx, &ctx->mapP, mp))) return rc;
    }
    // ---- launch shape ------------------------------------------------------------
    int maxsm = 0;
    CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    // operator tables go to shared memory when they are small next to the element data (<= 48 KB)
    size_t tbl = ell_stage_bytes(ctx->Vq.rows, ctx->Vq.width) + ell_stage_bytes(ctx->Proj.rows, ctx->Proj.width) +
                 ell_stage_bytes(ctx->Ph.rows, ctx->Ph.width) + ell_stage_bytes(ctx->Lf.rows, ctx->Lf.width) +
                 pair_stage_bytes(Nh, ctx->qp_width);
    // measured slower than L1-cached table reads on B200 (copying ~30 KB per 10-element CTA costs more than it
    // saves: quad N=3 rhs 4.9 -> 5.2 ms): off unless ESDG_STAGE_TABLES=1
    ctx->stage_tables = tbl <= 48 * 1024 && getenv("ESDG_STAGE_TABLES") != nullptr;
    if (!ctx->stage_tables) tbl = 0;
    int EB = std::max(1, ESDG_GENERIC_MAX_THREADS / Nh);
    if (d.collocated) {
        // two full rounds of items per phase for the CTA's threads: same residency (two 320-thread CTAs per SM by
        // registers) with half the barriers and per-CTA set-up per element -- quads N=3, 1024^2: 4.78 ms (EB=10)
        // -> 3.85 ms (EB=20) per rhs launch, 16 / 24 / 32 are slower (profiles/r1_generic_launch_shapes.txt).
        // The symmetric schedule needs one thread per row, so a forced schedule keeps one round.
        if (ctx->sym_rounds == 0) EB *= 2;
    } else {
        EB = std::min(EB, 4);                            // modal: see finalize_cns
    }
    if (const char* ev = getenv("ESDG_GENERIC_EB")) EB = std::max(1, atoi(ev));   // launch-shape experiments
    while (EB > 1 && generic_smem_bytes(DIM, EB, d.Ns, d.Nq, d.Nfq, Nh, d.collocated) + tbl > (size_t)maxsm / 2) --EB;
    ctx->EB = EB;
    ctx->table_off = (unsigned)generic_smem_bytes(DIM, EB, d.Ns, d.Nq, d.Nfq, Nh, d.collocated);
    ctx->smem_bytes = ctx->table_off + tbl;
    if (ctx->smem_bytes > (size_t)maxsm) return fail(ctx, ESDG_EINVAL, "esdg_finalize: element too large for shared memory");
    ctx->threads = std::min(ESDG_GENERIC_MAX_THREADS, ((EB * Nh + 31) / 32) * 32);
    if (const char* tv = getenv("ESDG_GENERIC_THREADS")) ctx->threads = std::max(32, (atoi(tv) / 32) * 32);   // launch-shape experiments
    ctx->smem_generic_traces = generic_traces_smem_bytes(DIM, EB, d.Ns, d.Nq, d.Nfq);
    if (DIM == 2) {
        CK(cudaFuncSetAttribute(k_generic_traces<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_generic_traces));
        CK(cudaFuncSetAttribute(k_generic_rhs<2, ESDG_GENERIC_COLL_REGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_bytes));
        CK(cudaFuncSetAttribute(k_generic_rhs<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_bytes));
    } else {
        CK(cudaFuncSetAttribute(k_generic_traces<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_generic_traces));
        CK(cudaFuncSetAttribute(k_generic_rhs<3, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_bytes));
    }
    // ---- state -------------------------------------------------------------------
    const size_t nstate = (size_t)NF * d.Ns * d.K;
    if ((rc = dev_alloc(ctx, &ctx->Q, nstate))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->res, nstate))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->rhs, nstate))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->Qin, nstate))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->traces, (size_t)NF * ctx->trace_stride))) return rc;
    ctx->n_partials = (int)std::max<long long>(grid_blocks(ctx), d.K + 4);   // >= blocks of either path, any split
    if ((rc = dev_alloc(ctx, &ctx->partials, (size_t)ctx->n_partials))) return rc;
    CK(cudaMemsetAsync(ctx->partials, 0, (size_t)ctx->n_partials * sizeof(double), ctx->stream));
    if ((rc = dev_alloc(ctx, &ctx->rhstest_dev, 2))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->bad_flag, 1))) return rc;
    CK(cudaMemsetAsync(ctx->bad_flag, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(ctx->res, 0, nstate * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->traces, 0, (size_t)NF * ctx->trace_stride * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->rhstest_dev, 0, 2 * sizeof(double), ctx->stream));
    if (ctx->cns.on) {
        if ((rc = finalize_cns(ctx))) return rc;
        if ((rc = cnsq_try_create(ctx))) return rc;
    } else if (!ctx->cns.wall.empty() || !ctx->cns.lid.empty() || !ctx->cns.inflow.empty() || !ctx->cns.outflow.empty()) {
        return fail(ctx, ESDG_EINVAL, "esdg_finalize: wall BCs need the CNS path (call esdg_set_viscosity; mu = 0 gives Euler with walls)");
    }
    if (!ctx->cns.on && (rc = hexfast_try_create(ctx))) return rc;
    if (!ctx->cns.on && (rc = quadfast_try_create(ctx))) return rc;
    if (ctx->hex) {
        bool diag = true;                                  // a structured box is axis-aligned by construction
        for (int64_t e = 0; e < d.K && diag && !ctx->structured; ++e)
            for (int k = 0; k < 9; ++k)
                if (k % 4 != 0 && geo[(size_t)e * ctx->geo_stride + k] != 0.0) { diag = false; break; }
        ctx->hex->diag = diag && d.affine && !getenv("ESDG_NO_DIAG");
    }
    ctx->finalized = true;
    return ESDG_OK;
}

static int require_ready(esdg_ctx* ctx, const char* who) {
    if (!ctx) return ESDG_EINVAL;
    if (!ctx->finalized) return fail(ctx, ESDG_EINVAL, std::string(who) + ": esdg_finalize has not been called");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, ESDG_ECUDA, cudaGetErrorString(e));
    return ESDG_OK;
}

int esdg_upload_state(esdg_ctx* ctx, const double* const* Q) {
    int rc = require_ready(ctx, "esdg_upload_state");
    if (rc) return rc;
    if (!Q) return fail(ctx, ESDG_EINVAL, "esdg_upload_state: Q is NULL");
    const size_t n = (size_t)ctx->d.Ns * ctx->d.K;
    const bool nodal = ctx->quad && ctx->quad->cns;      // state kept in the Gauss-nodal basis (kernels_cnsq.cuh)
    double* dst = nodal ? ctx->Qin : ctx->Q;
    for (int f = 0; f < ctx->NF; ++f)
        CK(cudaMemcpyAsync(dst + f * n, Q[f], n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (nodal && (rc = cnsq_basis(ctx, true, ctx->Qin, ctx->Q))) return rc;
    CK(cudaMemsetAsync(ctx->res, 0, ctx->NF * n * sizeof(double), ctx->stream));
    ctx->traces_valid = false;
    ctx->dopri.ready = false;
    return ESDG_OK;
}

int esdg_download_state(esdg_ctx* ctx, double* const* Q) {
    int rc = require_ready(ctx, "esdg_download_state");
    if (rc) return rc;
    if (!Q) return fail(ctx, ESDG_EINVAL, "esdg_download_state: Q is NULL");
    const size_t n = (size_t)ctx->d.Ns * ctx->d.K;
    const double* src = ctx->Q;
    if (ctx->quad && ctx->quad->cns) {
        if ((rc = cnsq_basis(ctx, false, ctx->Q, ctx->Qin))) return rc;
        src = ctx->Qin;
    }
    for (int f = 0; f < ctx->NF; ++f)
        CK(cudaMemcpyAsync(Q[f], src + f * n, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return check_bad(ctx);
}

int esdg_rhs(esdg_ctx* ctx, const double* const* Q, double* const* rhs_out, int compute_rhstest, double* rhstest) {
    int rc = require_ready(ctx, "esdg_rhs");
    if (rc) return rc;
    if (!rhs_out) return fail(ctx, ESDG_EINVAL, "esdg_rhs: rhs_out is NULL");
    if (ctx->d.n_halo != 0)
        return fail(ctx, ESDG_EINVAL, "esdg_rhs: context has halo points; drive it with esdg_compute_traces / esdg_stage_update");
    const size_t n = (size_t)ctx->d.Ns * ctx->d.K;
    const double* src = ctx->Q;
    const bool nodal = ctx->quad && ctx->quad->cns;
    if (Q) {
        double* stage = nodal ? ctx->rhs : ctx->Qin;
        for (int f = 0; f < ctx->NF; ++f)
            CK(cudaMemcpyAsync(stage + f * n, Q[f], n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (nodal && (rc = cnsq_basis(ctx, true, ctx->rhs, ctx->Qin))) return rc;
        src = ctx->Qin;
    }
    if (!(src == ctx->Q && ctx->traces_valid))
        if ((rc = do_traces(ctx, src))) return rc;
    if ((rc = do_stage(ctx, src, 0, 0, 0, 0, compute_rhstest, ESDG_PART_ALL))) return rc;
    ctx->traces_valid = (src == ctx->Q);
    const double* out = ctx->rhs;
    if (nodal) {                                   // the stage is done with its source: Qin is free again
        if ((rc = cnsq_basis(ctx, false, ctx->rhs, ctx->Qin))) return rc;
        out = ctx->Qin;
    }
    for (int f = 0; f < ctx->NF; ++f)
        CK(cudaMemcpyAsync(rhs_out[f], out + f * n, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    double t[2] = {0, 0};
    if (compute_rhstest) CK(cudaMemcpyAsync(t, ctx->rhstest_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = check_bad(ctx))) return rc;
    if (rhstest) {
        rhstest[0] = t[0];
        if (ctx->cns.on) rhstest[1] = t[1];              // rhstest_visc of rhsRK! (dg2D_CNS_cavity_optimized.jl:960-971)
    }
    return ESDG_OK;
}

static int comm_lsrk45(esdg_ctx* ctx, double dt, int nsteps, double* rhstest_last);

int esdg_lsrk45_step(esdg_ctx* ctx, double dt, int nsteps, double* rhstest_last) {
    int rc = require_ready(ctx, "esdg_lsrk45_step");
    if (rc) return rc;
    if (ctx->d.n_halo != 0 || ctx->comm) {
        if (!ctx->comm)
            return fail(ctx, ESDG_EINVAL, "esdg_lsrk45_step: context has halo points; attach a communicator (esdg_comm_init) or drive "
                                          "stages with esdg_stage_update");
        return comm_lsrk45(ctx, dt, nsteps, rhstest_last);
    }
    for (int s = 0; s < nsteps; ++s)
        for (int k = 0; k < 5; ++k) {
            const int want = (rhstest_last != nullptr) && (s == nsteps - 1) && (k == 4);  // dg2D_euler_quad.jl:202
            if (!ctx->traces_valid) {
                if ((rc = do_traces(ctx, ctx->Q))) return rc;
                ctx->traces_valid = true;
            }
            if ((rc = do_stage(ctx, ctx->Q, 1, RK4A[k], RK4B[k], dt, want, ESDG_PART_ALL))) return rc;
        }
    if (rhstest_last) {
        CK(cudaMemcpyAsync(rhstest_last, ctx->rhstest_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        return check_bad(ctx);
    }
    return ESDG_OK;
}

// ---- pipelined host step -----------------------------------------------------------------------------
// upload -> nsteps x 5 stages -> download as a WAVEFRONT over element chunks, so that the host->device and the
// device->host copies overlap the stages AND each other (PCIe is full duplex).  Stage s of a chunk needs stage s-1 of the
// chunks its mapP touches (their new traces; the same condition protects the double-buffered trace arrays against
// overwriting), so after 5 stages a chunk depends on a cone of +-4 neighbour layers only: chunks are uploaded outward
// from chunk 0, every (chunk, stage) task is issued to the one compute stream as soon as its predecessors have been
// issued, and a chunk's device->host copy starts behind its last stage while later chunks are still uploading.
static int build_chunks(esdg_ctx* ctx) {
    const esdg_desc& d = ctx->d;
    const long long K = d.K;
    if (ctx->structured && ctx->mapP_host.empty()) {              // device-generated maps: read them back once (1-based)
        std::vector<int> mp((size_t)d.Nfq * K);
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMemcpy(mp.data(), ctx->mapP, mp.size() * sizeof(int), cudaMemcpyDeviceToHost));
        ctx->mapP_host.resize(mp.size());
        for (size_t i = 0; i < mp.size(); ++i) ctx->mapP_host[i] = (int64_t)mp[i] + 1;
    }
    int nchunk;
    std::function<int(long long)> chunk_of;
    const char* ev = getenv("ESDG_HOST_CHUNKS");
    if (ctx->structured && ctx->skzn == ctx->sK[2] && ctx->sK[1] >= 8 && !(ev && atoi(ev) < 0)) {
        // structured slab: chunks along y (rows of Kx elements in every z-layer).  After five stages a chunk depends on
        // +-4 chunks whatever their size, so what shortens the dependency cone is MANY chunks: up to 64 of them, each a
        // strided element range (Kz segments of rows * Kx elements) that one launch / one 2-D copy per field covers
        const long long Kx = ctx->sK[0], Ky = ctx->sK[1];
        nchunk = (int)std::min<long long>(64, Ky / 2);
        if (ev && atoi(ev) > 0) nchunk = std::max(1, std::min<int>({64, atoi(ev), (int)Ky}));
        const long long rows = (Ky + nchunk - 1) / nchunk;
        ctx->chunk_begin.clear();
        for (long long j = 0; j < Ky; j += rows) ctx->chunk_begin.push_back(j * Kx);
        ctx->chunk_begin.push_back(Ky * Kx);
        ctx->chunk_seg_stride = Kx * Ky;
        ctx->chunk_nseg = ctx->skzn;
        nchunk = (int)ctx->chunk_begin.size() - 1;
        chunk_of = [=](long long e) { return (int)(((e / Kx) % Ky) / rows); };
    } else {
        const int EB = ctx->hex ? 6 * ctx->hex->EB_ws : 1;            // whole batches of either stage kernel per chunk
        nchunk = (int)std::min<long long>(16, K / std::max<long long>(4096, EB));
        if (ev && atoi(ev) != 0) nchunk = std::max(1, std::min(64, std::abs(atoi(ev))));   // experiments
        if (nchunk < 2) nchunk = 1;
        const long long per = ((K + nchunk - 1) / nchunk + EB - 1) / EB * EB;
        ctx->chunk_begin.clear();
        for (long long b = 0; b < K; b += per) ctx->chunk_begin.push_back(b);
        ctx->chunk_begin.push_back(K);
        ctx->chunk_seg_stride = 0;
        ctx->chunk_nseg = 1;
        nchunk = (int)ctx->chunk_begin.size() - 1;
        chunk_of = [=](long long e) {
            int c = (int)(e / per);
            return c < nchunk ? c : nchunk - 1;
        };
    }
    ctx->chunk_dep.assign(nchunk, 0ull);
    for (long long e = 0; e < K; ++e) {
        const int c = chunk_of(e);
        unsigned long long m = 1ull << c;
        for (int k = 0; k < d.Nfq; ++k) m |= 1ull << chunk_of((ctx->mapP_host[(size_t)(e * d.Nfq + k)] - 1) / d.Nfq);
        ctx->chunk_dep[c] |= m;
    }
    CK(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    ctx->chunk_up.resize(nchunk);
    ctx->chunk_done.resize(nchunk);
    for (int c = 0; c < nchunk; ++c) {
        CK(cudaEventCreateWithFlags(&ctx->chunk_up[c], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->chunk_done[c], cudaEventDisableTiming));
    }
    return ESDG_OK;
}

int esdg_lsrk45_host(esdg_ctx* ctx, const double* const* Q_in, double* const* Q_out, double dt, int nsteps) {
    int rc = require_ready(ctx, "esdg_lsrk45_host");
    if (rc) return rc;
    if (!Q_in || !Q_out || nsteps < 1) return fail(ctx, ESDG_EINVAL, "esdg_lsrk45_host: bad argument");
    if (ctx->d.n_halo != 0) return fail(ctx, ESDG_EINVAL, "esdg_lsrk45_host: single-GPU contexts only");
    if (!ctx->hex || ctx->cns.on) {                               // other paths: plain upload / steps / download
        if ((rc = esdg_upload_state(ctx, Q_in))) return rc;
        if ((rc = esdg_lsrk45_step(ctx, dt, nsteps, nullptr))) return rc;
        return esdg_download_state(ctx, Q_out);
    }
    if (ctx->chunk_begin.empty() && (rc = build_chunks(ctx))) return rc;
    HexFast* h = ctx->hex;
    const int nchunk = (int)ctx->chunk_begin.size() - 1;
    const size_t n = (size_t)ctx->d.Ns * ctx->d.K;
    const int Ns = ctx->d.Ns;
    const int nstage = 5 * nsteps;
    ctx->dopri.ready = false;
    // order the copy stream after everything already queued on the main stream
    cudaEvent_t fence = ctx->chunk_done[0];
    CK(cudaEventRecord(fence, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->h2d_stream, fence, 0));
    CK(cudaStreamWaitEvent(ctx->d2h_stream, fence, 0));
    CK(cudaMemsetAsync(ctx->res, 0, ctx->NF * n * sizeof(double), ctx->stream));
    h->prod ^= 1;                                                 // fresh trace / log buffers for the uploaded state
    const int bufU = h->prod;                                     // buffer of the traces of the uploaded state
    // upload order: breadth first from chunk 0 over the chunk graph (periodic slabs: 0, 1, C-1, 2, C-2, ...)
    std::vector<int> order;
    {
        std::vector<char> seen(nchunk, 0);
        order.push_back(0);
        seen[0] = 1;
        for (size_t i = 0; i < order.size() || (int)order.size() < nchunk; ++i) {
            if (i == order.size()) {                              // disconnected remainder
                for (int c = 0; c < nchunk; ++c)
                    if (!seen[c]) { order.push_back(c); seen[c] = 1; break; }
            }
            const int c = order[i];
            for (int q = 0; q < nchunk; ++q)
                if (!seen[q] && (ctx->chunk_dep[c] >> q & 1ull)) { order.push_back(q); seen[q] = 1; }
        }
    }
    std::vector<int> done(nchunk, 0);     // stages issued per chunk
    std::vector<char> up(nchunk, 0);      // uploaded and traces issued
    auto ready = [&](int c) {
        const int st = done[c];
        if (!up[c] || st >= nstage) return false;
        for (int q = 0; q < nchunk; ++q)
            if ((ctx->chunk_dep[c] >> q & 1ull) && (st == 0 ? !up[q] : done[q] < st)) return false;
        return true;
    };
    for (int idx = 0; idx < nchunk; ++idx) {
        const int u = order[idx];
        const long long b = ctx->chunk_begin[u], len = ctx->chunk_begin[u + 1] - b;
        const long long nseg = ctx->chunk_nseg, sst = ctx->chunk_seg_stride;
        for (int f = 0; f < ctx->NF; ++f) {
            if (nseg == 1)
                CK(cudaMemcpyAsync(ctx->Q + f * n + b * Ns, Q_in[f] + b * Ns, (size_t)len * Ns * sizeof(double),
                                   cudaMemcpyHostToDevice, ctx->h2d_stream));
            else
                CK(cudaMemcpy2DAsync(ctx->Q + f * n + b * Ns, (size_t)sst * Ns * sizeof(double), Q_in[f] + b * Ns,
                                     (size_t)sst * Ns * sizeof(double), (size_t)len * Ns * sizeof(double), (size_t)nseg,
                                     cudaMemcpyHostToDevice, ctx->h2d_stream));
        }
        CK(cudaEventRecord(ctx->chunk_up[u], ctx->h2d_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->chunk_up[u], 0));
        h->prod = bufU;
        h->cons = bufU ^ 1;
        if ((rc = hexfast_traces_range(ctx, ctx->Q, b, len * nseg, nseg == 1 ? 0 : len, sst))) return rc;
        up[u] = 1;
        // issue every task that has become ready, deepest stage first (ties: upload order): chunks near chunk 0 finish
        // all their stages -- and start their download -- while the outer chunks are still on the bus
        for (;;) {
            int best = -1;
            for (int k = 0; k < nchunk; ++k) {
                const int c = order[k];
                if (ready(c) && (best < 0 || done[c] > done[best])) best = c;
            }
            if (best < 0) break;
            const int c = best, st = done[c];
            const long long cb = ctx->chunk_begin[c], cl = ctx->chunk_begin[c + 1] - cb;
            h->cons = bufU ^ (st & 1);                            // stage st reads the traces stage st-1 wrote
            h->prod = h->cons ^ 1;
            if ((rc = hexfast_stage_range(ctx, ctx->Q, 1, RK4A[st % 5], RK4B[st % 5], dt, cb, cl * nseg, nseg == 1 ? 0 : cl, sst))) return rc;
            done[c] = st + 1;
            if (done[c] == nstage) {                              // finished: its download goes behind it
                CK(cudaEventRecord(ctx->chunk_done[c], ctx->stream));
                CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->chunk_done[c], 0));
                for (int f = 0; f < ctx->NF; ++f) {
                    if (nseg == 1)
                        CK(cudaMemcpyAsync(Q_out[f] + cb * Ns, ctx->Q + f * n + cb * Ns, (size_t)cl * Ns * sizeof(double),
                                           cudaMemcpyDeviceToHost, ctx->d2h_stream));
                    else
                        CK(cudaMemcpy2DAsync(Q_out[f] + cb * Ns, (size_t)sst * Ns * sizeof(double), ctx->Q + f * n + cb * Ns,
                                             (size_t)sst * Ns * sizeof(double), (size_t)cl * Ns * sizeof(double), (size_t)nseg,
                                             cudaMemcpyDeviceToHost, ctx->d2h_stream));
                }
            }
        }
    }
    for (int c = 0; c < nchunk; ++c)
        if (done[c] != nstage) return fail(ctx, ESDG_EINVAL, "esdg_lsrk45_host: internal scheduling error");
    h->prod = bufU ^ (nstage & 1);                                // traces of the final state
    h->cons = h->prod ^ 1;
    ctx->traces_valid = true;
    CK(cudaStreamSynchronize(ctx->d2h_stream));
    return check_bad(ctx);
}

int esdg_sync(esdg_ctx* ctx) {
    int rc = require_ready(ctx, "esdg_sync");
    if (rc) return rc;
    return check_bad(ctx);
}

// ---- multi-GPU plumbing ----------------------------------------------------------
static double* cur_traces(esdg_ctx* ctx) {
    return ctx->hex ? ctx->hex->trace[ctx->hex->prod] : (ctx->quad ? ctx->quad->trace[ctx->quad->prod] : ctx->traces);
}
static int cur_planes(const esdg_ctx* ctx) { return ctx->hex ? ESDG_HEX_TRACE_PLANES : (ctx->quad ? ESDG_QUAD_TRACE_PLANES : ctx->NF); }

int esdg_trace_planes(esdg_ctx* ctx) { return ctx ? cur_planes(ctx) : 0; }

const char* esdg_kernel_path(const esdg_ctx* ctx) {
    if (!ctx || !ctx->finalized) return "";
    if (ctx->cns.on) return ctx->quad ? "cns_quad_line" : "cns";
    if (ctx->hex) return ctx->hex->curved ? "hex_line_curved" : (ctx->hex->diag ? "hex_line_diag" : "hex_line");
    if (ctx->quad) return "quad_line";
    return "generic";
}

int esdg_set_send_list(esdg_ctx* ctx, const int64_t* idx, int64_t n) {
    int rc = require_ready(ctx, "esdg_set_send_list");
    if (rc) return rc;
    std::vector<int> h((size_t)n);
    std::vector<char> is_bnd((size_t)ctx->d.K, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t v = idx[i] - 1;
        if (v < 0 || v >= (int64_t)ctx->d.Nfq * ctx->d.K) return fail(ctx, ESDG_EINVAL, "esdg_set_send_list: index out of range");
        h[(size_t)i] = (int)v;
        is_bnd[(size_t)(v / ctx->d.Nfq)] = 1;
    }
    ctx->n_send = n;
    // elements whose traces travel: a leading and a trailing run (z-slab partition); anything else => no split
    long long lo = 0, hi = 0, cnt = 0;
    for (char c : is_bnd) cnt += c;
    while (lo < ctx->d.K && is_bnd[(size_t)lo]) ++lo;
    while (hi < ctx->d.K - lo && is_bnd[(size_t)(ctx->d.K - 1 - hi)]) ++hi;
    if (lo + hi != cnt) { lo = ctx->d.K; hi = 0; }   // scattered boundary set: treat every element as boundary
    ctx->bnd_lo = lo;
    ctx->bnd_hi = hi;
    if (ctx->send_idx) {                                   // a second call replaces the list
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->allocs.erase(std::remove(ctx->allocs.begin(), ctx->allocs.end(), (void*)ctx->send_idx), ctx->allocs.end());
        cudaFree(ctx->send_idx);
        ctx->send_idx = nullptr;
    }
    return dev_upload(ctx, &ctx->send_idx, h);
}

__global__ void k_pack_send(const double* __restrict__ traces, long long stride, const int* __restrict__ idx,
                            long long n, int nf, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = idx[i];
    for (int f = 0; f < nf; ++f) out[f * n + i] = traces[f * stride + j];
}

int esdg_compute_traces(esdg_ctx* ctx) {
    int rc = require_ready(ctx, "esdg_compute_traces");
    if (rc) return rc;
    if ((rc = do_traces(ctx, ctx->Q))) return rc;
    ctx->traces_valid = true;
    return ESDG_OK;
}

int esdg_pack_send(esdg_ctx* ctx, void* send_dev) {
    int rc = require_ready(ctx, "esdg_pack_send");
    if (rc) return rc;
    if (ctx->n_send == 0) return ESDG_OK;
    tick(ctx, "pack_send");
    k_pack_send<<<(unsigned)((ctx->n_send + 255) / 256), 256, 0, ctx->stream>>>(cur_traces(ctx), ctx->trace_stride, ctx->send_idx,
                                                                               ctx->n_send, cur_planes(ctx), (double*)send_dev);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

int esdg_halo_ptr(esdg_ctx* ctx, void** halo_dev, int64_t* field_stride) {
    int rc = require_ready(ctx, "esdg_halo_ptr");
    if (rc) return rc;
    if (halo_dev) *halo_dev = cur_traces(ctx) + (long long)ctx->d.Nfq * ctx->d.K;
    if (field_stride) *field_stride = ctx->trace_stride;
    return ESDG_OK;
}

int esdg_unpack_halo(esdg_ctx* ctx, const void* recv_dev) {
    int rc = require_ready(ctx, "esdg_unpack_halo");
    if (rc) return rc;
    if (ctx->d.n_halo == 0) return ESDG_OK;
    double* dst = cur_traces(ctx) + (long long)ctx->d.Nfq * ctx->d.K;
    CK(cudaMemcpy2DAsync(dst, ctx->trace_stride * sizeof(double), recv_dev, ctx->d.n_halo * sizeof(double),
                         ctx->d.n_halo * sizeof(double), cur_planes(ctx), cudaMemcpyDeviceToDevice, ctx->stream));
    return ESDG_OK;
}

// ---- CNS across GPUs: the second exchange of a stage (sigma traces) ----
int esdg_cns_sigma_pass(esdg_ctx* ctx, int compute_rhstest) {
    int rc = require_ready(ctx, "esdg_cns_sigma_pass");
    if (rc) return rc;
    if (!ctx->cns.on) return fail(ctx, ESDG_EINVAL, "esdg_cns_sigma_pass: not a CNS context");
    if (!ctx->traces_valid) return fail(ctx, ESDG_EINVAL, "esdg_cns_sigma_pass: call esdg_compute_traces (and exchange halos) after uploading a state");
    if ((rc = cns_sigma_pass(ctx, ctx->Q, compute_rhstest))) return rc;
    ctx->cns.sigma_valid = true;
    return ESDG_OK;
}

int esdg_pack_send_sigma(esdg_ctx* ctx, void* send_dev) {
    int rc = require_ready(ctx, "esdg_pack_send_sigma");
    if (rc) return rc;
    if (!ctx->cns.on) return fail(ctx, ESDG_EINVAL, "esdg_pack_send_sigma: not a CNS context");
    if (ctx->n_send == 0) return ESDG_OK;
    tick(ctx, "pack_send");
    k_pack_send<<<(unsigned)((ctx->n_send + 255) / 256), 256, 0, ctx->stream>>>(ctx->cns.sig_tr, ctx->trace_stride, ctx->send_idx,
                                                                               ctx->n_send, 6, (double*)send_dev);
    tock(ctx);
    CK(cudaGetLastError());
    return ESDG_OK;
}

int esdg_unpack_halo_sigma(esdg_ctx* ctx, const void* recv_dev) {
    int rc = require_ready(ctx, "esdg_unpack_halo_sigma");
    if (rc) return rc;
    if (!ctx->cns.on) return fail(ctx, ESDG_EINVAL, "esdg_unpack_halo_sigma: not a CNS context");
    if (ctx->d.n_halo == 0) return ESDG_OK;
    double* dst = ctx->cns.sig_tr + (long long)ctx->d.Nfq * ctx->d.K;
    CK(cudaMemcpy2DAsync(dst, ctx->trace_stride * sizeof(double), recv_dev, ctx->d.n_halo * sizeof(double),
                         ctx->d.n_halo * sizeof(double), 6, cudaMemcpyDeviceToDevice, ctx->stream));
    return ESDG_OK;
}

int esdg_stage_update(esdg_ctx* ctx, int stage, double dt, int compute_rhstest, int part) {
    int rc = require_ready(ctx, "esdg_stage_update");
    if (rc) return rc;
    if (stage < 0 || stage > 4) return fail(ctx, ESDG_EINVAL, "esdg_stage_update: stage must be 0..4");
    if (part < 0 || part > 2) return fail(ctx, ESDG_EINVAL, "esdg_stage_update: part must be 0 (all), 1 (boundary) or 2 (interior)");
    if (!ctx->traces_valid) return fail(ctx, ESDG_EINVAL, "esdg_stage_update: call esdg_compute_traces (and exchange halos) after uploading a state");
    return do_stage(ctx, ctx->Q, 1, RK4A[stage], RK4B[stage], dt, compute_rhstest, part);
}

int esdg_read_rhstest(esdg_ctx* ctx, double* rhstest) {
    int rc = require_ready(ctx, "esdg_read_rhstest");
    if (rc) return rc;
    CK(cudaMemcpyAsync(rhstest, ctx->rhstest_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return check_bad(ctx);
}

// ---- DOPRI45 adaptive driver (dg2D_CNS_cavity_optimized.jl:919-934, :974-1053) --------------------
namespace {
const double DP_A[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.2, 0, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0},
    {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};          // :920-927
const double DP_E[7] = {71.0 / 57600.0, 0.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0,
                        22.0 / 525.0, -1.0 / 40.0};                                              // :931

struct DopriArgs {
    const double* K[7];
    double c[7];
    int ns;
};

// Qtmp = Q + dt * sum_s a[s] K_s   (k accumulated in the reference's order s = 1..INTRK-1, :1005-1010)
__global__ void k_dopri_combine(const double* __restrict__ Q, double* __restrict__ Qtmp, DopriArgs a, double dt, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double k = 0.0;
    for (int s = 0; s < a.ns; ++s) k = k + a.c[s] * a.K[s][i];
    Qtmp[i] = Q[i] + dt * k;
}

// partial sums of (|sum_s E[s] K_s| / (errTol (1 + |Q|)))^2   (:1015-1023), one partial per CTA, fixed order
__global__ void k_dopri_err(const double* __restrict__ Q, DopriArgs a, double err_tol, long long n, double* __restrict__ partial) {
    __shared__ double sh[256];
    double t = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        double e = 0.0;
        for (int s = 0; s < 7; ++s) e = e + a.c[s] * a.K[s][i];
        const double sc = fabs(e) / (err_tol * (1.0 + fabs(Q[i])));
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

int dopri_rhs(esdg_ctx* ctx, const double* src, double* out, int want) {
    int rc;
    ctx->rhs_target = out;
    rc = do_traces(ctx, src);
    if (!rc) rc = do_stage(ctx, src, 0, 0, 0, 0, want, ESDG_PART_ALL);
    ctx->rhs_target = nullptr;
    ctx->traces_valid = false;
    return rc;
}
}  // namespace

int esdg_dopri45_init(esdg_ctx* ctx) {
    int rc = require_ready(ctx, "esdg_dopri45_init");
    if (rc) return rc;
    if (ctx->d.n_halo != 0) return fail(ctx, ESDG_EINVAL, "esdg_dopri45_init: single-GPU contexts only");
    DopriState& D = ctx->dopri;
    const size_t n = (size_t)ctx->NF * ctx->d.Ns * ctx->d.K;
    if (!D.K[0]) {
        for (int s = 0; s < 7; ++s)
            if ((rc = dev_alloc(ctx, &D.K[s], n))) return rc;
        if ((rc = dev_alloc(ctx, &D.Qtmp, n))) return rc;
        D.n_err_blocks = (int)std::min<size_t>(1024, (n + 255) / 256);
        if ((rc = dev_alloc(ctx, &D.err_partial, (size_t)D.n_err_blocks))) return rc;
        if ((rc = dev_alloc(ctx, &D.err_dev, 2))) return rc;
    }
    D.prev_err = 0.0;
    D.attempts = 0;
    if ((rc = dopri_rhs(ctx, ctx->Q, D.K[0], 0))) return rc;       // FSAL start (:996-997)
    D.ready = true;
    return check_bad(ctx);
}

int esdg_dopri45_step(esdg_ctx* ctx, double* t, double* dt, double dt0, double err_tol, int* accepted, double* err_est,
                      double* rhstest) {
    int rc = require_ready(ctx, "esdg_dopri45_step");
    if (rc) return rc;
    DopriState& D = ctx->dopri;
    if (!D.ready) return fail(ctx, ESDG_EINVAL, "esdg_dopri45_step: call esdg_dopri45_init after uploading the state");
    if (!t || !dt || !(*dt > 0.0) || !(err_tol > 0.0)) return fail(ctx, ESDG_EINVAL, "esdg_dopri45_step: bad argument");
    const long long n = (long long)ctx->NF * ctx->d.Ns * ctx->d.K;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    for (int k = 1; k < 7; ++k) {                                    // INTRK = 2:7 (:1004-1013)
        DopriArgs a;
        a.ns = k;
        for (int s = 0; s < 7; ++s) { a.K[s] = D.K[s]; a.c[s] = s < k ? DP_A[k][s] : 0.0; }
        tick(ctx, "dopri_combine");
        k_dopri_combine<<<blocks, 256, 0, ctx->stream>>>(ctx->Q, D.Qtmp, a, *dt, n);
        tock(ctx);
        CK(cudaGetLastError());
        if ((rc = dopri_rhs(ctx, D.Qtmp, D.K[k], k == 6))) return rc;
    }
    DopriArgs e;
    e.ns = 7;
    for (int s = 0; s < 7; ++s) { e.K[s] = D.K[s]; e.c[s] = DP_E[s]; }
    if (ctx->quad && ctx->quad->cns) {               // the script's norm is taken over ITS coefficients
        if ((rc = cnsq_dopri_err(ctx, e.K, e.c, err_tol, D.n_err_blocks, D.err_partial))) return rc;
    } else {
        tick(ctx, "dopri_err");
        k_dopri_err<<<D.n_err_blocks, 256, 0, ctx->stream>>>(ctx->Q, e, err_tol, n, D.err_partial);
        tock(ctx);
        CK(cudaGetLastError());
    }
    tick(ctx, "reduce_partials");
    k_reduce_partials<<<1, 256, 0, ctx->stream>>>(D.err_partial, D.n_err_blocks, D.err_dev);
    tock(ctx);
    CK(cudaGetLastError());
    double h[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(h, D.err_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h + 1, ctx->rhstest_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = check_bad(ctx))) return rc;
    const double err = std::sqrt(h[0] / (double)n);                  // length(Q[1])*4 = all entries (:1024)
    const int ok = err < 1.0;
    if (ok) {                                                        // :1025-1029
        CK(cudaMemcpyAsync(ctx->Q, D.Qtmp, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        *t += *dt;
        std::swap(D.K[0], D.K[6]);                                   // FSAL
        ctx->traces_valid = false;
    }
    const double order = 5.0;
    double dtnew = .8 * *dt * std::pow(.9 / err, .4 / (order + 1));  // P controller (:1031)
    if (D.attempts > 0) dtnew *= std::pow(D.prev_err / std::max(1e-14, err), .3 / (order + 1));   // PI (:1032-1034)
    *dt = std::max(std::min(10 * dt0, dtnew), 1e-9);                 // :1035
    D.prev_err = err;
    D.attempts += 1;
    if (accepted) *accepted = ok;
    if (err_est) *err_est = err;
    if (rhstest) { rhstest[0] = h[1]; rhstest[1] = ctx->cns.on ? h[2] : 0.0; }
    return ESDG_OK;
}

int esdg_read_rhstest_visc(esdg_ctx* ctx, double* rhstest_visc) {
    int rc = require_ready(ctx, "esdg_read_rhstest_visc");
    if (rc) return rc;
    if (!ctx->cns.on) return fail(ctx, ESDG_EINVAL, "esdg_read_rhstest_visc: not a CNS context");
    CK(cudaMemcpyAsync(rhstest_visc, ctx->rhstest_dev + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return check_bad(ctx);
}

// ---- introspection ---------------------------------------------------------------
int64_t esdg_kernel_launches(const esdg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int esdg_enable_timing(esdg_ctx* ctx, int on) {
    if (!ctx) return ESDG_EINVAL;
    ctx->timing = on != 0;
    return ESDG_OK;
}

int esdg_kernel_time_ms(esdg_ctx* ctx, const char* kernel, double* total_ms, int64_t* launches) {
    if (!ctx || !kernel) return ESDG_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& t : ctx->timers) {
        float ms = 0;
        cudaEventElapsedTime(&ms, t.a, t.b);
        auto& acc = ctx->time_acc[t.name];
        acc.first += ms;
        acc.second += 1;
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    ctx->timers.clear();
    auto it = ctx->time_acc.find(kernel);
    if (total_ms) *total_ms = it == ctx->time_acc.end() ? 0.0 : it->second.first;
    if (launches) *launches = it == ctx->time_acc.end() ? 0 : it->second.second;
    return ESDG_OK;
}

// ---- multi-GPU inside the boundary: the context owns an NCCL communicator and a communication stream ------------------
// (SURVEY 8b / 8e).  NCCL is resolved at run time (dlopen of libnccl.so.2 -- already in the process when the host is
// PyTorch or NCCL.jl), so the library has no link-time dependency on it and loads on machines without NCCL.
}  // extern "C"

#include <dlfcn.h>

namespace {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId_t { char internal[128]; };
enum { NCCL_DOUBLE = 8, NCCL_SUM = 0 };
struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

const char* load_nccl() {
    if (g_nccl.ok) return nullptr;
    void* h = nullptr;
    const char* names[] = {getenv("ESDG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names)
        if (nm && !h) h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return "libnccl.so.2 not found (set ESDG_NCCL_LIB or load NCCL in the host process first)";
#define ESDG_NCCL_SYM(field, sym)                                         \
    *(void**)(&g_nccl.field) = dlsym(h, sym);                             \
    if (!g_nccl.field) return "libnccl: missing symbol " sym;
    ESDG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    ESDG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    ESDG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    ESDG_NCCL_SYM(GroupStart, "ncclGroupStart")
    ESDG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    ESDG_NCCL_SYM(Send, "ncclSend")
    ESDG_NCCL_SYM(Recv, "ncclRecv")
    ESDG_NCCL_SYM(AllReduce, "ncclAllReduce")
    ESDG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ESDG_NCCL_SYM
    g_nccl.ok = true;
    return nullptr;
}
}  // namespace

struct EsdgComm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    std::vector<int> nb_rank;               // neighbours in send-list order
    std::vector<long long> send_off, send_cnt, recv_off, recv_cnt;   // per neighbour, in trace points
    cudaStream_t stream = nullptr;          // communication stream
    cudaEvent_t packed = nullptr, landed = nullptr;
    double* sendbuf = nullptr;              // [neighbour][plane][count]
    int* send_slot = nullptr;               // send-list entry -> first entry of its neighbour's block
    int* send_cntv = nullptr;               // send-list entry -> size of its neighbour's block
    int max_planes = 0;
    bool pending = false;                   // an exchange is in flight on the communication stream
    bool fresh = true;                      // state uploaded, traces not yet computed / exchanged
};

#define NCK(call)                                                                                        \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != 0) return fail(ctx, ESDG_ENCCL, std::string(#call ": ") + g_nccl.GetErrorString(r_));  \
    } while (0)

static void comm_destroy(esdg_ctx* ctx) {
    EsdgComm* c = ctx->comm;
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
    if (c->packed) cudaEventDestroy(c->packed);
    if (c->landed) cudaEventDestroy(c->landed);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    ctx->comm = nullptr;
}

namespace {
// send-list entry i of neighbour b goes to sendbuf[planes*off_b + f*cnt_b + (i - off_b)]: one contiguous message per
// (neighbour, plane), no staging copy between the gather and ncclSend
__global__ void k_pack_send_nb(const double* __restrict__ traces, long long stride, const int* __restrict__ idx,
                               const int* __restrict__ slot, const int* __restrict__ cntv, long long n, int nf,
                               double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = idx[i], c = cntv[i];
    const long long ob = slot[i], local = i - ob;   // slot[i] = first send-list entry of the neighbour the entry belongs to
    for (int f = 0; f < nf; ++f) out[ob * nf + (long long)f * c + local] = traces[f * stride + j];
}

// start the exchange of `planes` trace planes of `tr` (halo tail after Nfq*K): pack on the compute stream, NCCL
// send / recv on the communication stream behind it
int comm_exchange_start(esdg_ctx* ctx, double* tr, int planes) {
    EsdgComm* c = ctx->comm;
    if (planes > c->max_planes) return fail(ctx, ESDG_EINVAL, "comm: too many trace planes");
    if (ctx->n_send > 0) {
        tick(ctx, "pack_send");
        k_pack_send_nb<<<(unsigned)((ctx->n_send + 255) / 256), 256, 0, ctx->stream>>>(tr, ctx->trace_stride, ctx->send_idx, c->send_slot,
                                                                                      c->send_cntv, ctx->n_send, planes, c->sendbuf);
        tock(ctx);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(c->packed, ctx->stream));
    CK(cudaStreamWaitEvent(c->stream, c->packed, 0));
    const long long tail = (long long)ctx->d.Nfq * ctx->d.K;
    const int nnb = (int)c->nb_rank.size();
    NCK(g_nccl.GroupStart());
    for (int b = 0; b < nnb; ++b)
        for (int f = 0; f < planes; ++f)
            if (c->send_cnt[b] > 0 && c->nb_rank[b] >= 0)
                NCK(g_nccl.Send(c->sendbuf + c->send_off[b] * planes + (long long)f * c->send_cnt[b], (size_t)c->send_cnt[b], NCCL_DOUBLE,
                                c->nb_rank[b], c->comm, c->stream));
    // receives in reverse neighbour order: with two ranks and a periodic direction both messages of a plane travel
    // between the same pair, and the peer's "to its lower neighbour" block is my "from my upper neighbour" block
    for (int b = nnb - 1; b >= 0; --b)
        for (int f = 0; f < planes; ++f)
            if (c->recv_cnt[b] > 0 && c->nb_rank[b] >= 0)
                NCK(g_nccl.Recv(tr + (long long)f * ctx->trace_stride + tail + c->recv_off[b], (size_t)c->recv_cnt[b], NCCL_DOUBLE,
                                c->nb_rank[b], c->comm, c->stream));
    NCK(g_nccl.GroupEnd());
    CK(cudaEventRecord(c->landed, c->stream));
    c->pending = true;
    return ESDG_OK;
}

int comm_exchange_finish(esdg_ctx* ctx) {
    EsdgComm* c = ctx->comm;
    if (c->pending) {
        CK(cudaStreamWaitEvent(ctx->stream, c->landed, 0));
        c->pending = false;
    }
    return ESDG_OK;
}
}  // namespace

// LSRK45 on a partitioned mesh (one context per rank).  Euler, tensor-product hexes: per stage the boundary-layer
// elements first, their new traces packed and sent while the interior elements are updated, the halo received straight
// into the tail of the trace planes the NEXT stage reads.  Other paths: stage, then exchange.  CNS: two exchanges per
// stage (sigma traces, entropy-variable traces).
static int comm_lsrk45(esdg_ctx* ctx, double dt, int nsteps, double* rhstest_last) {
    EsdgComm* c = ctx->comm;
    int rc;
    const int planes = cur_planes(ctx);
    if (c->fresh || !ctx->traces_valid) {
        if ((rc = do_traces(ctx, ctx->Q))) return rc;
        ctx->traces_valid = true;
        if ((rc = comm_exchange_start(ctx, cur_traces(ctx), planes))) return rc;
        if ((rc = comm_exchange_finish(ctx))) return rc;
        c->fresh = false;
    }
    const bool overlap = (ctx->hex != nullptr || (ctx->quad != nullptr && !ctx->quad->cns)) && !getenv("ESDG_COMM_NO_OVERLAP");
    for (int s = 0; s < nsteps; ++s)
        for (int k = 0; k < 5; ++k) {
            const int want = (rhstest_last != nullptr) && (s == nsteps - 1) && (k == 4);
            if (ctx->cns.on) {
                if ((rc = cns_sigma_pass(ctx, ctx->Q, want))) return rc;
                ctx->cns.sigma_valid = true;
                if ((rc = comm_exchange_start(ctx, ctx->cns.sig_tr, 6))) return rc;
                if ((rc = comm_exchange_finish(ctx))) return rc;
                if ((rc = do_stage(ctx, ctx->Q, 1, RK4A[k], RK4B[k], dt, want, ESDG_PART_ALL))) return rc;
                if ((rc = comm_exchange_start(ctx, cur_traces(ctx), planes))) return rc;
                if ((rc = comm_exchange_finish(ctx))) return rc;
            } else if (overlap) {
                if ((rc = do_stage(ctx, ctx->Q, 1, RK4A[k], RK4B[k], dt, want, ESDG_PART_BOUNDARY))) return rc;
                if ((rc = comm_exchange_start(ctx, cur_traces(ctx), planes))) return rc;
                if ((rc = do_stage(ctx, ctx->Q, 1, RK4A[k], RK4B[k], dt, want, ESDG_PART_INTERIOR))) return rc;
                if ((rc = comm_exchange_finish(ctx))) return rc;
            } else {
                if ((rc = do_stage(ctx, ctx->Q, 1, RK4A[k], RK4B[k], dt, want, ESDG_PART_ALL))) return rc;
                if ((rc = comm_exchange_start(ctx, cur_traces(ctx), planes))) return rc;
                if ((rc = comm_exchange_finish(ctx))) return rc;
            }
        }
    if (rhstest_last) {
        // entropy residual of the whole mesh: sum of the ranks' partials (2 doubles: total, viscous)
        NCK(g_nccl.AllReduce(ctx->rhstest_dev, ctx->rhstest_dev, 2, NCCL_DOUBLE, NCCL_SUM, c->comm, ctx->stream));
        CK(cudaMemcpyAsync(rhstest_last, ctx->rhstest_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        return check_bad(ctx);
    }
    return ESDG_OK;
}

extern "C" {

int esdg_comm_unique_id(void* id128) {
    if (!id128) return ESDG_EINVAL;
    if (load_nccl()) return ESDG_ENCCL;
    ncclUniqueId_t id;
    if (g_nccl.GetUniqueId(&id) != 0) return ESDG_ENCCL;
    std::memcpy(id128, &id, sizeof id);
    return ESDG_OK;
}

int esdg_comm_init(esdg_ctx* ctx, const void* id128, int nranks, int rank, int n_neighbours, const int* nb_rank,
                   const int64_t* nb_send, const int64_t* nb_recv) {
    int rc = require_ready(ctx, "esdg_comm_init");
    if (rc) return rc;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks || n_neighbours < 0 || (n_neighbours && (!nb_rank || !nb_send || !nb_recv)))
        return fail(ctx, ESDG_EINVAL, "esdg_comm_init: bad argument");
    if (const char* e = load_nccl()) return fail(ctx, ESDG_ENCCL, std::string("esdg_comm_init: ") + e);
    long long ns = 0, nr = 0;
    for (int b = 0; b < n_neighbours; ++b) {
        if (nb_rank[b] < -1 || nb_rank[b] >= nranks || nb_send[b] < 0 || nb_recv[b] < 0) return fail(ctx, ESDG_EINVAL, "esdg_comm_init: bad neighbour entry");
        ns += nb_send[b];
        nr += nb_recv[b];
    }
    if (ns != ctx->n_send) return fail(ctx, ESDG_EINVAL, "esdg_comm_init: neighbour send counts do not add up to the send list (esdg_set_send_list)");
    if (nr != ctx->d.n_halo) return fail(ctx, ESDG_EINVAL, "esdg_comm_init: neighbour receive counts do not add up to desc.n_halo");
    comm_destroy(ctx);
    EsdgComm* c = new EsdgComm();
    ctx->comm = c;
    c->nranks = nranks;
    c->rank = rank;
    std::vector<int> slot((size_t)ns), cntv((size_t)ns);
    long long so = 0, ro = 0;
    for (int b = 0; b < n_neighbours; ++b) {
        c->nb_rank.push_back(nb_rank[b]);
        c->send_off.push_back(so);
        c->send_cnt.push_back(nb_send[b]);
        c->recv_off.push_back(ro);
        c->recv_cnt.push_back(nb_recv[b]);
        for (long long i = 0; i < nb_send[b]; ++i) { slot[(size_t)(so + i)] = (int)so; cntv[(size_t)(so + i)] = (int)nb_send[b]; }
        so += nb_send[b];
        ro += nb_recv[b];
    }
    c->max_planes = std::max(8, ctx->NF + 2);
    if ((rc = dev_upload(ctx, &c->send_slot, slot))) return rc;
    if ((rc = dev_upload(ctx, &c->send_cntv, cntv))) return rc;
    if ((rc = dev_alloc(ctx, &c->sendbuf, (size_t)std::max<long long>(1, ns) * c->max_planes))) return rc;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->packed, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->landed, cudaEventDisableTiming));
    // The halo messages are small (a few MB per neighbour and stage) and hide behind the interior elements: ONE p2p
    // channel (= one NCCL CTA per direction) keeps NCCL out of the way of the persistent stage kernel, which takes every
    // register of the SMs it runs on (hex_fast.cu leaves comm_sms SMs free for it).  Measured on 2 x B200, 128x128x16
    // elements per GPU: default channels 23.6 ms per step, 1 channel + 1 free SM 22.25 ms (1 GPU: 21.7 ms).
    setenv("NCCL_MAX_P2P_NCHANNELS", "1", 0);
    setenv("NCCL_MIN_P2P_NCHANNELS", "1", 0);
    ncclUniqueId_t id;
    std::memcpy(&id, id128, sizeof id);
    NCK(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    return ESDG_OK;
}

int esdg_comm_destroy(esdg_ctx* ctx) {
    if (!ctx) return ESDG_EINVAL;
    comm_destroy(ctx);
    return ESDG_OK;
}

}  // extern "C"
