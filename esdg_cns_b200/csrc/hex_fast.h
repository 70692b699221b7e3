// Host side of the tensor-product hex path (hex_common.cuh, kernels_hex_line.cuh, kernels_hex_ws.cuh).
#pragma once
#include "esdg_ctx.h"

constexpr int ESDG_HEX_TRACE_PLANES = 8;   // == esdg::HEX_TP (hex_common.cuh)

struct HexFast {
    int n = 0;                 // N + 1
    bool diag = false;         // axis-aligned elements: only rxJ, syJ, tzJ non-zero
    bool curved = false;       // desc.affine = 0: per-point geometric factors, one-shot kernel with line_compute_curved
    double* geoL = nullptr;    // curved: rxJ.. line-major, [e][(n+2)*3][NL]
    double T[64], lm[8], lp[8], w[8];
    double* trace[2] = {nullptr, nullptr};   // 8 planes x trace_stride each
    double* logs[2] = {nullptr, nullptr};    // 2 planes x Nq*K each (log rho, log beta at volume nodes)
    int prod = 0;              // buffer most recently produced (pack/unpack act on it)
    int cons = 0;              // buffer the running stage consumes
    size_t smem_stage = 0, smem_traces = 0, smem_curved = 0;
    int threads = 0, EB = 1;           // one-shot stage kernel launch shape (kernels_hex_line.cuh)
    int threads_tr = 0, EB_tr = 1;     // k_hex_traces launch shape
    int threads_ws = 0, EB_ws = 1;     // warp-specialised persistent stage kernel (kernels_hex_ws.cuh)
    size_t smem_ws = 0;
    int sms = 0;                       // persistent grid size
    int comm_sms = 2;                  // SMs left free for NCCL while the interior part of a split stage runs (ESDG_COMM_SMS)
    int kernel_choice = 0;             // 0: by range size, 1: one-shot, 2: persistent (ESDG_HEX_KERNEL)
};

enum { ESDG_PART_ALL = 0, ESDG_PART_BOUNDARY = 1, ESDG_PART_INTERIOR = 2 };

// Decide whether the caller's operators are the Kronecker products the hex kernels assume;
// on success ctx->hex is set (device buffers allocated).  Returns an esdg_status; a mismatch is
// NOT an error (ctx->hex stays null and the generic kernels are used).
int hexfast_try_create(esdg_ctx* ctx);
void hexfast_destroy(esdg_ctx* ctx);
// traces + log stash of state Q into a fresh buffer (becomes `prod`)
int hexfast_traces(esdg_ctx* ctx, const double* Q);
// element-range versions for the pipelined host step (buffers selected by the caller through h->prod / h->cons)
// (seg_len, seg_stride): strided ranges, see HexParams (0, 0 = contiguous)
int hexfast_traces_range(esdg_ctx* ctx, const double* Q, long long e_begin, long long e_count, long long seg_len = 0, long long seg_stride = 0);
int hexfast_stage_range(esdg_ctx* ctx, const double* Q, int mode, double a, double b, double dt, long long e_begin, long long e_count,
                        long long seg_len = 0, long long seg_stride = 0);
// one rhs evaluation (mode 0) or fused LSRK stage (mode 1) over a part of the elements
int hexfast_stage(esdg_ctx* ctx, const double* Q, int mode, double a, double b, double dt, int want_rhstest, int part);
