// Tensor-product hex stage, WARP-SPECIALISED PERSISTENT kernel (the product path of BASELINE configs 4 / 5).
//
// One CTA per SM walks over batches of EB elements.  Its 16 warps form two teams with different register budgets
// (setmaxnreg): 8 LINE warps (one line problem per thread and batch, ~170 registers: the n+2 nodes and 5n
// accumulators live in registers) and 8 POINT warps (~88 registers) that do everything else -- they issue the loads
// of batch i+1 (own traces as TMA bulk copies, exterior traces gathered through mapP with cp.async), compute its
// volume primitives (A), and finish batch i-1 (E: rhs, LSRK update, entropy variables; F: traces of the new state)
// WHILE the line warps work on batch i.  The line phase carries ~70 % of the FP64 instructions; in the one-shot
// kernel (kernels_hex_line.cuh) the per-point phases, the loads and the block barriers of a CTA leave the FP64
// pipe to the one other resident CTA for two thirds of its lifetime (44 % pipe utilisation, profiles/r1_*).  Here
// the line warps never wait once the pipeline is full, the point phases run beside them, and nothing is
// synchronised block-wide: hand-over is by mbarriers on double-buffered shared arrays.
//
//   full_tp[b]   (point -> line)  traces of the batch in WF[b], P[b]  (1 expect_tx arrival + tx bytes + cp.async arrivals)
//   full_wv[b]   (point -> line)  WV[b], geo[b] written
//   empty_in[b]  (line -> point)  inputs of buffer b consumed
//   full_a3      (line -> point)  line results of the batch in A3
//   empty_a3     (point -> line)  A3 consumed by the update phase
//
// Bitwise identical to the one-shot kernel: same per-item device functions, same summation order.
#pragma once
#include "kernels_hex_line.cuh"

namespace esdg {

// Team sizes and register budgets (setmaxnreg acts on warpgroups: multiples of 128 threads; 256 x LREGS + NPT x PREGS
// must fit the CTA's register pool = (256 + NPT) x the launch count).  N = 4: the line problem needs ~184 registers
// (n+2 nodes, 5n accumulators, three pairs in flight), which leaves room for 4 point warps at 136.  Measured on the
// bench slab (profiles/r2_ws_variants.txt): 8 point warps at 72-88 registers 5.8-7.3 ms, 4 at 104: 7.6, 4 at 120: 4.27,
// 4 at 136: 4.11, 4 at 152 (lines 176): 4.48, 4 at 168 (lines 168): 5.15.  N <= 3: the line problem is smaller and the point team
// is the busier one: 168 / 168 (no register hand-over at all) 2.27 ms per stage on 64^3 elements at N = 3 against 2.36 at 184 / 136;
// 160 / 184: 2.74, 152 / 200: 3.29 (line spills); 8 point warps at 128 / 120: 2.92.
#ifndef ESDG_WS_NPT5
#define ESDG_WS_NPT5 128
#endif
#ifndef ESDG_WS_LREGS5
#define ESDG_WS_LREGS5 184
#endif
#ifndef ESDG_WS_PREGS5
#define ESDG_WS_PREGS5 136
#endif
#ifndef ESDG_WS_NPT4
#define ESDG_WS_NPT4 128
#endif
#ifndef ESDG_WS_LREGS4
#define ESDG_WS_LREGS4 168
#endif
#ifndef ESDG_WS_PREGS4
#define ESDG_WS_PREGS4 168
#endif

template <int n>
struct WsDims {
    using G = HexGeom<n>;
    static constexpr int NLT = 256, NPT = n == 5 ? ESDG_WS_NPT5 : ESDG_WS_NPT4, THREADS = NLT + NPT;
    static constexpr int LREGS = n == 5 ? ESDG_WS_LREGS5 : ESDG_WS_LREGS4, PREGS = n == 5 ? ESDG_WS_PREGS5 : ESDG_WS_PREGS4;
    static constexpr int LAUNCH_REGS = (65536 / THREADS) / 8 * 8;   // what every thread holds at launch; the teams move up / down from it
    static constexpr int EB = NLT / G::NL;                  // n = 5: 3 elements (225 lines), 4: 5, 3: 9, 2: 21
    static constexpr int EN = EB * G::Nq, EF = EB * G::Nfq;
    static constexpr int ENP = wv_stride<n, EB>();                                 // plane stride of WV (even)
    static constexpr int WV_DOUBLES = 7 * ENP;
    static constexpr int IN_DOUBLES = WV_DOUBLES + HEX_TP * EF;                   // WV, WF of one buffer (even)
    static constexpr int GEO_DOUBLES = EB * HEX_SGEO;                              // geometry records: ring of three batches
    static constexpr int SMEM_DOUBLES = 2 * IN_DOUBLES + HEX_TP * EF + 3 * GEO_DOUBLES + 15 * EN + 5 * EN + 3 * n;
};

#ifdef ESDG_PHASE_TIMING   // debug build only: cycles per activity (thread 0 of each team), summed over CTAs at kernel end
__device__ unsigned long long g_phase_cycles[16];
#define WS_MARK(k)                                                                          \
    do {                                                                                    \
        if ((threadIdx.x & 255) == 0) {                                                     \
            const long long t_ = clock64();                                                 \
            s_cyc[k] += (unsigned long long)(t_ - t_phase);                                 \
            t_phase = t_;                                                                   \
        }                                                                                   \
    } while (0)
#else
#define WS_MARK(k)
#endif

__device__ __forceinline__ void bar_sync_points(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_sync_lines(int nthreads) { asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); }
#ifndef ESDG_WS_PRIMS_ON_LINES
#define ESDG_WS_PRIMS_ON_LINES 1   // item A (conservative -> primitive, in place) by the line team instead of the point team
#endif

template <int n, bool DIAG>
__global__ void __launch_bounds__(WsDims<n>::THREADS, 1) k_hex_stage_ws(const __grid_constant__ HexParams p) {
    using D = WsDims<n>;
    using G = HexGeom<n>;
    constexpr int Nq = G::Nq, Nfq = G::Nfq, NL = G::NL, EB = D::EB, EN = D::EN, EF = D::EF, NLT = D::NLT, NPT = D::NPT;
    extern __shared__ __align__(16) double sm[];
    __shared__ unsigned long long full_tp[2], full_wv[2], empty_in[2], full_a3, empty_a3;
    __shared__ double red[NPT / 32];
    double* sIn = sm;                           // 2 x { WV[7][ENP], WF[8][EF] }
    double* sPx = sIn + 2 * D::IN_DOUBLES;      // P[8][EF] exterior traces: every line thread fetches and reads only its own two points
    double* sGeo3 = sPx + HEX_TP * EF;    // 3 x geo[EB][40]: batch j uses slot j % 3 (its update runs two steps after its loads)
    double* sA3 = sGeo3 + 3 * D::GEO_DOUBLES;   // [3][5][EN]
    double* sV = sA3 + 15 * EN;                 // [5][EN]
    double* s_lm = sV + 5 * EN;
    double* s_lp = s_lm + n;
    double* s_w = s_lp + n;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&full_tp[b], 2 + NPT);            // two expect_tx arrivals (traces, Q + stash) + one cp.async arrival per point thread
            mbar_init(&full_wv[b], NPT);
            mbar_init(&empty_in[b], NLT);
        }
        mbar_init(&full_a3, NLT);
        mbar_init(&empty_a3, NPT);
        mbar_fence_init();
    }
    if (tid < n) {
        s_lm[tid] = p.lm[tid];
        s_lp[tid] = p.lp[tid];
        s_w[tid] = p.w[tid];
    }
    __syncthreads();
    const long long nbatch = range_batches(p, EB);
#ifdef ESDG_PHASE_TIMING
    __shared__ unsigned long long s_cyc[16];
    if (tid < 16) s_cyc[tid] = 0;
    __syncthreads();
    long long t_phase = clock64();
#endif

    if (tid < NLT) {
        // ================================ line warps ================================================
        if constexpr (D::LREGS >= D::LAUNCH_REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(D::LREGS));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(D::LREGS));
        const double c_igm = 1.0 / (2.0 * (p.gamma - 1.0));
        // exterior traces of this thread's two face points: indices read one batch ahead, copies issued as soon as the
        // thread is done with the previous ones (no other thread touches these shared-memory slots)
        int ipn[2] = {0, 0};
        bool bad_l = false;
        {
            long long e0;
            int nb;
            range_batch(p, EB, blockIdx.x, e0, nb);
            if (tid < nb * NL) {
                line_mapP<n, EB>(tid, e0, p, ipn);
                line_gather<n, EB>(tid, ipn, p, sPx);
            }
        }
        unsigned it = 0;
        for (long long bi = blockIdx.x; bi < nbatch; bi += gridDim.x, ++it) {
            const int b = it & 1;
            long long e0, e1 = 0;
            int nb, nb1 = 0;
            range_batch(p, EB, bi, e0, nb);
            if (bi + gridDim.x < nbatch) range_batch(p, EB, bi + gridDim.x, e1, nb1);
            const double* sWV0 = sIn + b * D::IN_DOUBLES;
            const double* sWV = sWV0 + ((e0 * Nq) & 1);       // batch stored with the parity of its first global node index
            const double* sWF = sWV0 + D::WV_DOUBLES;
            const double* sGeo = sGeo3 + (it % 3) * D::GEO_DOUBLES;
            const bool active = tid < nb * NL, active1 = tid < nb1 * NL;
            double acc[n][5];
            if (active1) line_mapP<n, EB>(tid, e1, p, ipn);
            WS_MARK(0);
            mbar_wait(&full_wv[b], (it >> 1) & 1);
            WS_MARK(1);
            mbar_wait(&full_tp[b], (it >> 1) & 1);
#if ESDG_WS_PRIMS_ON_LINES
            {   // A: the point team is the busier one (profiles/r2_phase_timing.txt): the 256 line threads convert the batch
                double* wv = const_cast<double*>(sWV);
                for (int k = tid; k < nb * Nq; k += NLT) node_prims_inplace<n, EB>(k, p, wv, bad_l);
                bar_sync_lines(NLT);
            }
#endif
            WS_MARK(2);
#if defined(ESDG_WS_EXPERIMENT) && ESDG_WS_EXPERIMENT == 2   // timing experiment: point team alone
#pragma unroll
            for (int a = 0; a < n; ++a)
#pragma unroll
                for (int f = 0; f < 5; ++f) acc[a][f] = 0.0 * sWV[a * 5 + f];
            cp_async_wait_all();
#else
            if (active) line_compute<n, EB, DIAG, true, ESDG_LINE_WC_WS>(tid, sWV, sWF, sPx, sGeo, p, c_igm, acc);
#endif
            mbar_arrive(&empty_in[b]);
            if (active1) line_gather<n, EB>(tid, ipn, p, sPx);
            WS_MARK(3);
            if (it > 0) mbar_wait(&empty_a3, (it - 1) & 1);
            WS_MARK(4);
            if (active) line_store<n, EB>(tid, acc, sA3);
            mbar_arrive(&full_a3);
            WS_MARK(5);
        }
        if (bad_l) atomicOr(p.bad, 1);
    } else {
        // ================================ point warps ===============================================
        if constexpr (D::PREGS > D::LAUNCH_REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(D::PREGS));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(D::PREGS));
        const int pt = tid - NLT;
        constexpr int MAXF = (EF + NPT - 1) / NPT;
        const bool bulk = trace_runs_aligned(p);
        const bool bulkq = ((p.q_stride & 1) == 0) && ((((size_t)p.Q | (size_t)p.Lcur) & 15) == 0);
        bool bad = false;
        int ipv[MAXF] = {};   // (the exterior traces are gathered by the line threads)
        // E and F of a batch whose line results are (about to be) in A3
        auto finish = [&](long long e0, int nb, unsigned j) {
            const double* sGeo = sGeo3 + (j % 3) * D::GEO_DOUBLES;
            double tsum = 0.0;
            constexpr int RE = (EN + NPT - 1) / NPT;
            NodeIn cur, nxt;
            if (pt < nb * Nq) node_load(e0 * Nq + pt, p, cur);     // in flight while the line warps finish the batch
            WS_MARK(6);
            mbar_wait(&full_a3, j & 1);
            WS_MARK(7);
#pragma unroll
            for (int r = 0; r < RE; ++r) {
                const int k = pt + r * NPT;
                if (r + 1 < RE && k + NPT < nb * Nq) node_load(e0 * Nq + k + NPT, p, nxt);
                if (k < nb * Nq) node_update<n, EB>(k, e0 * Nq + k, p, cur, sA3, sV, sGeo, s_w, tsum, bad);
                cur = nxt;
            }
            mbar_arrive(&empty_a3);
            WS_MARK(8);
            if (p.mode == 1) {
                bar_sync_points(NPT);
                WS_MARK(9);
                for (int l = pt; l < nb * NL; l += NPT) line_traces<n, EB>(l, e0, p, sV, sGeo, s_lm, s_lp);
                WS_MARK(10);
            }
            if (p.want_rhstest) {
                for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(0xffffffffu, tsum, o);
                if ((pt & 31) == 0) red[pt >> 5] = tsum;
                bar_sync_points(NPT);
                if (pt == 0) {
                    double t = 0.0;
                    for (int wv = 0; wv < NPT / 32; ++wv) t += red[wv];
                    p.partial[e0] = t;
                }
            }
            bar_sync_points(NPT);   // V (and red) free for the next batch
            WS_MARK(11);
        };
        // loads of a batch: own traces (bulk copies), exterior traces (gather), Q and the log stash straight into the
        // WV planes (converted in place by `prims`), L2 prefetch of what the update phase will read from global memory
        auto issue = [&](long long e0, int nb, unsigned j) {
            const int b = j & 1;
            double* sWV = sIn + b * D::IN_DOUBLES;
            double* sWF = sWV + D::WV_DOUBLES;
            issue_trace_loads<n, EB, NPT, false>(pt, nb, e0, p, sWF, nullptr, &full_tp[b], bulk, ipv);
            // Q (5 planes) and the log stash (2) into the WV planes.  Node k of the batch lives at WV[f][off + k] with
            // off = parity of its global node index, so that the even-aligned middle of every run moves as ONE 16-byte
            // aligned bulk copy; at most one leading / trailing double per plane goes by cp.async
            const long long g0 = e0 * Nq;
            const int cnt = nb * Nq, off = (int)(g0 & 1);
            if (bulkq) {
                const int lead = off, mid = (cnt - lead) & ~1, tail = cnt - lead - mid;
                if (pt == 32) mbar_expect_tx(&full_tp[b], 7u * (unsigned)(mid * 8));
                if (pt >= 32 && pt < 39) {
                    const int f = pt - 32;
                    const double* src = (f < 5 ? p.Q + f * p.q_stride : p.Lcur + (f - 5) * p.q_stride) + g0;
                    double* dst = sWV + f * D::ENP + off;
                    bulk_g2s(dst + lead, src + lead, (unsigned)(mid * 8), &full_tp[b]);
                    if (lead) cp_async8(dst, src);
                    if (tail) cp_async8(dst + cnt - 1, src + cnt - 1);
                }
            } else {
                if (pt == 32) mbar_arrive(&full_tp[b]);
                for (int k = pt; k < cnt; k += NPT) {
                    const long long g = g0 + k;
#pragma unroll
                    for (int f = 0; f < 5; ++f) cp_async8(sWV + f * D::ENP + off + k, p.Q + f * p.q_stride + g);
                    cp_async8(sWV + 5 * D::ENP + off + k, p.Lcur + g);
                    cp_async8(sWV + 6 * D::ENP + off + k, p.Lcur + p.q_stride + g);
                }
            }
            cp_async_mbar_arrive(&full_tp[b]);
            stage_geo_async<EB>(pt, NPT, nb, e0, p, sGeo3 + (j % 3) * D::GEO_DOUBLES);
            if (p.mode == 1) {
                for (int k = pt * 16; k < nb * Nq; k += NPT * 16) {   // one 128-byte line per 16 doubles
#pragma unroll
                    for (int f = 0; f < 5; ++f) prefetch_l2(p.res + f * p.q_stride + e0 * Nq + k);
                }
            }
        };
        // A: geometry records + conservative -> primitive in place (each thread converts the nodes it copied itself)
        auto prims = [&](long long e0, int nb, unsigned j) {
            const int b = j & 1;
            double* sWV = sIn + b * D::IN_DOUBLES;
            double* sGeo = sGeo3 + (j % 3) * D::GEO_DOUBLES;
            cp_async_wait_all();
            stage_geo_finish<EB>(pt, NPT, nb, sGeo);
#if !ESDG_WS_PRIMS_ON_LINES
            mbar_wait(&full_tp[b], (j >> 1) & 1);          // Q and the stash have landed (with the traces)
            sWV += (e0 * Nq) & 1;
            for (int k = pt; k < nb * Nq; k += NPT) node_prims_inplace<n, EB>(k, p, sWV, bad);
#else
            (void)sWV;
#endif
            mbar_arrive(&full_wv[b]);
        };
        const unsigned nloc = (unsigned)((nbatch - blockIdx.x + gridDim.x - 1) / gridDim.x);   // batches of this CTA
        auto batch_e0 = [&](unsigned i) {
            long long e0;
            int nb;
            range_batch(p, EB, (long long)blockIdx.x + (long long)i * gridDim.x, e0, nb);
            return e0;
        };
        auto batch_nb = [&](unsigned i) {
            long long e0;
            int nb;
            range_batch(p, EB, (long long)blockIdx.x + (long long)i * gridDim.x, e0, nb);
            return nb;
        };
        // step j runs beside the line warps' batch j-1: loads of batch j, update + traces of batch j-2, primitives of
        // batch j (single call sites, so that the three lambdas are inlined)
#if defined(ESDG_WS_EXPERIMENT) && ESDG_WS_EXPERIMENT == 1   // timing experiment: line team alone (garbage data)
        for (unsigned j = 0; j < nloc + 2; ++j) {
            if (j < nloc) {
                const int b = j & 1;
                if (j >= 2) mbar_wait(&empty_in[b], ((j >> 1) - 1) & 1);
                if (pt == 0 || pt == 32) mbar_arrive(&full_tp[b]);
                mbar_arrive(&full_tp[b]);
                mbar_arrive(&full_wv[b]);
            }
            if (j >= 2) {
                mbar_wait(&full_a3, (j - 2) & 1);
                mbar_arrive(&empty_a3);
            }
        }
#else
        for (unsigned j = 0; j < nloc + 2; ++j) {
            if (j < nloc) {
                const int b = j & 1;
                WS_MARK(6);
                if (j >= 2) mbar_wait(&empty_in[b], ((j >> 1) - 1) & 1);
                WS_MARK(12);
                issue(batch_e0(j), batch_nb(j), j);
                WS_MARK(13);
            }
            if (j >= 2) finish(batch_e0(j - 2), batch_nb(j - 2), j - 2);
            if (j < nloc) {
                prims(batch_e0(j), batch_nb(j), j);
                WS_MARK(14);
            }
        }
#endif
        if (bad) atomicOr(p.bad, 1);
    }
#ifdef ESDG_PHASE_TIMING
    __syncthreads();
    if (tid < 16) atomicAdd(&g_phase_cycles[tid], s_cyc[tid]);
#endif
}

}  // namespace esdg
