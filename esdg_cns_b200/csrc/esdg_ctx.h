// Host-side context behind the opaque `esdg_ctx` of include/esdg_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/esdg_b200.h"
#include "device_tables.h"

struct HostMat {
    int64_t rows = 0, cols = 0;
    std::vector<double> a;  // column-major
    double at(int64_t i, int64_t j) const { return a[i + rows * j]; }
};

struct DevEll {
    int rows = 0, width = 0;
    int* cnt = nullptr;
    int* col = nullptr;
    double* val = nullptr;
    esdg::Ell view() const { return esdg::Ell{rows, width, cnt, col, val}; }
};

struct KernelTimer {
    std::string name;
    cudaEvent_t a, b;
};

// Compressible Navier-Stokes extension (kernels_cns.cuh): settings given before esdg_finalize and
// the device data built by it.
struct CnsState {
    bool on = false;                      // esdg_set_viscosity was called
    double lambda = 0, mu = 0, Pr = 0, Re = 0;
    int viscous_dissp = 0, bctype = 1;
    std::vector<int64_t> wall, lid, mapB; // 1-based linear indices into Nfq x K (Julia)
    std::vector<double> vlid;
    std::vector<int64_t> inflow, outflow; // shock-tube boundary nodes (dg2D_CNS_modalESDG.jl:165-166)
    double inflow_state[4] = {1, 0, 0, 1}; // rho, u, v, p of the inflow
    bool have_mapB = false;
    DevEll Pq, Vf, Dr, Ds;
    esdg::Dense dPq{}, dVf{}, dDr{}, dDs{};
    unsigned char* bc = nullptr;
    double* vlid_dev = nullptr;
    double *sig = nullptr, *sig_tr = nullptr, *partial_visc = nullptr, *partial_vrhs = nullptr;
    bool sigma_valid = false;             // distributed stage: sigma pass done, rhs pass pending
};

// DOPRI45 driver state (dg2D_CNS_cavity_optimized.jl:919-934, :974-1053)
struct DopriState {
    bool ready = false;
    double* K[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // rhsQrk[1..7]
    double* Qtmp = nullptr;
    double *err_partial = nullptr, *err_dev = nullptr;
    int n_err_blocks = 0;
    double prev_err = 0.0;
    long long attempts = 0;    // `i` of the script: total steps attempted
};

struct esdg_ctx {
    esdg_desc d{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool finalized = false;
    std::string err;

    std::map<std::string, HostMat> ops, geom;
    std::vector<int64_t> mapP_host;
    std::vector<int64_t> send_list;

    // derived sizes
    int Nh = 0, Nfp = 0, NF = 0, EB = 1, threads = 256, geo_stride = 0;
    size_t smem_bytes = 0;
    size_t smem_cns_sigma = 0, smem_cns_traces = 0, smem_cns_rhs = 0;   // compact per-kernel layouts (kernels_cns.cuh)
    size_t smem_generic_traces = 0;                   // k_generic_traces: U, V and one face array
    int stage_tables = 0;          // operator tables copied to shared memory by the kernels
    unsigned table_off = 0;        // byte offset of the table region in dynamic shared memory
    long long trace_stride = 0;

    // device tables
    DevEll Vq, Proj, Ph, Lf;
    esdg::Dense dVq{}, dProj{}, dProjF{}, dPh{}, dLf{};   // modal path (tensor-core batch GEMMs)
    int *qp_cnt = nullptr, *qp_col = nullptr;
    double* qp_c = nullptr;
    int qp_width = 0;
    int sym_rounds = 0;              // symmetric pair schedule (device_tables.h); 0 = row-wise
    int* sym_dst = nullptr;
    double* sym_c = nullptr;
    double* wq = nullptr;
    double* geo = nullptr;
    double *geoG = nullptr, *geoJ = nullptr, *geoN = nullptr;   // curved hexes (desc.affine = 0)
    int* mapP = nullptr;
    int* send_idx = nullptr;
    long long n_send = 0;

    // state
    double *Q = nullptr, *res = nullptr, *rhs = nullptr, *traces = nullptr;
    double *Qin = nullptr;  // staging for esdg_rhs(Q given)
    double *partials = nullptr, *rhstest_dev = nullptr;
    int n_partials = 0;
    int* bad_flag = nullptr;
    bool traces_valid = false;

    // specialised hex path (kernels_hex.cuh); null if the operators do not match
    struct HexFast* hex = nullptr;
    // specialised quad path (kernels_quad_line.cuh); null if the operators do not match
    struct QuadFast* quad = nullptr;
    CnsState cns;
    DopriState dopri;
    double* rhs_target = nullptr;   // where mode-0 rhs evaluations write (default: rhs)
    // pipelined host step (esdg_lsrk45_host): copy streams, per-chunk events, chunk dependency masks
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> chunk_up, chunk_done;
    std::vector<unsigned long long> chunk_dep;   // bit c' set: the stage of this chunk reads traces of chunk c'
    std::vector<long long> chunk_begin; // nchunk + 1 element offsets (contiguous chunks) ...
    long long chunk_seg_stride = 0;     // ... or, > 0: chunk c = segments [chunk_begin[c], chunk_begin[c+1]) + s * chunk_seg_stride, s < chunk_nseg
    long long chunk_nseg = 1;
    long long bnd_lo = 0, bnd_hi = 0;   // leading / trailing elements whose traces are sent to neighbours
    struct EsdgComm* comm = nullptr;    // NCCL communicator + neighbour lists (esdg_comm_init)
    // structured hex mesh generated on the device (esdg_set_structured_hex): global element counts, box, periodicity,
    // local z-slab; maps and geometric factors never cross the ABI
    bool structured = false;
    long long sK[3] = {0, 0, 0}, skz0 = 0, skzn = 0;
    double sbox[6] = {0, 0, 0, 0, 0, 0};
    int sper[3] = {1, 1, 1};
    bool res_zero = true;

    // accounting
    int64_t launches = 0;
    bool timing = false;
    std::vector<KernelTimer> timers;
    std::map<std::string, std::pair<double, int64_t>> time_acc;

    std::vector<void*> allocs;
};
