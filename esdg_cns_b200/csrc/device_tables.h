// Device-side description of one discretisation (tables built by esdg_finalize).
#pragma once
#include <stdint.h>

#include "kernels_mma.cuh"

namespace esdg {

// Row-compressed (ELL) sparse matrix: row r holds `cnt[r]` entries at
// col[r*width + k], val[r*width + k].  Dense operators (triangles) are simply
// full-width rows.  Built from the caller's dense matrices with the reference's
// drop tolerance (dg2D_euler_quad.jl:62-63, :76-78).
struct Ell {
    int rows, width;
    const int* cnt;
    const int* col;
    const double* val;
};

// Union sparsity of the skew hybridized SBP operators (Qrsids / Qnzids,
// dg2D_euler_quad.jl:64, dg3D_euler_hex.jl:58) with the (r,s,t) coefficients of
// each entry: c[(r*width+k)*3 + d].
struct PairEll {
    int rows, width;
    const int* cnt;
    const int* col;
    const double* c;
};

struct GenericParams {
    // sizes
    int Ns, Nq, Nfq, Nh, Nfaces, Nfp;
    int collocated;
    int EB;                 // elements per CTA
    long long K;            // local elements
    long long trace_stride; // doubles per field of the trace array (Nfq*K + n_halo)
    double gamma, lf_scale, volume_scale;
    double lg1, igm1x2;     // log(gamma-1), 1/(2(gamma-1)) (host-computed)
    // operators
    Ell Vq;                 // Nq x Ns   (modal only)
    Ell Proj;               // collocated: Ef (Nfq x Nq); modal: VhP (Nh x Nq)
    Ell Ph;                 // Ns x Nh
    Ell Lf;                 // Ns x Nfq
    PairEll Qp;             // Nh rows
    // symmetric schedule of the same pairs (skew operators, symmetric flux: each unordered pair once):
    // in round t the thread of row r evaluates F(r, sym_dst[t*Nh + r]) (or idles when it is < 0) with the
    // coefficients sym_c[(t*Nh + r)*3 + d] = Q_d[r][dst]; no row is the destination of two threads in a round
    int sym_rounds;         // 0: row-wise evaluation of Qp
    const int* sym_dst;
    const double* sym_c;
    // modal path: the same operators dense + zero-padded for the tensor-core batch GEMMs (kernels_mma.cuh);
    // dPh carries volume_scale, dProjF = face rows of VhP
    Dense dVq, dProj, dProjF, dPh, dLf;
    const double* wq;       // Nq
    // geometry, per element: [e*geo_stride + ...]
    //   0..DIM*DIM-1 : G[d][k] = d(ref k)/d(x_d)*J  i.e. (rxJ,sxJ,txJ | ryJ,syJ,tyJ | rzJ,szJ,tzJ)
    //   DIM*DIM      : J
    //   then per face: nxJ,nyJ[,nzJ],sJ
    const double* geo;
    int geo_stride;
    // curved hexes (desc.affine = 0; dg3D_euler_hex.jl:90-101, :145-151): geometric factors at every hybridized
    // point, J at every node, scaled normals at every face point.  Per element: [e][9][Nh], [e][Nq], [e][4][Nfq].
    int curved;
    const double *geoG, *geoJ, *geoN;
    const int* mapP;        // [K][Nfq] 0-based index into the per-field trace array
    // state / work arrays (field-major: base + f*stride)
    const double* Q;        // Ns*K per field
    long long q_stride;
    double* traces;         // trace_stride per field
    double* rhs;            // output of rhs mode (Ns*K per field) or nullptr
    double* Qout;           // RK mode: updated state
    double* res;            // RK mode: LSRK register
    double rk_a, rk_b, dt;
    int mode;               // 0: write rhs ; 1: fused LSRK stage update
    int want_rhstest;
    double* rhstest_partial; // one double per CTA
    int* bad_flag;
    // operator tables are copied into shared memory at kernel start when they fit (byte offset of the table
    // region in dynamic shared memory; 0 = leave them in global memory)
    int stage_tables;
    unsigned table_off;
};

}  // namespace esdg
