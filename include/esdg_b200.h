/* esdg_b200.h -- C ABI of libesdg_b200.so
 *
 * B200-native (sm_100a) replacement for ONE hot path of yiminllin/ESDG-CNS: the
 * semi-discrete entropy-stable DG right-hand side and the low-storage RK loop
 * that drives it.  The reference has no FFI; its boundary is the Julia call
 * shape
 *     rhsQ, rhstest = rhs(Q, md::MeshData, ops, flux_fun, compute_rhstest)
 *         examples/dg2D_euler_quad.jl:141, dg2D_euler_tri.jl:130, dg3D_euler_hex.jl:167
 *     for INTRK=1:5  resQ = rk4a*resQ + dt*rhsQ ; Q += rk4b*resQ
 *         examples/dg2D_euler_quad.jl:198-207
 * A Julia shim pulls the arrays out of rd::RefElemData
 * (src/SetupDG.jl:38-75), md::MeshData (:77-115) and the `ops` tuple and hands
 * them over as plain column-major double / int64 arrays.  No torch types, no
 * exceptions across the boundary; every function returns an esdg_status.
 *
 * All matrices are COLUMN-MAJOR doubles exactly as Julia stores them; all index
 * arrays are int64 and 1-BASED exactly as Julia stores them.  Host pointers are
 * borrowed for the duration of the call only.
 */
#ifndef ESDG_B200_H
#define ESDG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct esdg_ctx esdg_ctx;

typedef enum {
    ESDG_OK = 0,
    ESDG_EINVAL = 1,  /* bad argument / missing operator / mesh not as declared */
    ESDG_ECUDA = 2,   /* CUDA runtime error (esdg_last_error has the string)    */
    ESDG_ENCCL = 3,   /* NCCL call failed / libnccl not found (esdg_comm_*)      */
    ESDG_ENOMEM = 4,  /* device allocation failed                               */
    ESDG_ESTATE = 5   /* non-positive density/pressure met in rhs: the Julia
                         DomainError of log/sqrt (euler_variables.jl:67, :8)   */
} esdg_status;

enum { ESDG_TRI = 0, ESDG_QUAD = 1, ESDG_HEX = 2 };

/* Discretisation descriptor: the script-top globals + array sizes.
 * (dg2D_euler_quad.jl:21-24,91 ; dg3D_euler_hex.jl:20-23,98) */
typedef struct {
    int32_t dim;        /* 2 or 3                                                           */
    int32_t elem;       /* ESDG_TRI / ESDG_QUAD / ESDG_HEX                                  */
    int32_t N;          /* polynomial degree                                                */
    int32_t Nfields;    /* dim + 2                                                          */
    int32_t Ns;         /* rows of each Q[f]: Nq (quad/hex scripts) or Np (tri/CNS scripts) */
    int32_t Nq;         /* volume quadrature points  (length of rd.wq)                      */
    int32_t Nfq;        /* face quadrature points, all faces (length of rd.wf)              */
    int32_t Nfaces;     /* rd.Nfaces                                                        */
    int32_t collocated; /* 1: state lives at quadrature nodes, volume part unprojected
                              (dg2D_euler_quad.jl:149-151, dg3D_euler_hex.jl:176-178)
                           0: Vq / VhP projection of every hybridized point
                              (dg2D_euler_tri.jl:138-140)                                   */
    int32_t affine;     /* 1: geofacs constant per element (`vgeo[1,e]`, dg2D_euler_quad.jl:175);
                           verified by esdg_finalize.  0 (collocated hexes only): curved
                           elements, geofacs at every hybridized point averaged per pair
                           (dg3D_euler_hex.jl:90-101, :145-151), J per node, normals per face
                           point; esdg_set_geometry then needs the full Nh / Nq / Nfq x K arrays */
    int64_t K;          /* elements owned by this context (md.K, or the local slab)        */
    int64_t n_halo;     /* extra trace points appended after Nfq*K that mapP may point to
                           (neighbour-rank face traces); 0 on a single GPU                  */
    double gamma;       /* 1.4 (EntropyStableEuler.jl:9)                                    */
    double lf_scale;    /* c in LFc = c*max(lam,lamP)*sJ : .5 quad(:165) .25 tri(:156) 0 hex(:193) */
    double volume_scale;/* factor on Ph*QF : 2 (quad :180, tri :171); 1 when the caller's Ph
                           already holds it (hex, dg3D_euler_hex.jl:96,208)                 */
} esdg_desc;

/* ---- lifecycle -------------------------------------------------------------- */
int esdg_create(const esdg_desc* desc, int device, esdg_ctx** out);
int esdg_destroy(esdg_ctx* ctx);
const char* esdg_last_error(const esdg_ctx* ctx);
/* Run all work of this context on an existing CUDA stream (cudaStream_t as void*);
 * default is a stream the context owns. */
int esdg_set_stream(esdg_ctx* ctx, void* cuda_stream);

/* ---- setup: replaces the `@unpack ... = rd/md` + `ops` tuple of rhs() ----------
 * name: "Ef"  (Nfq x Nq, Vf*Pq; dg2D_euler_quad.jl:50)           collocated only
 *       "Vq"  (Nq x Ns), "VhP" (Nh x Nq)   (dg2D_euler_tri.jl:58)    modal only
 *       "Ph"  (Ns x Nh)  (dg2D_euler_quad.jl:77 / dg2D_euler_tri.jl:57)
 *       "Lf"  (Ns x Nfq) quad/hex Lf (:78) or rd.LIFT (SetupDG.jl:192)
 *       "Qrh","Qsh","Qth" (Nh x Nh) skew hybridized SBP operators (:60-61)
 *       "wq"  (Nq x 1)   rd.wq  (only used for rhstest through wJq)
 *       CNS path only (rd fields, src/SetupDG.jl:38-75): "Pq" (Ns x Nq), "Vf" (Nfq x Ns),
 *       "Dr","Ds" (Ns x Ns); "Lf" is rd.LIFT
 * Sparse Julia matrices are passed densified (`Matrix(A)`); sparsity is re-detected
 * with the reference's drop tolerance 1e-12 (dg2D_euler_quad.jl:62-63). */
int esdg_set_operator(esdg_ctx* ctx, const char* name, const double* colmajor, int64_t rows, int64_t cols);

/* name: "rxJ","sxJ","txJ","ryJ","syJ","tyJ","rzJ","szJ","tzJ" (rows = Nh, Np or 1; row 1 is
 *        read, the `[1,e]` of dg2D_euler_quad.jl:175), "J" (rows = Ns or 1),
 *       "nxJ","nyJ","nzJ","sJ" (rows = Nfq, or Nfaces when constant per face),
 *       "wJq" (Nq x K or 1 x K: then wq[i]*wJq[e] is used; SetupDG.jl:307)
 * cols must equal desc.K. */
int esdg_set_geometry(esdg_ctx* ctx, const char* name, const double* colmajor, int64_t rows, int64_t cols);

/* md.mapP (Nfq x K, 1-based linear index into an Nfq x (K [+halo]) array; SetupDG.jl:294-298). */
int esdg_set_maps(esdg_ctx* ctx, const int64_t* mapP, int64_t n);

/* ---- compressible Navier-Stokes (examples/CompressibleNS/dg2D_CNS_cavity_optimized.jl) ----------
 * Calling esdg_set_viscosity before esdg_finalize turns the context into the CNS rhs
 *     rhsQ, rhstest, rhstest_visc = rhsRK!(Q, rd, md, Re, BCTYPE, ops, flux, inviscid_dissp, viscous_dissp, ...)   (:955-972)
 * = rhs_inviscid! (:447-528) + rhs_viscous! (BR1, :749-849).  2D, Vq/Pq projection (collocated = 0).
 * lambda, mu, Pr are the arguments of init_visc_fxn (:613; lambda's sign is flipped inside exactly as
 * the reference does, :614), Re scales the penalisation tau = -1/Re/v4 (:821), viscous_dissp is the
 * script flag (:30).  inviscid_dissp (:29) maps to desc.lf_scale = .25 (on, :508) or 0 (off). */
int esdg_set_viscosity(esdg_ctx* ctx, double lambda, double mu, double Pr, double Re, int viscous_dissp);
/* Wall boundary conditions, init_BC_funs (:135-264).  bctype = BCTYPE (:27): 1 adiabatic no-slip,
 * 2 isothermal, 3 slip/reflective.  wall_idx / lid_idx: 1-based linear indices into the Nfq x K face
 * arrays (the `wall` and `lid` subsets of md.mapB, :139-140), vlid: lid speed per lid node (:147).
 * The inviscid mirror state (:157-176), the entropy-variable BCs (:178-216) and the stress BCs
 * (:218-262) are applied at these nodes. */
int esdg_set_wall_bc(esdg_ctx* ctx, int bctype, const int64_t* wall_idx, int64_t nwall, const int64_t* lid_idx,
                     const double* vlid, int64_t nlid);
/* Inflow / outflow nodes of the viscous shock tube (examples/CompressibleNS/dg2D_CNS_modalESDG.jl, init_BC_funs
 * :161-218).  inflow_idx = `leftwall`: exterior primitive state fixed to state = (rho, u, v, p) (:172-175), exterior
 * entropy variables v_ufun of it (:191-197).  outflow_idx = `rightwall`: exterior = interior (:177-180, :199-202).
 * On both sets the Lax-Friedrichs coefficient is zero (impose_BCs_lam!, :183-188) and the exterior stress equals the
 * interior one (:205-215).  1-based linear indices into the Nfq x K face arrays; mapP may connect these nodes to
 * anything (the script makes the box periodic first, :80-86).  Can be combined with esdg_set_wall_bc. */
int esdg_set_inflow_outflow_bc(esdg_ctx* ctx, const int64_t* inflow_idx, int64_t n_in, const double* state /* [4] */,
                               const int64_t* outflow_idx, int64_t n_out);
/* md.mapB (src/node_map_functions.jl:53): nodes where the penalisation switches to its wall form
 * (:829-839).  Default when not called: wall + lid. */
int esdg_set_boundary_nodes(esdg_ctx* ctx, const int64_t* mapB, int64_t nB);

/* Validates that everything required by the descriptor was supplied, compresses the
 * operators and uploads them.  Must be called once before any hot call. */
/* GPU-side setup of a structured hex mesh, INSTEAD of esdg_set_geometry + esdg_set_maps: mapP and
 * the geometric factors of an axis-aligned box of K3[0] x K3[1] x K3[2] elements (element index x fastest,
 * src/UniformHexMesh.jl:59-63; faces r-, r+, s-, s+, t-, t+; matching face points share their on-face index, which is
 * what connect_mesh.jl + node_map_functions.jl:23-55,139-213 produce for the r-fastest node ordering) are generated
 * on the device by esdg_finalize -- a 128^3 mesh needs six integers instead of 2.5 GB of Int64 maps.  This context holds
 * the layers kz_begin .. kz_begin + kz_count - 1; trace points owned by the slab below / above map to the halo slots
 * Nfq K + (i + Kx j) Nfp + m  (+ Kx Ky Nfp for the upper slab): desc.n_halo = 2 Kx Ky Nfp for a slab, 0 for the whole mesh. */
int esdg_set_structured_hex(esdg_ctx* ctx, const int64_t K3[3], const double box[6] /* x0,x1,y0,y1,z0,z1 */, const int periodic[3],
                            int64_t kz_begin, int64_t kz_count);
int esdg_finalize(esdg_ctx* ctx);

/* ---- hot path ------------------------------------------------------------------
 * Q / rhs_out: arrays of Nfields HOST pointers, each Ns*K doubles (the Julia
 * Vector{Matrix{Float64}}).  Q == NULL evaluates the resident state.  Synchronous.
 * rhstest[0] receives sum_f sum(wJq .* VU[f] .* rhs[f]) when compute_rhstest != 0
 * (dg2D_euler_quad.jl:186-191).  CNS contexts: rhstest must hold 2 doubles, (rhstest, rhstest_visc)
 * of rhsRK! (dg2D_CNS_cavity_optimized.jl:960-971). */
int esdg_rhs(esdg_ctx* ctx, const double* const* Q, double* const* rhs_out, int compute_rhstest, double* rhstest);

/* State residency (Q and the LSRK register resQ live in HBM between calls). */
int esdg_upload_state(esdg_ctx* ctx, const double* const* Q);   /* also zeroes resQ (:197) */
int esdg_download_state(esdg_ctx* ctx, double* const* Q);

/* nsteps x 5 stages of LSRK45 on the resident state (dg2D_euler_quad.jl:196-207,
 * coefficients src/CommonUtils.jl:29-49).  Stream-ordered: returns without waiting
 * unless rhstest_last != NULL (then rhstest of the final stage is returned). */
int esdg_lsrk45_step(esdg_ctx* ctx, double dt, int nsteps, double* rhstest_last);

/* The same loop for HOST state (the Julia `Q`, `Vector{Matrix{Float64}}`): Q_in -> nsteps LSRK45 steps ->
 * Q_out (may alias Q_in).  On the tensor-product hex path the state travels in element chunks and the stages run as
 * a wavefront behind the upload (stage s of a chunk needs stage s-1 of the chunks its mapP touches), with each
 * chunk's download behind its last stage, so both copy directions and the kernels overlap; the result is bitwise
 * that of esdg_upload_state + esdg_lsrk45_step + esdg_download_state.  Pinned host memory is needed for the copies
 * to overlap.  Other paths: plain upload / steps / download.  Synchronous. */
int esdg_lsrk45_host(esdg_ctx* ctx, const double* const* Q_in, double* const* Q_out, double dt, int nsteps);

/* DOPRI45 adaptive driver of the CNS scripts (dg2D_CNS_cavity_optimized.jl:919-934 coefficients,
 * :974-1053 loop): FSAL, Hairer error semi-norm with tolerance err_tol (the script's errTol = 1e-5,
 * :1017), step accepted iff errEst < 1, P / PI step-size controller clipped to [1e-9, 10*dt0].
 * esdg_dopri45_init evaluates the first stage of the resident state (:996-997); each
 * esdg_dopri45_step is ONE iteration of the `while t < T` loop: it attempts a step of *dt, updates
 * *t (if accepted) and *dt (always, to the controller's next value) and reports errEst; rhstest[2]
 * receives (rhstest, rhstest_visc) of the last stage (:1039-1042).  Works on Euler contexts too. */
int esdg_dopri45_init(esdg_ctx* ctx);
int esdg_dopri45_step(esdg_ctx* ctx, double* t, double* dt, double dt0, double err_tol, int* accepted, double* err_est,
                      double* rhstest);

int esdg_sync(esdg_ctx* ctx);   /* waits; reports ESDG_ESTATE if the rhs met rho<=0 or p<=0 */

/* ---- multi-GPU plumbing (one context per GPU; the host exchanges halos) ---------
 * Face traces (entropy-projected conservative variables at face points,
 * `Uf` of dg2D_euler_quad.jl:150) live in one device array of
 * Nfields x (Nfq*K + n_halo) doubles, field-major.  The host library (torch.distributed
 * / NCCL in the Python mirror, NCCL.jl in Julia) sends packed boundary traces and
 * receives straight into the halo tail. */
int esdg_set_send_list(esdg_ctx* ctx, const int64_t* trace_idx /*1-based*/, int64_t n);
int esdg_compute_traces(esdg_ctx* ctx);                     /* traces of the resident Q      */
int esdg_pack_send(esdg_ctx* ctx, void* send_dev /* Nfields*n doubles, device */);
int esdg_halo_ptr(esdg_ctx* ctx, void** halo_dev, int64_t* field_stride);
/* copy a received buffer of Nfields x n_halo doubles (device) into the halo tail */
int esdg_unpack_halo(esdg_ctx* ctx, const void* recv_dev);
/* Distributed LSRK stepping, driven by the caller:
 *   after upload:  esdg_compute_traces ; esdg_pack_send ; <exchange> ; esdg_unpack_halo
 *   every stage :  esdg_stage_update(stage) ; esdg_pack_send ; <exchange> ; esdg_unpack_halo */
/* part: 0 = all local elements; 1 = only the elements whose traces are in the send list (launch
 * first, then pack + start the exchange); 2 = the remaining interior elements (overlaps the exchange).
 * After the call(s) the traces of the NEW state are valid locally: pack / exchange / unpack them
 * before the next stage. */
int esdg_stage_update(esdg_ctx* ctx, int stage /*0..4*/, double dt, int compute_rhstest, int part);
/* CNS contexts with halo points (rhsRK! has two inter-element dependences: the entropy-variable traces
 * `VUP = VUf[mapP]` dg2D_CNS_cavity_optimized.jl:776 and the stress traces `sigma_f[mapP]` :591-600), so one
 * distributed stage is
 *   esdg_cns_sigma_pass ; esdg_pack_send_sigma ; <exchange> ; esdg_unpack_halo_sigma ;
 *   esdg_stage_update(stage, part = 0) ; esdg_pack_send ; <exchange> ; esdg_unpack_halo
 * The sigma buffers hold 6 planes (sigma_x, sigma_y of fields 2..4) per trace point, the entropy-variable
 * buffers esdg_trace_planes() = 4.  compute_rhstest must agree between the two calls of a stage. */
int esdg_cns_sigma_pass(esdg_ctx* ctx, int compute_rhstest);
int esdg_pack_send_sigma(esdg_ctx* ctx, void* send_dev /* 6*n doubles, device */);
int esdg_unpack_halo_sigma(esdg_ctx* ctx, const void* recv_dev /* 6*n_halo doubles, device */);
/* number of double planes per trace point in the pack/unpack buffers (5 = Uf of the generic path,
 * 8 = rho,u,v,w,beta,log rho,log beta,lambda of the tensor-product hex path) */
int esdg_trace_planes(esdg_ctx* ctx);

/* ---- multi-GPU inside the boundary -----------------------------------------------------------------
 * One context per rank (= per GPU); the context owns an NCCL communicator and a communication stream.  After
 * esdg_comm_init, esdg_lsrk45_step runs the partitioned mesh by itself: per stage the boundary-layer elements are
 * updated first, their new traces are gathered into one contiguous message per (neighbour, plane) and sent with
 * ncclSend while the interior elements are updated; ncclRecv lands directly in the halo tail of the trace planes the
 * next stage reads (no staging copies), and rhstest is all-reduced over the ranks.  The Julia caller needs no stepper of
 * its own: create the id on rank 0 (esdg_comm_unique_id), broadcast its 128 bytes (MPI.jl / Distributed), call
 * esdg_comm_init on every rank.  NCCL is resolved with dlopen("libnccl.so.2") (ESDG_NCCL_LIB overrides the name).
 *   neighbours : in the order their blocks appear in the send list (esdg_set_send_list) and in the halo tail
 *   nb_send[b] : number of send-list entries that go to neighbour b      (sum = length of the send list)
 *   nb_recv[b] : number of halo points that come from neighbour b        (sum = desc.n_halo)
 *   nb_rank[b] = -1: a block without a partner (physical boundary): nothing is sent or received, offsets still advance
 * Stream contract: every kernel of a context runs on ITS stream (esdg_set_stream, default: an own non-blocking stream);
 * device buffers handed to esdg_pack_send / esdg_unpack_halo by a host-driven exchange must be produced / consumed on
 * that stream (or ordered against it with events). */
int esdg_comm_unique_id(void* id128 /* out: 128 bytes */);
int esdg_comm_init(esdg_ctx* ctx, const void* id128, int nranks, int rank, int n_neighbours, const int* nb_rank,
                   const int64_t* nb_send, const int64_t* nb_recv);
int esdg_comm_destroy(esdg_ctx* ctx);
/* which kernel family a finalized context runs: "hex_line_diag" (tensor-product hexes, axis-aligned elements),
 * "hex_line" (general affine hexes), "hex_line_curved" (desc.affine = 0: per-pair averaged geometric factors), "quad_line" (tensor-product quads, dg2D_euler_quad.jl), "generic" (table-driven),
 * "cns_quad_line" (Navier-Stokes on quads, state kept in the Gauss-nodal basis between upload and download -- invisible
 * at this boundary: every entry point takes and returns the caller's coefficients), "cns" (table-driven modal kernels:
 * triangles, inflow / outflow contexts).  ESDG_FORCE_GENERIC=1 / ESDG_CURVED_GENERIC=1 / ESDG_CNS_MODAL=1 in the environment select the
 * table-driven families. */
const char* esdg_kernel_path(const esdg_ctx* ctx);
int esdg_read_rhstest(esdg_ctx* ctx, double* rhstest);
int esdg_read_rhstest_visc(esdg_ctx* ctx, double* rhstest_visc);   /* CNS contexts */

/* ---- 1-D compressible Navier-Stokes with SSPRK(3,3) -----------------------------------------
 * examples/CompressibleNS/dg1D_CNS_modalESDG.jl: modal ESDG on K intervals, `rhs_ESDG(U, ..., t)` (:338-353 =
 * rhs_inviscid :220-268 + rhs_viscous :270-336) and the loop `while t < T` (:377-391).  The script's operators
 * (rd.Vq, rd.Pq, rd.Vf, rd.LIFT, rd.Dr of init_reference_interval; VhP, Ph, Qrh_skew of :142-156) are handed over
 * as column-major matrices; J = (XR-XL)/K/2, rxJ = 1.  The states the script imposes on the two domain ends (the
 * exact solution at time t, :339-349) are passed per call as
 *   bc12 = { rhoL, uL, betaL, rhoR, uR, betaR, v1L, v2L, v3L, v1R, v2R, v3R }
 * so that the (host-side, scalar) bisection of `exact_sol_viscous_shocktube` stays in the caller. */
typedef struct esdg1d_ctx esdg1d_ctx;
int esdg1d_create(int device, int N, int64_t K, int Nq, const double* Vq, const double* Pq, const double* Vf, const double* LIFT,
                  const double* Dr, const double* VhP, const double* Ph, const double* Qrh_skew, double J, double rxJ, double gamma,
                  double mu, double lambda, double kappa_over_cv, esdg1d_ctx** out);
int esdg1d_destroy(esdg1d_ctx* ctx);
const char* esdg1d_last_error(const esdg1d_ctx* ctx);
int esdg1d_upload_state(esdg1d_ctx* ctx, const double* const* U /* 3 host arrays, (N+1) x K column-major */);
int esdg1d_download_state(esdg1d_ctx* ctx, double* const* U);
int esdg1d_rhs(esdg1d_ctx* ctx, const double* bc12, double* const* rhs_out);          /* rhs_ESDG of the resident state */
int esdg1d_ssprk3_step(esdg1d_ctx* ctx, double dt, const double* bc12);               /* one pass of the loop body      */
int64_t esdg1d_kernel_launches(const esdg1d_ctx* ctx);

/* ---- introspection ---------------------------------------------------------- */
int64_t esdg_kernel_launches(const esdg_ctx* ctx);          /* kernels launched so far      */
int esdg_kernel_time_ms(esdg_ctx* ctx, const char* kernel, double* total_ms, int64_t* launches);
int esdg_enable_timing(esdg_ctx* ctx, int on);              /* CUDA-event timing per kernel */
/* diagnostic: the kernels' own log (x > 0, normal) and exp (|x| < 700) evaluated on n host values */
int esdg_debug_log_exp(int device, const double* x, int64_t n, double* log_out, double* exp_out);
/* diagnostic: the tensor-product kernels' own two-point flux (logmean.jl:14-28 + euler_fluxes.jl:51-89) on n
 * caller-supplied pairs.  UL, UR: [7][n] host doubles (rho, u, v, w, beta, log rho, log beta); F_out: [4][5][n] =
 * the axis-aligned form (x flux) followed by the general form contracted with e_x, e_y, e_z */
int esdg_debug_pair_flux(int device, const double* UL, const double* UR, int64_t n, double gamma, double* F_out);
/* DFMA micro-benchmark: non-tensor FP64 peak of the device in TFLOP/s (roofline denominator) */
int esdg_measure_fp64_peak(int device, double* tflops);
const char* esdg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ESDG_B200_H */
