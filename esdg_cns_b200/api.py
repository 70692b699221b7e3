"""Host-side mirror of the reference's rhs / LSRK interface over the C ABI.

The reference's boundary is the Julia call shape

    rhsQ, rhstest = rhs(Q, md::MeshData, ops, flux_fun, compute_rhstest=false)
        examples/dg2D_euler_quad.jl:141 ; dg2D_euler_tri.jl:130 ; dg3D_euler_hex.jl:167
    for INTRK = 1:5 ... resQ = rk4a*resQ + dt*rhsQ ; Q += rk4b*resQ      (:198-207)

``rhs`` / ``lsrk45!``-style functions below take the same arguments (``Q`` a list
of Nfields ``Ns x K`` arrays, ``md`` with the MeshData field names, ``ops`` the
operator bundle) and marshal them exactly as the Julia ``ccall`` shim of
INTEGRATION.md does.  All arithmetic happens in libesdg_b200.so on the GPU.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib as L

ELEM_CODE = {"tri": L.ESDG_TRI, "quad": L.ESDG_QUAD, "hex": L.ESDG_HEX}


def rk45_coeffs():
    """Carpenter-Kennedy LSRK45 (src/CommonUtils.jl:29-49); the library holds the same literals."""
    rk4a = np.array([0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                     -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0])
    rk4b = np.array([1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                     1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                     2277821191437.0 / 14882151754819.0])
    rk4c = np.array([0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                     2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0, 1.0])
    return rk4a, rk4b, rk4c


def euler_fluxes(*_a, **_k):
    """Marker for the ``flux_fun`` argument of ``rhs``: the entropy-conservative flux of
    examples/EntropyStableEuler/euler_fluxes.jl is compiled into the kernels."""
    raise NotImplementedError("euler_fluxes is evaluated on the GPU; pass it to rhs() as flux_fun")


class Solver:
    """One GPU context: operators, geometry and maps resident in HBM."""

    def __init__(self, md, ops, device=0, n_halo=0, stream=None, map_index_base=0, cns=None, affine=None):
        """``cns``: None (Euler) or a namespace with the CNS script globals ``Re, mu, lam, Pr, bctype,
        inviscid_dissp, viscous_dissp`` (dg2D_CNS_cavity_optimized.jl:21-36) and the boundary node
        sets ``wall, lid, vlid, mapB`` (0-based unless map_index_base=1; :139-147)."""
        lib = L.load()
        self._lib = lib
        self._ctx = C.c_void_p()
        dim = 3 if ops.elem == "hex" else 2
        nf = dim + 2
        self.nf = nf
        collocated = 1 if (ops.elem in ("quad", "hex") and not getattr(ops, "modal", False)) else 0
        self.cns = cns
        Nq = int(np.asarray(ops.wq).size)
        Nfq = int(np.asarray(ops.Lf).shape[1])
        Ns = int(np.asarray(ops.Ph).shape[0])
        self.Ns, self.K = Ns, int(md.K)
        # curved elements (hex script only, dg3D_euler_hex.jl:90-101, :145-151): volume geometric factors that vary
        # inside an element; the 2-D scripts read `[1,e]` whatever the mesh
        if affine is not None:
            affine = int(affine)                 # caller's claim; esdg_finalize verifies a claimed affine mesh
        elif getattr(md, "structured", False):
            affine = 1
        else:
            affine = 1
            if dim == 3 and collocated:
                scale = float(np.abs(np.asarray(md.J)).max()) ** (2.0 / 3.0)
                for n in ("rxJ", "sxJ", "txJ", "ryJ", "syJ", "tyJ", "rzJ", "szJ", "tzJ"):
                    a = np.atleast_2d(getattr(md, n))
                    if a.shape[0] > 1 and float(np.abs(a - a[:1]).max()) > 1e-10 * scale:
                        affine = 0
                        break
        self.affine = affine
        d = L.esdg_desc(dim=dim, elem=ELEM_CODE[ops.elem], N=int(ops.N), Nfields=nf, Ns=Ns, Nq=Nq, Nfq=Nfq,
                        Nfaces=int(ops.Nfaces), collocated=collocated, affine=affine, K=int(md.K), n_halo=int(n_halo),
                        gamma=1.4, volume_scale=float(ops.volume_scale),
                        lf_scale=float(ops.lf_scale) if (cns is None or cns.inviscid_dissp) else 0.0)
        self.desc = d
        code = lib.esdg_create(C.byref(d), int(device), C.byref(self._ctx))
        if code != L.ESDG_OK:
            msg = lib.esdg_last_error(self._ctx) if self._ctx else b"esdg_create failed"
            lib.esdg_destroy(self._ctx)
            self._ctx = C.c_void_p()
            raise L.EsdgError(code, (msg or b"").decode())
        self._finalizer = weakref.finalize(self, lib.esdg_destroy, self._ctx)
        if stream is not None:
            self._ck(lib.esdg_set_stream(self._ctx, C.c_void_p(int(stream))))
        # operators (Julia: Matrix(A) of the ops tuple / rd fields)
        names = ["Qrh", "Qsh", "Qth"][:dim]
        for n, A in zip(names, ops.Qskew):
            self._set_op(n, A)
        self._set_op("Ph", ops.Ph)
        self._set_op("Lf", ops.Lf)
        self._set_op("wq", np.asarray(ops.wq).reshape(-1, 1))
        if collocated:
            self._set_op("Ef", ops.Ef)
        else:
            self._set_op("Vq", ops.Vq)
            self._set_op("VhP", ops.VhP)
        # geometry + maps: host arrays (the script's MeshData), or -- structured hex meshes -- generated on the device
        if getattr(md, "structured", False):
            K3 = np.ascontiguousarray(md.Kxyz, dtype=np.int64)
            box = np.ascontiguousarray([v for ab in md.domain for v in ab], dtype=np.float64)
            per = np.ascontiguousarray([int(bool(v)) for v in md.periodic], dtype=np.int32)
            self._ck(lib.esdg_set_structured_hex(self._ctx, K3.ctypes.data, box.ctypes.data, per.ctypes.data,
                                                 int(md.zrange[0]), int(md.zrange[1] - md.zrange[0])))
        else:
            gnames = ["rxJ", "sxJ", "ryJ", "syJ", "J", "nxJ", "nyJ", "sJ"] if dim == 2 else \
                     ["rxJ", "sxJ", "txJ", "ryJ", "syJ", "tyJ", "rzJ", "szJ", "tzJ", "J", "nxJ", "nyJ", "nzJ", "sJ"]
            for n in gnames:
                a = L.f64(np.atleast_2d(getattr(md, n)))
                self._ck(lib.esdg_set_geometry(self._ctx, n.encode(), a.ctypes.data, a.shape[0], a.shape[1]))
            mapP = np.asfortranarray(np.asarray(md.mapP, dtype=np.int64) + (1 - map_index_base))
            self._ck(lib.esdg_set_maps(self._ctx, mapP.ctypes.data, mapP.size))
        if cns is not None:
            for n in ("Pq", "Vf", "Dr", "Ds"):
                self._set_op(n, getattr(ops, n))
            self._ck(lib.esdg_set_viscosity(self._ctx, float(cns.lam), float(cns.mu), float(cns.Pr), float(cns.Re),
                                            int(bool(cns.viscous_dissp))))
            shift = 1 - map_index_base
            wall = np.ascontiguousarray(np.asarray(cns.wall, dtype=np.int64) + shift)
            lid = np.ascontiguousarray(np.asarray(cns.lid, dtype=np.int64) + shift)
            vlid = np.ascontiguousarray(np.asarray(cns.vlid, dtype=np.float64))
            self._ck(lib.esdg_set_wall_bc(self._ctx, int(cns.bctype), wall.ctypes.data, wall.size, lid.ctypes.data,
                                          vlid.ctypes.data, lid.size))
            if getattr(cns, "mapB", None) is not None:
                mb = np.ascontiguousarray(np.asarray(cns.mapB, dtype=np.int64) + shift)
                self._ck(lib.esdg_set_boundary_nodes(self._ctx, mb.ctypes.data, mb.size))
            if getattr(cns, "inflow", None) is not None:
                # shock-tube node sets `leftwall` / `rightwall` and the inflow state (dg2D_CNS_modalESDG.jl:161-218)
                fin = np.ascontiguousarray(np.asarray(cns.inflow, dtype=np.int64) + shift)
                fout = np.ascontiguousarray(np.asarray(cns.outflow, dtype=np.int64) + shift)
                st = np.ascontiguousarray(np.asarray(cns.inflow_state, dtype=np.float64))
                self._ck(lib.esdg_set_inflow_outflow_bc(self._ctx, fin.ctypes.data, fin.size, st.ctypes.data,
                                                        fout.ctypes.data, fout.size))
        self._ck(lib.esdg_finalize(self._ctx))

    # -- helpers ---------------------------------------------------------------
    def _ck(self, code):
        L.check(self._ctx, code)

    def _set_op(self, name, A):
        a = L.f64(np.atleast_2d(A))
        self._ck(self._lib.esdg_set_operator(self._ctx, name.encode(), a.ctypes.data, a.shape[0], a.shape[1]))

    def _host_fields(self, Q):
        arrs = [L.f64(q) for q in Q]
        if len(arrs) != self.nf or any(a.shape != (self.Ns, self.K) for a in arrs):
            raise ValueError(f"Q must be {self.nf} arrays of shape ({self.Ns},{self.K})")
        return arrs

    # -- hot path ----------------------------------------------------------------
    def rhs(self, Q=None, compute_rhstest=False):
        """Drop-in for ``rhs(Q,md,ops,flux,compute_rhstest)``: returns (rhsQ, rhstest)."""
        out = [np.empty((self.Ns, self.K), order="F") for _ in range(self.nf)]
        rt = (C.c_double * 2)(0.0, 0.0)
        if Q is None:
            qp = None
        else:
            qin = self._host_fields(Q)
            qp = L.ptr_array(qin)
        self._ck(self._lib.esdg_rhs(self._ctx, qp, L.ptr_array(out), int(bool(compute_rhstest)), rt))
        if self.cns is not None:
            return out, float(rt[0]), float(rt[1])
        return out, float(rt[0])

    def upload(self, Q):
        qin = self._host_fields(Q)
        self._ck(self._lib.esdg_upload_state(self._ctx, L.ptr_array(qin)))
        self._ck(self._lib.esdg_sync(self._ctx))

    def download(self, out=None):
        if out is None:
            out = [np.empty((self.Ns, self.K), order="F") for _ in range(self.nf)]
        else:
            # raw pointers cross the ABI: the caller's arrays must be exactly what the library writes
            # (Ns x K float64, column-major); a copy made here would never reach the caller
            if len(out) != self.nf:
                raise ValueError(f"download: expected {self.nf} field arrays")
            for a in out:
                if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.shape == (self.Ns, self.K)
                        and a.flags.f_contiguous and a.flags.writeable):
                    raise ValueError(f"download: every output must be a writeable float64 array of shape ({self.Ns}, {self.K}) "
                                     "in column-major (Fortran) order")
        self._ck(self._lib.esdg_download_state(self._ctx, L.ptr_array(out)))
        return out

    def lsrk45(self, dt, nsteps=1, want_rhstest=False):
        """``nsteps`` LSRK45 steps on the resident state (dg2D_euler_quad.jl:196-207)."""
        if want_rhstest:
            rt = C.c_double(0.0)
            self._ck(self._lib.esdg_lsrk45_step(self._ctx, float(dt), int(nsteps), C.byref(rt)))
            return float(rt.value)
        self._ck(self._lib.esdg_lsrk45_step(self._ctx, float(dt), int(nsteps), None))
        return None

    def lsrk45_host(self, Q, dt, nsteps=1, out=None):
        """``nsteps`` LSRK45 steps of the HOST state ``Q`` (list of Ns x K column-major arrays); the result is
        written to ``out`` (default: in place).  Upload and download are pipelined with the first / last stage."""
        qin = self._host_fields(Q)
        qout = qin if out is None else self._host_fields(out)
        if any(a is not b for a, b in zip(qin, Q)) and out is None:
            raise ValueError("lsrk45_host works in place: pass column-major float64 arrays (or give out=)")
        self._ck(self._lib.esdg_lsrk45_host(self._ctx, L.ptr_array(qin), L.ptr_array(qout), float(dt), int(nsteps)))
        return qout

    def dopri45(self, T, dt, err_tol=1e-5, max_attempts=10 ** 9):
        """The adaptive time loop of dg2D_CNS_cavity_optimized.jl:974-1053 on the resident state:
        ``while t < T`` DOPRI45 steps with FSAL and the PI controller.  Returns the histories the
        script keeps (``thist, dthist, errhist, rhstesthist, vischist``) as a dict."""
        self._ck(self._lib.esdg_dopri45_init(self._ctx))
        t, h, dt0 = C.c_double(0.0), C.c_double(float(dt)), float(dt)
        ok, err, rt = C.c_int(0), C.c_double(0.0), (C.c_double * 2)(0.0, 0.0)
        hist = {"t": [0.0], "dt": [float(dt)], "err": [0.0], "rhstest": [0.0], "rhstest_visc": [0.0], "accepted": []}
        n = 0
        while t.value < T and n < max_attempts:
            self._ck(self._lib.esdg_dopri45_step(self._ctx, C.byref(t), C.byref(h), dt0, float(err_tol), C.byref(ok),
                                                 C.byref(err), rt))
            n += 1
            hist["t"].append(t.value); hist["dt"].append(h.value); hist["err"].append(err.value)
            hist["rhstest"].append(rt[0]); hist["rhstest_visc"].append(rt[1]); hist["accepted"].append(bool(ok.value))
        return hist

    def sync(self):
        self._ck(self._lib.esdg_sync(self._ctx))

    def set_stream(self, cuda_stream):
        """Run every kernel of this context on the given CUDA stream (esdg_set_stream)."""
        self._ck(self._lib.esdg_set_stream(self._ctx, C.c_void_p(int(cuda_stream))))

    # -- multi-GPU plumbing --------------------------------------------------------
    def set_send_list(self, idx0):
        idx = np.ascontiguousarray(np.asarray(idx0, dtype=np.int64) + 1)
        self._ck(self._lib.esdg_set_send_list(self._ctx, idx.ctypes.data, idx.size))

    def comm_init(self, unique_id, nranks, rank, nb_rank, nb_send, nb_recv):
        """Attach an NCCL communicator (esdg_comm_init): afterwards ``lsrk45`` steps the partitioned mesh by itself."""
        r = np.ascontiguousarray(nb_rank, dtype=np.int32)
        ns = np.ascontiguousarray(nb_send, dtype=np.int64)
        nr = np.ascontiguousarray(nb_recv, dtype=np.int64)
        uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        assert uid.size == 128
        self._ck(self._lib.esdg_comm_init(self._ctx, uid.ctypes.data, int(nranks), int(rank), int(r.size),
                                          r.ctypes.data, ns.ctypes.data, nr.ctypes.data))

    def compute_traces(self):
        self._ck(self._lib.esdg_compute_traces(self._ctx))

    def pack_send(self, dev_ptr):
        self._ck(self._lib.esdg_pack_send(self._ctx, C.c_void_p(int(dev_ptr))))

    def halo_ptr(self):
        p = C.c_void_p()
        s = C.c_int64()
        self._ck(self._lib.esdg_halo_ptr(self._ctx, C.byref(p), C.byref(s)))
        return int(p.value), int(s.value)

    def unpack_halo(self, dev_ptr):
        self._ck(self._lib.esdg_unpack_halo(self._ctx, C.c_void_p(int(dev_ptr))))

    PART_ALL, PART_BOUNDARY, PART_INTERIOR = 0, 1, 2

    def stage_update(self, stage, dt, compute_rhstest=False, part=0):
        self._ck(self._lib.esdg_stage_update(self._ctx, int(stage), float(dt), int(bool(compute_rhstest)), int(part)))

    def trace_planes(self):
        return int(self._lib.esdg_trace_planes(self._ctx))

    def kernel_path(self):
        return self._lib.esdg_kernel_path(self._ctx).decode()

    def hex_diag(self):
        return self.kernel_path() == "hex_line_diag"

    # CNS across GPUs: the sigma traces are the second exchange of a stage (include/esdg_b200.h)
    SIGMA_PLANES = 6

    def cns_sigma_pass(self, compute_rhstest=False):
        self._ck(self._lib.esdg_cns_sigma_pass(self._ctx, int(bool(compute_rhstest))))

    def pack_send_sigma(self, dev_ptr):
        self._ck(self._lib.esdg_pack_send_sigma(self._ctx, C.c_void_p(int(dev_ptr))))

    def unpack_halo_sigma(self, dev_ptr):
        self._ck(self._lib.esdg_unpack_halo_sigma(self._ctx, C.c_void_p(int(dev_ptr))))

    def read_rhstest(self):
        rt = C.c_double(0.0)
        self._ck(self._lib.esdg_read_rhstest(self._ctx, C.byref(rt)))
        return float(rt.value)

    def read_rhstest_visc(self):
        rt = C.c_double(0.0)
        self._ck(self._lib.esdg_read_rhstest_visc(self._ctx, C.byref(rt)))
        return float(rt.value)

    # -- introspection ---------------------------------------------------------------
    def kernel_launches(self):
        return int(self._lib.esdg_kernel_launches(self._ctx))

    def enable_timing(self, on=True):
        self._ck(self._lib.esdg_enable_timing(self._ctx, int(on)))

    def kernel_time_ms(self, name):
        t = C.c_double()
        n = C.c_int64()
        self._ck(self._lib.esdg_kernel_time_ms(self._ctx, name.encode(), C.byref(t), C.byref(n)))
        return float(t.value), int(n.value)

    def close(self):
        self._finalizer()


_solver_cache = []   # [(md, ops, Solver)], most recent last; the functional API keeps 4 contexts alive


def _cached_solver(md, ops, cns=None):
    for ent in _solver_cache:
        if ent[0] is md and ent[1] is ops and ent[2].cns is cns:     # exact match, None included: rhs() after rhsRK()
            return ent[2]                                            # on the same (md, ops) must not get the CNS context
    ent = (md, ops, Solver(md, ops, cns=cns))
    _solver_cache.append(ent)
    while len(_solver_cache) > 4:
        _solver_cache.pop(0)[2].close()
    return ent[2]


def rhs(Q, md, ops, flux_fun=euler_fluxes, compute_rhstest=False):
    """``rhsQ, rhstest = rhs(Q, md, ops, flux_fun, compute_rhstest)`` on the GPU.

    Same call shape as the reference scripts; ``flux_fun`` must be ``euler_fluxes``
    (the only flux the reference passes).
    """
    if flux_fun is not euler_fluxes:
        raise NotImplementedError("only the entropy-conservative euler_fluxes is compiled into the kernels")
    return _cached_solver(md, ops).rhs(Q, compute_rhstest)


def lsrk45_steps(Q, md, ops, dt, nsteps=1, want_rhstest=False):
    """The reference time loop (dg2D_euler_quad.jl:196-212) for ``nsteps`` steps; returns (Q, rhstest)."""
    s = _cached_solver(md, ops)
    s.upload(Q)
    rt = s.lsrk45(dt, nsteps, want_rhstest)
    return s.download(), rt


def rhsRK(Q, rd, md, Re, BCTYPE, ops, flux_fun, inviscid_dissp, viscous_dissp, cns):
    """``rhsQ, rhstest, rhstest_visc = rhsRK!(Q, rd, md, Re, BCTYPE, ops, flux, inviscid_dissp,
    viscous_dissp, <work arrays>)`` of examples/CompressibleNS/dg2D_CNS_cavity_optimized.jl:955-972.

    ``cns`` carries what the reference script keeps in globals / closures: ``mu, lam, Pr``
    (:33-36, as passed to init_visc_fxn) and the wall node sets ``wall, lid, vlid, mapB`` (:139-147).
    ``ops`` additionally holds rd's ``Pq, Vf, Dr, Ds`` (the reference unpacks them from ``rd``).
    """
    if flux_fun is not euler_fluxes:
        raise NotImplementedError("only the entropy-conservative euler_fluxes is compiled into the kernels")
    if (cns.Re, cns.bctype, bool(cns.inviscid_dissp), bool(cns.viscous_dissp)) != \
            (Re, BCTYPE, bool(inviscid_dissp), bool(viscous_dissp)):
        raise ValueError("rhsRK: Re/BCTYPE/dissipation flags disagree with the cns settings")
    for n in ("Pq", "Vf", "Dr", "Ds"):
        if not hasattr(ops, n):
            setattr(ops, n, getattr(rd, n))
    return _cached_solver(md, ops, cns).rhs(Q, True)
