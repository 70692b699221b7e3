"""Face-trace halo exchange for the z-slab element partition.

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests) is the plumbing.  Per stage each rank sends the entropy-projected traces of its bottom
layer's t=-1 faces to the slab below and of its top layer's t=+1 faces to the slab above, and
receives the matching layers into the halo tail that ``mapP`` points at
(``setup_dg.uniform_hex_mesh_md(zslab=...)``: lower neighbour's block first, then the upper's).
"""
import numpy as np


def attach_comm(solver, md, dist, torch, device):
    """Give a slab context its own NCCL communicator (include/esdg_b200.h: esdg_comm_*).  ``torch.distributed`` only
    carries the 128-byte unique id from rank 0 to the others; the halo exchange itself -- pack, ncclSend / ncclRecv
    straight into the halo tail, overlap with the interior elements, all-reduce of rhstest -- then runs inside
    ``esdg_lsrk45_step`` (``Solver.lsrk45``).  A Julia host does the same with MPI.jl / Distributed for the id."""
    from . import _lib as L
    import ctypes as C
    send = np.concatenate([md.send_lower, md.send_upper])
    solver.set_send_list(send)
    uid = torch.zeros(128, dtype=torch.uint8)
    if dist.get_rank() == 0:
        buf = (C.c_ubyte * 128)()
        rc = L.load().esdg_comm_unique_id(buf)
        if rc:
            raise L.EsdgError(rc, "esdg_comm_unique_id failed (libnccl.so.2 not loadable?)")
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    uid = uid.to(device)
    dist.broadcast(uid, 0)
    half = int(md.n_halo) // 2
    # neighbour order = block order in the send list and in the halo tail: lower first, then upper (-1: no partner)
    nb = [-1 if md.lower_rank is None else md.lower_rank, -1 if md.upper_rank is None else md.upper_rank]
    ns = [md.send_lower.size, md.send_upper.size]
    nr = [half, half]
    solver.comm_init(uid.cpu().numpy().tobytes(), dist.get_world_size(), dist.get_rank(), nb, ns, nr)


def _share_stream(solver, torch):
    """The host-driven exchange mixes the context's kernels (pack / unpack on the CONTEXT's stream) with tensor copies
    and NCCL calls on torch's current stream: both must be the same stream, or the copies can read the send buffer
    before the pack kernel has written it.  (The in-library exchange, attach_comm, orders its own streams.)"""
    solver.set_stream(torch.cuda.current_stream().cuda_stream)


class SlabHalo:
    def __init__(self, md, nf, dist, torch, device):
        self.dist, self.torch = dist, torch
        self.n_lo, self.n_up = int(md.send_lower.size), int(md.send_upper.size)
        self.n_send = self.n_lo + self.n_up
        self.n_halo = int(md.n_halo)
        self.send_list = np.concatenate([md.send_lower, md.send_upper])   # local trace points, 0-based
        self.lower, self.upper = md.lower_rank, md.upper_rank
        kw = dict(dtype=torch.float64, device=device)
        self.send = torch.empty((nf, self.n_send), **kw)
        self.recv = torch.zeros((nf, self.n_halo), **kw)
        self._s_lo = torch.empty((nf, self.n_lo), **kw)
        self._s_up = torch.empty((nf, self.n_up), **kw)
        self._r_lo = torch.empty((nf, self.n_halo // 2), **kw)
        self._r_up = torch.empty((nf, self.n_halo // 2), **kw)

    def exchange(self):
        """self.send (filled by the caller) -> self.recv."""
        return self.finish(self.start())

    def start(self):
        dist = self.dist
        self._s_lo.copy_(self.send[:, :self.n_lo])
        self._s_up.copy_(self.send[:, self.n_lo:])
        ops = []
        # Sends are posted (to lower, to upper) and receives (from upper, from lower): with two ranks
        # and a periodic z direction both messages travel between the same pair and match in order.
        if self.lower is not None:
            ops.append(dist.P2POp(dist.isend, self._s_lo, self.lower))
        if self.upper is not None:
            ops.append(dist.P2POp(dist.isend, self._s_up, self.upper))
        if self.upper is not None:
            ops.append(dist.P2POp(dist.irecv, self._r_up, self.upper))
        if self.lower is not None:
            ops.append(dist.P2POp(dist.irecv, self._r_lo, self.lower))
        return dist.batch_isend_irecv(ops) if ops else []

    def finish(self, works):
        for w in works:
            w.wait()
        h = self.n_halo // 2
        self.recv[:, :h].copy_(self._r_lo)
        self.recv[:, h:].copy_(self._r_up)
        return self.recv


class GpuSlabStepper:
    """LSRK45 on a z-slab.  Per stage: boundary-layer elements first, their new traces are packed and
    the NCCL exchange starts while the interior elements are still being updated; the received halo
    is unpacked before the next stage."""

    def __init__(self, solver, md, dist, torch, device, overlap=True):
        self.s = solver
        _share_stream(solver, torch)
        self.halo = SlabHalo(md, solver.trace_planes(), dist, torch, device)
        solver.set_send_list(self.halo.send_list)
        self.overlap = overlap
        self.fresh = True          # state uploaded, traces not yet computed / exchanged

    def _exchange(self):
        self.s.pack_send(self.halo.send.data_ptr())
        self.halo.exchange()
        self.s.unpack_halo(self.halo.recv.data_ptr())

    def mark_fresh(self):
        self.fresh = True

    def step(self, dt, want_rhstest=False):
        s = self.s
        if self.fresh:
            s.compute_traces()
            self._exchange()
            self.fresh = False
        for k in range(5):
            want = want_rhstest and k == 4
            if self.overlap:
                s.stage_update(k, dt, want, s.PART_BOUNDARY)
                s.pack_send(self.halo.send.data_ptr())
                works = self.halo.start()
                s.stage_update(k, dt, want, s.PART_INTERIOR)
                self.halo.finish(works)
                s.unpack_halo(self.halo.recv.data_ptr())
            else:
                s.stage_update(k, dt, want, s.PART_ALL)
                self._exchange()


class GpuSlabStepperCNS:
    """LSRK45 of the compressible Navier-Stokes rhs (rhsRK!, dg2D_CNS_cavity_optimized.jl:955-972) on a slab
    of elements.  rhsRK! has two inter-element dependences, so a stage has two exchanges:
    the projected entropy-variable traces (4 planes; they serve the inviscid flux, the BR1 gradient and the
    penalisation) and the stress traces of the BR1 divergence (6 planes)."""

    def __init__(self, solver, md, dist, torch, device):
        self.s = solver
        _share_stream(solver, torch)
        self.halo_v = SlabHalo(md, solver.trace_planes(), dist, torch, device)
        self.halo_s = SlabHalo(md, solver.SIGMA_PLANES, dist, torch, device)
        solver.set_send_list(self.halo_v.send_list)
        self.fresh = True

    def mark_fresh(self):
        self.fresh = True

    def _exchange_v(self):
        self.s.pack_send(self.halo_v.send.data_ptr())
        self.halo_v.exchange()
        self.s.unpack_halo(self.halo_v.recv.data_ptr())

    def _exchange_sigma(self):
        self.s.pack_send_sigma(self.halo_s.send.data_ptr())
        self.halo_s.exchange()
        self.s.unpack_halo_sigma(self.halo_s.recv.data_ptr())

    def step(self, dt, want_rhstest=False):
        s = self.s
        if self.fresh:
            s.compute_traces()
            self._exchange_v()
            self.fresh = False
        for k in range(5):
            want = want_rhstest and k == 4
            s.cns_sigma_pass(want)
            self._exchange_sigma()
            s.stage_update(k, dt, want, s.PART_ALL)   # rhs pass + entropy-variable traces of the new state
            self._exchange_v()


def exchange_in_process(halos):
    """The exchange of ``SlabHalo.exchange`` for slabs that live in ONE process (contexts on the same GPU or
    on several GPUs driven by one host thread): ``halos[r]`` is rank r's SlabHalo with ``send`` filled;
    device-to-device copies take the place of the NCCL send/recv pairs.  All contexts must share the stream
    (or be synchronised by the caller)."""
    for r, h in enumerate(halos):
        half = h.n_halo // 2
        if h.lower is not None:
            lo = halos[h.lower]
            h.recv[:, :half].copy_(lo.send[:, lo.n_lo:])       # the lower slab's send_upper block
        if h.upper is not None:
            up = halos[h.upper]
            h.recv[:, half:].copy_(up.send[:, :up.n_lo])       # the upper slab's send_lower block
    return [h.recv for h in halos]


class LocalSlabsCNS:
    """GpuSlabStepperCNS for several slabs held by one process (same call sequence per stage; the two
    exchanges are device-to-device copies).  Used by the single-GPU test of the distributed CNS path."""

    def __init__(self, solvers, mds, torch, device):
        self.s = solvers
        self.hv = [SlabHalo(md, s.trace_planes(), None, torch, device) for s, md in zip(solvers, mds)]
        self.hs = [SlabHalo(md, s.SIGMA_PLANES, None, torch, device) for s, md in zip(solvers, mds)]
        for s, h in zip(solvers, self.hv):
            s.set_send_list(h.send_list)
        self.fresh = True

    def _exchange_v(self):
        for s, h in zip(self.s, self.hv):
            s.pack_send(h.send.data_ptr())
        exchange_in_process(self.hv)
        for s, h in zip(self.s, self.hv):
            s.unpack_halo(h.recv.data_ptr())

    def _exchange_sigma(self):
        for s, h in zip(self.s, self.hs):
            s.pack_send_sigma(h.send.data_ptr())
        exchange_in_process(self.hs)
        for s, h in zip(self.s, self.hs):
            s.unpack_halo_sigma(h.recv.data_ptr())

    def step(self, dt, want_rhstest=False):
        if self.fresh:
            for s in self.s:
                s.compute_traces()
            self._exchange_v()
            self.fresh = False
        for k in range(5):
            want = want_rhstest and k == 4
            for s in self.s:
                s.cns_sigma_pass(want)
            self._exchange_sigma()
            for s in self.s:
                s.stage_update(k, dt, want, s.PART_ALL)
            self._exchange_v()


class LocalSlabs:
    """GpuSlabStepper for several z-slabs of an Euler mesh held by one process: the same per-stage call sequence
    (boundary layers -> pack -> exchange -> interior -> unpack, or the plain one), with device-to-device copies
    for the exchange.  Single-GPU coverage of the distributed hex path (esdg_stage_update parts 1/2, the
    7-plane halo, pack / unpack)."""

    def __init__(self, solvers, mds, torch, device, overlap=True):
        self.s = solvers
        self.h = [SlabHalo(md, s.trace_planes(), None, torch, device) for s, md in zip(solvers, mds)]
        for s, h in zip(solvers, self.h):
            s.set_send_list(h.send_list)
        self.overlap = overlap
        self.fresh = True

    def _pack(self):
        for s, h in zip(self.s, self.h):
            s.pack_send(h.send.data_ptr())

    def _unpack(self):
        exchange_in_process(self.h)
        for s, h in zip(self.s, self.h):
            s.unpack_halo(h.recv.data_ptr())

    def step(self, dt, want_rhstest=False):
        if self.fresh:
            for s in self.s:
                s.compute_traces()
            self._pack()
            self._unpack()
            self.fresh = False
        for k in range(5):
            want = want_rhstest and k == 4
            if self.overlap:
                for s in self.s:
                    s.stage_update(k, dt, want, s.PART_BOUNDARY)
                self._pack()
                for s in self.s:
                    s.stage_update(k, dt, want, s.PART_INTERIOR)
                self._unpack()
            else:
                for s in self.s:
                    s.stage_update(k, dt, want, s.PART_ALL)
                self._pack()
                self._unpack()
