#!/usr/bin/env python
"""bench.py -- FP64 DOF*RK-stage updates/s of the ESDG rhs + LSRK45 hot path on B200.

Workload (BASELINE.json config 5, the one the north-star target is quoted on): 3D Taylor-Green
vortex, entropy-stable flux differencing on hexes, N=4, element-partitioned in z-slabs of
128 x 128 x 16 elements per GPU (weak scaling: 8 GPUs = the full 128^3 mesh).  One "step" is
one LSRK45 time step = 5 fused rhs+stage updates of every local DOF.

    python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun)
    python bench.py --impl reference --steps K --warmup W    (reference algorithm on the host CPU)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FP64 DOF*RK-stage updates/sec (3D Taylor-Green, hex N=4, entropy-stable flux differencing + LSRK45)"
UNIT = "DOF*stage/s"
N_DEG = 4
LOCAL_MESH = (128, 128, 16)       # elements per GPU
CPU_SAMPLE_MESH = (24, 24, 24)    # bounded CPU sample of the same workload
CFL = .5


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes_per_element_stage(N, nf=5):
    from esdg_cns_b200 import roofline
    return roofline.alg_bytes("hex", N)


def alg_flops_per_element_stage(N):
    from esdg_cns_b200 import roofline
    return roofline.alg_flops("hex", N)


def measured_traffic(kernel, K):
    """(dram bytes per launch, provenance) from the committed ncu capture; None when the capture is stale."""
    from esdg_cns_b200 import roofline
    return roofline.measured_traffic(kernel, K)


def tgv_dt(N, K1D_global):
    CN = (N + 1) * (N + 2) * 3 / 2            # dg3D_euler_hex.jl:115-117
    return CFL * 2 / (CN * K1D_global)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference ALGORITHM (sparse-row Hadamard sum evaluating (i,j) and (j,i), unfused passes,
    as written in examples/dg3D_euler_hex.jl:122-222) restated in C + OpenMP, on the host cores.
    Julia is not installed in this image, so this is a port (kind='port'), not the Julia binary."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle, physics as P, solvers as SV   # the one other place bench may execute oracle/
    from esdg_cns_b200 import setup_dg as SD
    N = N_DEG
    mesh = CPU_SAMPLE_MESH
    prd, pops = SD.reference_hex(N)
    pmd = SD.uniform_hex_mesh_md(prd, mesh, domain=((-np.pi, np.pi),) * 3)
    x, y, z = SD.hex_node_coords(prd, pmd)
    rho, u, v, w, p = SD.taylor_green(x, y, z)
    Q = SD.primitive_to_conservative(rho, (u, v, w), p)
    _all_host_threads(c_oracle)
    co = c_oracle.COracle(pmd, pops)
    dt = tgv_dt(N, 128)
    dof = co.nf * co.Nq * co.K
    for _ in range(args.warmup):
        co.lsrk45(Q, dt, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        Q, _ = co.lsrk45(Q, dt, 1)
    el = time.perf_counter() - t0
    val = dof * 5 * args.steps / el
    sample = f"{mesh[0]}x{mesh[1]}x{mesh[2]} elements (K={co.K}) of the same TGV hex N={N} workload, {args.steps} LSRK45 steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"tgv_hex_N{N} bounded CPU sample {mesh[0]}x{mesh[1]}x{mesh[2]}", "N": N,
                   "note": "reference algorithm (dg3D_euler_hex.jl rhs + LSRK45) restated in C/OpenMP; Julia unavailable"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": co.threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _all_host_threads(c_oracle):
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host thread it can."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    c_oracle.load().esdg_oracle_set_threads(int(n))


def cpu_baseline_leg(seconds_budget=25.0):
    from oracle import c_oracle
    from esdg_cns_b200 import setup_dg as SD
    _all_host_threads(c_oracle)
    N, mesh = N_DEG, CPU_SAMPLE_MESH
    prd, pops = SD.reference_hex(N)
    pmd = SD.uniform_hex_mesh_md(prd, mesh, domain=((-np.pi, np.pi),) * 3)
    x, y, z = SD.hex_node_coords(prd, pmd)
    Q = SD.primitive_to_conservative(*(lambda r, u, v, w, p: (r, (u, v, w), p))(*SD.taylor_green(x, y, z)))
    co = c_oracle.COracle(pmd, pops)
    dt = tgv_dt(N, 128)
    t0 = time.perf_counter()
    steps = 0
    while True:
        Q, _ = co.lsrk45(Q, dt, 1)
        steps += 1
        el = time.perf_counter() - t0
        if el > seconds_budget / 2 or steps >= 4:
            break
    dof = co.nf * co.Nq * co.K
    return {"value": dof * 5 * steps / el, "unit": UNIT, "cores": co.threads, "kind": "port",
            "sample": f"{mesh[0]}x{mesh[1]}x{mesh[2]} elements (K={co.K}) of the same TGV hex N={N} workload, "
                      f"{steps} LSRK45 steps, {el:.1f} s, C/OpenMP restatement of dg3D_euler_hex.jl (Julia unavailable)"}


def bind_to_gpu_cpus(index):
    """Multi-GPU runs: keep this rank (and the pinned host buffers it first-touches) on the CPUs NVML reports as local to
    its GPU, so that the host<->device copies of the end-to-end leg do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------- the other BASELINE configs
def other_configs(device, peak_gbs):
    """BASELINE configs 4, 1, 2 (operators at timing scale) and 3, plus config 5's element on a curved mesh, timed in the
    same process (a few seconds each), so that
    their numbers are driver-visible: DOF*stage/s with the state resident, ms per LSRK stage from CUDA events around
    the kernels, ALGORITHMIC bytes per element-stage (esdg_cns_b200/roofline.py) and the fraction of the measured HBM
    peak they amount to."""
    from types import SimpleNamespace
    from esdg_cns_b200 import api, setup_dg as SD, roofline as R
    out = []

    def run(name, solver, Q, dt, nf, Ns, Nfq, d, kernels, K, cns=False, nsteps=4, B=None):
        solver.upload(Q)
        solver.lsrk45(dt, 2)
        solver.sync()
        solver.enable_timing(True)
        solver.lsrk45(dt, nsteps)
        solver.sync()
        ms = sum(solver.kernel_time_ms(k)[0] for k in kernels)
        finite = all(np.isfinite(q).all() for q in solver.download())
        solver.close()
        per_stage = ms / (5 * nsteps)
        B = B if B is not None else R.alg_bytes_sizes(nf, Ns, Nfq, d, cns)
        out.append({"config": name, "elements": int(K), "dof": int(nf * Ns * K), "ms_per_stage": per_stage,
                    "value": nf * Ns * K / (per_stage * 1e-3), "unit": UNIT, "alg_bytes_per_element_stage": B,
                    "achieved_gbs": B * K / (per_stage * 1e-3) / 1e9, "frac": B * K / (per_stage * 1e-3) / 1e9 / peak_gbs,
                    "kernels": list(kernels), "finite": bool(finite)})

    # config 4: hex N=3, 64^3, isentropic vortex extruded in z
    N = 3
    prd, pops = SD.reference_hex(N)
    md = SD.structured_hex_md(prd, (64, 64, 64), domain=((0, 20), (-5, 5), (-5, 5)))
    x, y, z = SD.hex_node_coords(prd, md)
    rho, u, v, p = SD.vortex(x, y)
    Q = SD.primitive_to_conservative(rho, (u, v, 0 * u), p)
    run("4: hex N=3 64^3 vortex (tensor-product kernels)", api.Solver(md, pops, device=device), Q, tgv_dt(N, 64) * 5 / np.pi, 5,
        (N + 1) ** 3, 6 * (N + 1) ** 2, 3, ["hex_stage"], md.K)
    del md, Q, x, y, z
    # config 5's element on a CURVED mesh: hex N=4, 32^3 trilinear elements, geometric factors at every
    # hybridized point averaged per pair (dg3D_euler_hex.jl:145-151)
    prd4, pops4 = SD.reference_hex(4)
    md = SD.curved_hex_mesh_md(prd4, (32, 32, 32), domain=((-np.pi, np.pi),) * 3, amp=.1)
    rho, u, v, w, p = SD.taylor_green(md.xq, md.yq, md.zq)
    Q = SD.primitive_to_conservative(rho, (u, v, w), p)
    sc = api.Solver(md, pops4, device=device)
    run("5c: hex N=4 32^3 curved elements (%s)" % sc.kernel_path(), sc, Q, tgv_dt(4, 32), 5, 125, 150, 3, ["hex_stage"], md.K,
        B=R.alg_bytes_curved_hex(4))
    del md, Q, rho, u, v, w, p
    # config 1 operators at timing scale: collocated quads N=3, 1024^2, periodic vortex
    rd, ops = SD.reference_quad(N)
    md = SD.uniform_quad_mesh_md(rd, 1024, 1024, domain=((0, 20), (-5, 5)))
    rho, u, v, p = SD.vortex(md.xq, md.yq)
    Q = SD.primitive_to_conservative(rho, (u, v), p)
    sq = api.Solver(md, ops, device=device)
    tp = sq.kernel_path() == "quad_line"
    run("1: quad N=3 1024^2 vortex (%s kernels)" % ("tensor-product" if tp else "generic"), sq, Q, 1e-4, 4, (N + 1) ** 2, 4 * (N + 1), 2,
        ["quad_stage"] if tp else ["generic_rhs", "generic_traces"], md.K)
    del md, Q
    # config 2 at timing scale: modal triangles N=4, 2 x 256^2 elements, periodic vortex (dense hybridized operators: table-driven
    # kernels with DMMA batch GEMMs and the symmetric pair schedule)
    rdt, opst = SD.reference_tri(4)
    md = SD.uniform_tri_mesh_md(rdt, 256, 256, domain=((0, 20), (-5, 5)))
    rho, u, v, p = SD.vortex(md.x, md.y)
    Q = SD.primitive_to_conservative(rho, (u, v), p)
    run("2: tri N=4 2x256^2 vortex (modal, generic kernels)", api.Solver(md, opst, device=device), Q, 1e-4, 4, rdt.r.size, rdt.rf.size, 2,
        ["generic_rhs", "generic_traces"], md.K)
    del md, Q
    # config 3: CNS lid-driven cavity, modal quads N=3, 512^2, BCTYPE 1, both dissipations on
    rd, ops = SD.reference_quad_modal(N)
    md = SD.uniform_quad_mesh_md(rd, 512, 512, periodic=(False, False))
    bc = SD.cavity_bc_sets(md)
    Re = 1000.0
    st = SimpleNamespace(Re=Re, mu=1 / Re, lam=-2 / 3 / Re, Pr=.71, bctype=1, inviscid_dissp=True, viscous_dissp=True,
                         wall=bc.wall, lid=bc.lid, vlid=bc.vlid, mapB=bc.mapB)
    one = np.ones_like(md.x)
    Q = SD.primitive_to_conservative(one, (0 * one, 0 * one), one / (.3 ** 2 * 1.4))
    run("3: CNS quad N=3 512^2 cavity (BR1, wall BCs)", api.Solver(md, ops, cns=st, device=device), Q, 1e-5, 4, rd.r.size,
        rd.rf.size, 2, ["cns_traces", "cns_sigma", "cns_rhs"], md.K, cns=True)
    return out


# ------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from esdg_cns_b200 import api, halo as H, setup_dg as SD

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None   # pinned host buffers on the GPU's own NUMA node
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=device)
    N = N_DEG
    Kx, Ky, Kzl = LOCAL_MESH
    if args.small:
        Kx, Ky, Kzl = 16, 16, 8
    Kz = Kzl * world
    prd, pops = SD.reference_hex(N)
    # always the periodic box [-pi,pi]^3 so the Taylor-Green field is smooth across the periodic wrap;
    # with 8 GPUs the mesh is the 128^3 of config 5, with fewer the elements are stretched in z
    dom = ((-np.pi, np.pi),) * 3
    # maps and geometric factors are generated on the device (esdg_set_structured_hex): no 1.3 GB of Int64 mapP per GPU
    md = SD.structured_hex_md(prd, (Kx, Ky, Kz), domain=dom, zslab=(rank, world) if world > 1 else None)
    x, y, z = SD.hex_node_coords(prd, md)
    rho, u, v, w, p = SD.taylor_green(x, y, z)
    Q0 = SD.primitive_to_conservative(rho, (u, v, w), p)
    del x, y, z, rho, u, v, w, p
    nf, Ns, K = 5, (N + 1) ** 3, md.K
    dof = nf * Ns * K
    dt = tgv_dt(N, Kx)

    stream = torch.cuda.current_stream()
    s = api.Solver(md, pops, device=local_rank, n_halo=md.n_halo, stream=stream.cuda_stream)
    # pinned host buffers (the user's arrays for the end-to-end leg); at N = 1 the process is bound to the GPU's own CPUs only
    # while they are allocated and first touched (NUMA placement), so that the CPU baseline keeps every host core
    aff0 = os.sched_getaffinity(0) if world == 1 else None
    if world == 1:
        numa = bind_to_gpu_cpus(local_rank)
    pin = [torch.zeros((K, Ns), dtype=torch.float64).pin_memory() for _ in range(nf)]
    if aff0 is not None:
        os.sched_setaffinity(0, aff0)
    host = [t.numpy().T for t in pin]                   # Ns x K column-major views
    for h, q in zip(host, Q0):
        h[...] = q
    del Q0
    s.upload(host)
    # N > 1: the context owns its NCCL communicator (esdg_comm_init) and esdg_lsrk45_step does the halo exchange
    # itself; --py-halo keeps the host-driven exchange (torch.distributed p2p + staging copies) for comparison
    stepper = None
    if world > 1:
        if args.py_halo:
            stepper = H.GpuSlabStepper(s, md, dist, torch, device, overlap=not args.no_overlap)
        else:
            if args.no_overlap:
                os.environ["ESDG_COMM_NO_OVERLAP"] = "1"
            H.attach_comm(s, md, dist, torch, device)

    def step():
        if stepper is None:
            s.lsrk45(dt, 1)
        else:
            stepper.step(dt)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()          # started before the warm-up: the first nvidia-smi call on a fresh box takes ~1 s
    for _ in range(max(args.warmup, 3)):
        step()
    s.sync()
    if sampler:
        sampler.rows.clear()     # keep only samples taken under the timed load
    # ---- device-resident timing ---------------------------------------------------------
    s.enable_timing(True)
    l0 = s.kernel_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = s.kernel_launches() - l0
    s.enable_timing(False)
    ktimes = {k: s.kernel_time_ms(k) for k in ("hex_stage", "hex_traces", "generic_rhs", "generic_traces", "pack_send")}
    # ---- end to end through the host-buffer API ---------------------------------------------
    def e2e_step():
        if world == 1:
            # one call of the host-buffer API: H2D of this step's input from pinned memory, 5 stages, D2H of the
            # result -- all inside the timed region; the library runs the stages as a wavefront over element chunks
            # behind the upload and starts each chunk's download behind its last stage
            s.lsrk45_host(host, dt, 1)
            return
        s.upload(host)           # H2D of this step's input from pinned memory (+ sync)
        if stepper is not None:
            stepper.mark_fresh()
        step()
        s.download(host)         # D2H of the step's result (synchronous)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 10))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None

    tmax = torch.tensor([ms, e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(tmax[0]), float(tmax[1])
    Qf = s.download()
    finite = all(np.isfinite(q).all() for q in Qf)

    if rank == 0:
        peak, peak_src = load_peaks()
        value = dof * world * 5 * args.steps / (ms * 1e-3)
        dom_name = max(ktimes, key=lambda k: ktimes[k][0])
        dom_ms, dom_n = ktimes[dom_name]
        B_alg = alg_bytes_per_element_stage(N)
        share = {"hex_stage": 1.0, "generic_rhs": (4 * Ns + 6 * (N + 1) ** 2) / (5 * Ns + 12 * (N + 1) ** 2)}.get(dom_name, 1.0)
        # one "launch" of the roofline = the launches that cover all K local elements once (1 on a single GPU;
        # boundary layers + interior when the stage is split for halo overlap)
        per_launch_s = dom_ms * 1e-3 / (args.steps * 5)
        achieved = B_alg * share * K / per_launch_s / 1e9
        from esdg_cns_b200 import _lib
        import ctypes
        pk = ctypes.c_double(0.0)
        _lib.load().esdg_measure_fp64_peak(local_rank, ctypes.byref(pk))
        F_alg = alg_flops_per_element_stage(N)
        fp64_tf = F_alg * K / per_launch_s / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"tgv_hex_N{N}_{Kx}x{Ky}x{Kzl}_per_gpu", "N": N, "elements_per_gpu": K,
                       "global_mesh": [Kx, Ky, Kz], "dof_per_gpu": dof, "stages_per_step": 5, "dt": dt,
                       "l2_policy": f"state {dof * 8 * 2 / 1e9:.2f} GB/GPU streamed every stage (>> 126 MB L2), no flush needed",
                       "parallelism": f"z-slab x{world}", "halo_exchange": ("none" if world == 1 else "host-driven (torch.distributed)" if args.py_halo else "in-library NCCL (esdg_comm_init)"), "finite": finite, "cpus_bound_per_rank": numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(dom_name, K)[0], "traffic_source": measured_traffic(dom_name, K)[1],
                         "kernel": dom_name, "kernel_ms_per_launch": per_launch_s * 1e3,
                         "kernel_launches_timed": dom_n, "kernel_share_of_step": dom_ms / ms,
                         "alg_bytes_per_element_stage": B_alg * share, "peak_source": peak_src,
                         "ns_per_element_stage": 1e9 * ms * 1e-3 / (args.steps * 5 * K)},
            "e2e": {"value": dof * world * 5 * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": dof * 8, "d2h_bytes_per_step": dof * 8,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps},
            "fp64": {"achieved_tflops": fp64_tf, "peak_tflops_measured_dfma": pk.value, "peak_tflops_nominal": 37.2,
                     "peak_source": "esdg_measure_fp64_peak: own DFMA micro-benchmark run on this GPU in this process (csrc/peak.cu); MEASURED_PEAKS.json has no FP64 entry",
                     "frac_of_measured": fp64_tf / pk.value if pk.value else None,
                     "alg_flops_per_element_stage": F_alg,
                     "note": "the stage kernel is FP64-pipe/latency bound, not HBM bound (profiles/)"},
            "gpu_launches": launches, "clocks": clocks,
            "kernels_ms": {k: {"total_ms": v[0], "launches": v[1]} for k, v in ktimes.items() if v[1]},
        }
        if world == 1 and not args.small and not args.no_other:
            s.close()
            del s
            torch.cuda.empty_cache()
            line["other_configs"] = other_configs(local_rank, peak)
        if not args.no_cpu and world == 1:      # reported on rank 0 at N = 1 only
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="16x16x8 elements per GPU (debugging only; not a bench number)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-other", action="store_true", help="skip the other_configs block (BASELINE configs 1-4 at timing scale, curved hexes)")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: exchange halos without overlapping interior work")
    ap.add_argument("--py-halo", action="store_true", help="multi-GPU: host-driven halo exchange (torch.distributed) instead of the context's own NCCL communicator")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
