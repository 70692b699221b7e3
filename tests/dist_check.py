"""Multi-GPU z-slab check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py [--overlap 0|1] [--N 2] [--K 6 6 8]

Every rank advances its slab with GpuSlabStepper (NCCL halo exchange); rank 0 also advances the
whole mesh on its own GPU with the single-context path and compares the gathered slabs with it.
The partitioned run must reproduce the single-GPU state to round-off (same kernels, same
per-element arithmetic; only the trace transport differs).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from esdg_cns_b200 import api, halo as H, setup_dg as SD

    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=2)
    ap.add_argument("--K", type=int, nargs=3, default=[6, 5, 8])
    ap.add_argument("--overlap", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--generic", type=int, default=0)
    ap.add_argument("--cns", type=int, default=0, help="1: compressible Navier-Stokes cavity on y-slabs of a quad mesh")
    ap.add_argument("--quad", type=int, default=0, help="1: Euler on y-slabs of a periodic quad mesh (tensor-product quad kernels), K = Kx Ky 0")
    ap.add_argument("--ccomm", type=int, default=0, help="1: the context's own NCCL communicator (esdg_comm_init) instead of the host-driven exchange")
    a = ap.parse_args()
    if a.generic:
        os.environ["ESDG_FORCE_GENERIC"] = "1"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    if a.cns:
        ok = cns_check(a, rank, world, lr, dev, torch, dist, api, H, SD)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, 0)
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if int(flag[0]) == 1 else 1)
    N, Kxyz = a.N, tuple(a.K)
    if a.quad:
        dom = ((0.0, 20.0), (-5.0, 5.0))
        prd, pops = SD.reference_quad(N)
        md = SD.uniform_quad_mesh_md(prd, Kxyz[0], Kxyz[1], domain=dom, yslab=(rank, world))
        Q0 = SD.primitive_to_conservative(*(lambda r, u, v, p: (r, (u, v), p))(*SD.vortex(md.xq, md.yq)))
    else:
        dom = ((-np.pi, np.pi),) * 3
        prd, pops = SD.reference_hex(N)
        md = SD.uniform_hex_mesh_md(prd, Kxyz, domain=dom, zslab=(rank, world))
        x, y, z = SD.hex_node_coords(prd, md)
        Q0 = SD.primitive_to_conservative(*(lambda r, u, v, w, p: (r, (u, v, w), p))(*SD.taylor_green(x, y, z)))
    dt = 2e-3
    s = api.Solver(md, pops, device=lr, n_halo=md.n_halo, stream=torch.cuda.current_stream().cuda_stream)
    if a.ccomm:
        if not a.overlap:
            os.environ["ESDG_COMM_NO_OVERLAP"] = "1"
        H.attach_comm(s, md, dist, torch, dev)
        s.upload(Q0)
        rt = 0.0
        for k in range(a.steps):
            rt = s.lsrk45(dt, 1, want_rhstest=(k == a.steps - 1))
        rt_local = torch.tensor([rt], dtype=torch.float64, device=dev)     # already all-reduced inside the library
    else:
        st = H.GpuSlabStepper(s, md, dist, torch, dev, overlap=bool(a.overlap))
        s.upload(Q0)
        for k in range(a.steps):
            st.step(dt, want_rhstest=(k == a.steps - 1))
        rt_local = torch.tensor([s.read_rhstest()], dtype=torch.float64, device=dev)
        dist.all_reduce(rt_local)                      # entropy residual: sum of the slab partials
    Ql = np.stack(s.download())                    # (5, Nq, K_local)
    t = torch.from_numpy(np.ascontiguousarray(Ql)).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, parts, dst=0)
    ok = True
    if rank == 0:
        full = np.concatenate([p.cpu().numpy() for p in parts], axis=2)
        if a.quad:
            mdf = SD.uniform_quad_mesh_md(prd, Kxyz[0], Kxyz[1], domain=dom)
            Qf = SD.primitive_to_conservative(*(lambda r, u, v, p: (r, (u, v), p))(*SD.vortex(mdf.xq, mdf.yq)))
        else:
            mdf = SD.uniform_hex_mesh_md(prd, Kxyz, domain=dom)
            xf, yf, zf = SD.hex_node_coords(prd, mdf)
            Qf = SD.primitive_to_conservative(*(lambda r, u, v, w, p: (r, (u, v, w), p))(*SD.taylor_green(xf, yf, zf)))
        s1 = api.Solver(mdf, pops, device=lr)
        s1.upload(Qf)
        rt1 = s1.lsrk45(dt, a.steps, want_rhstest=True)
        ref = np.stack(s1.download())
        err = np.abs(full - ref).max() / np.abs(ref).max()
        rterr = abs(float(rt_local[0]) - rt1)
        print(f"dist_check world={world} N={N} K={Kxyz} quad={a.quad} path={s.kernel_path()} overlap={a.overlap} generic={a.generic}: "
              f"state max rel diff {err:.3e}; rhstest {float(rt_local[0]):.12e} vs single-GPU {rt1:.12e}")
        ok = err < 1e-13 and rterr < 1e-10 * max(1.0, abs(rt1))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


def cns_check(a, rank, world, lr, dev, torch, dist, api, H, SD):
    """CNS lid-driven cavity (config 3 shape) on y-slabs: two halo exchanges per stage (entropy-variable
    traces, sigma traces) over NCCL; must equal the single-GPU run bit for bit."""
    from types import SimpleNamespace
    N, Kx, Ky = a.N, a.K[0], a.K[1] * world
    prd, pops = SD.reference_quad_modal(N)

    def settings(md):
        bc = SD.cavity_bc_sets(md)
        return SimpleNamespace(Re=100.0, mu=1e-2, lam=-2e-2 / 3, Pr=.71, bctype=1, inviscid_dissp=True,
                               viscous_dissp=True, wall=bc.wall, lid=bc.lid, vlid=bc.vlid, mapB=bc.mapB)

    def state(md):
        x, y = md.x, md.y
        rho = 1 + .1 * np.sin(np.pi * x) * np.cos(np.pi * y)
        u, v = .2 * np.cos(np.pi * x), -.1 * np.sin(np.pi * y) * np.cos(np.pi * x)
        p = 1 / (.3 ** 2 * 1.4) * (1 + .05 * np.cos(np.pi * (x + y)))
        return SD.primitive_to_conservative(rho, (u, v), p)
    md = SD.uniform_quad_mesh_md(prd, Kx, Ky, periodic=(False, False), yslab=(rank, world))
    dt = 2e-4
    s = api.Solver(md, pops, device=lr, n_halo=md.n_halo, cns=settings(md), stream=torch.cuda.current_stream().cuda_stream)
    if a.ccomm:
        H.attach_comm(s, md, dist, torch, dev)
        s.upload(state(md))
        for k in range(a.steps):
            s.lsrk45(dt, 1, want_rhstest=(k == a.steps - 1))
        rt_local = torch.tensor([s.read_rhstest(), s.read_rhstest_visc()], dtype=torch.float64, device=dev)
    else:
        st = H.GpuSlabStepperCNS(s, md, dist, torch, dev)
        s.upload(state(md))
        for k in range(a.steps):
            st.step(dt, want_rhstest=(k == a.steps - 1))
        rt_local = torch.tensor([s.read_rhstest(), s.read_rhstest_visc()], dtype=torch.float64, device=dev)
        dist.all_reduce(rt_local)
    t = torch.from_numpy(np.ascontiguousarray(np.stack(s.download()))).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, parts, dst=0)
    ok = True
    if rank == 0:
        full = np.concatenate([p.cpu().numpy() for p in parts], axis=2)
        mdf = SD.uniform_quad_mesh_md(prd, Kx, Ky, periodic=(False, False))
        s1 = api.Solver(mdf, pops, device=lr, cns=settings(mdf))
        s1.upload(state(mdf))
        rt1 = s1.lsrk45(dt, a.steps, want_rhstest=True)
        rtv1 = s1.read_rhstest_visc()
        ref = np.stack(s1.download())
        diff = np.abs(full - ref).max()
        print(f"dist_check CNS world={world} N={N} K={Kx}x{Ky}: state max diff {diff:.3e}; rhstest {float(rt_local[0]):.12e} vs "
              f"{rt1:.12e}; rhstest_visc {float(rt_local[1]):.12e} vs {rtv1:.12e}")
        ok = diff == 0.0 and abs(float(rt_local[0]) - rt1) < 1e-10 * max(1.0, abs(rt1)) and \
            abs(float(rt_local[1]) - rtv1) < 1e-10 * max(1.0, abs(rtv1))
    return ok


if __name__ == "__main__":
    main()
