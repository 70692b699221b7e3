#!/usr/bin/env python
"""2D paths at timing scale: Euler on collocated quads (config 1 operators, periodic vortex) and
CNS on modal quads (config 3: lid-driven cavity, BCTYPE=1, both dissipations on), LSRK45 stages/s."""
import os, sys, time
from types import SimpleNamespace
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esdg_cns_b200 import api, setup_dg as SD


def run(name, s, Q, dt, kernels, Np, K):
    s.upload(Q)
    s.lsrk45(dt, 2)
    s.sync()
    s.enable_timing(True)
    t0 = time.perf_counter()
    nsteps = 10
    s.lsrk45(dt, nsteps)
    s.sync()
    wall = time.perf_counter() - t0
    Q1 = s.download()
    dof = Np * K * 4
    print(f"{name}: K={K} dof={dof} {wall/nsteps/5*1e3:.3f} ms/stage {dof*5*nsteps/wall*1e-9:.2f} GDOF*stage/s finite={all(np.isfinite(q).all() for q in Q1)}")
    for k in kernels:
        ms, nl = s.kernel_time_ms(k)
        if nl:
            print(f"    {k:16s} {ms/nl:.3f} ms/launch x {nl}")
    s.close()


N = 3
K1D = int(os.environ.get("K1D", "512"))
# Euler, collocated quads, periodic vortex (config 1 at scale)
rd, ops = SD.reference_quad(N)
md = SD.uniform_quad_mesh_md(rd, 2 * K1D, 2 * K1D, domain=((0, 20), (-5, 5)))
rho, u, v, p = SD.vortex(md.xq, md.yq)
Q = SD.primitive_to_conservative(rho, (u, v), p)
run("euler quad N=3 (generic kernels)", api.Solver(md, ops), Q, 1e-4, ["generic_rhs", "generic_traces"], 16, md.K)
# CNS, modal quads, cavity (config 3 at scale)
rd, ops = SD.reference_quad_modal(N)
md = SD.uniform_quad_mesh_md(rd, K1D, K1D, periodic=(False, False))
bc = SD.cavity_bc_sets(md)
Re = 1000.0
st = SimpleNamespace(Re=Re, mu=1 / Re, lam=-2 / 3 / Re, Pr=.71, bctype=1, inviscid_dissp=True, viscous_dissp=True,
                     wall=bc.wall, lid=bc.lid, vlid=bc.vlid, mapB=bc.mapB)
one = np.ones_like(md.x)
Q = SD.primitive_to_conservative(one, (0 * one, 0 * one), one / (.3 ** 2 * 1.4))
dt = 1e-5   # timing only (the reference adapts dt with DOPRI45; the fixed-CFL dt of :39-42 is viscous-unstable at K1D=512)
run("cns quad N=3 cavity", api.Solver(md, ops, cns=st), Q, dt, ["cns_traces", "cns_sigma", "cns_rhs"], 16, md.K)
