import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from types import SimpleNamespace
from esdg_cns_b200 import api, setup_dg as SD
N = 3; Kx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rd, ops = SD.reference_quad_modal(N)
md = SD.uniform_quad_mesh_md(rd, Kx, Kx, periodic=(False, False))
bc = SD.cavity_bc_sets(md)
Re = 1000.0
st = SimpleNamespace(Re=Re, mu=1 / Re, lam=-2 / 3 / Re, Pr=.71, bctype=1, inviscid_dissp=True, viscous_dissp=True,
                     wall=bc.wall, lid=bc.lid, vlid=bc.vlid, mapB=bc.mapB)
one = np.ones_like(md.x)
Q = SD.primitive_to_conservative(one, (0 * one, 0 * one), one / (.3 ** 2 * 1.4))
s = api.Solver(md, ops, cns=st)
print(s.kernel_path())
s.upload(Q); s.lsrk45(1e-5, 2); s.sync()
s.enable_timing(True)
s.lsrk45(1e-5, 4); s.sync()
tot = 0
for k in ("cns_traces", "cns_sigma", "cns_rhs"):
    ms, nl = s.kernel_time_ms(k)
    print(f"  {k:12s} {ms / max(nl, 1):.4f} ms x {nl}")
    tot += ms
per = tot / 20
dof = 4 * 16 * md.K
print(f"cns quad N=3 {Kx}^2: {per:.4f} ms/stage {dof / per / 1e6:.2f} GDOF*stage/s finite={all(np.isfinite(q).all() for q in s.download())}")
