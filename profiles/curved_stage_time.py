"""ms per LSRK stage of the curved-hex line kernel against the table-driven kernels on the same curved mesh (B200)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from esdg_cns_b200 import api, setup_dg as SD
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
Kx = int(sys.argv[2]) if len(sys.argv) > 2 else 32
prd, pops = SD.reference_hex(N)
dom = ((-np.pi, np.pi),) * 3
md = SD.curved_hex_mesh_md(prd, (Kx, Kx, Kx), domain=dom, amp=.1)
rho, u, v, w, p = SD.taylor_green(md.xq, md.yq, md.zq)
Q = SD.primitive_to_conservative(rho, (u, v, w), p)
s = api.Solver(md, pops)
print(s.kernel_path(), "K =", md.K)
s.upload(Q); s.lsrk45(1e-4, 1); s.sync()
best = 1e9
for rep in range(3):
    s.sync(); t0 = time.perf_counter()
    s.lsrk45(1e-4, 2); s.sync()
    best = min(best, (time.perf_counter() - t0) * 1e3 / 10)
dof = 5 * (N + 1) ** 3 * md.K
print(f"curved hex N={N} {Kx}^3: {best:.4f} ms/stage  {dof / best / 1e6:.2f} GDOF*stage/s  {best * 1e6 / md.K:.1f} ns per element-stage; finite={all(np.isfinite(q).all() for q in s.download())}")
