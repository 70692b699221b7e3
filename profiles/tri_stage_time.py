"""ms per LSRK stage of BASELINE config 2 at timing scale (modal triangles, table-driven kernels) on B200."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from esdg_cns_b200 import api, setup_dg as SD
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
Kx = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rd, ops = SD.reference_tri(N)
md = SD.uniform_tri_mesh_md(rd, Kx, Kx, domain=((0, 20), (-5, 5)))
rho, u, v, p = SD.vortex(md.x, md.y)
Q = SD.primitive_to_conservative(rho, (u, v), p)
s = api.Solver(md, ops)
print(s.kernel_path(), "K =", md.K)
s.upload(Q); s.lsrk45(1e-4, 2); s.sync()
s.enable_timing(True)
s.lsrk45(1e-4, 4); s.sync()
tot = 0
for k in ("generic_traces", "generic_rhs"):
    ms, nl = s.kernel_time_ms(k)
    print(f"  {k:16s} {ms / max(nl, 1):.4f} ms x {nl}")
    tot += ms
per = tot / 20
dof = 4 * rd.r.size * md.K
print(f"tri N={N} 2x{Kx}^2: {per:.4f} ms/stage {dof / per / 1e6:.2f} GDOF*stage/s finite={all(np.isfinite(q).all() for q in s.download())}")
