import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from esdg_cns_b200 import api, setup_dg as SD
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Kx = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rd, ops = SD.reference_quad(N)
md = SD.uniform_quad_mesh_md(rd, Kx, Kx, domain=((0, 20), (-5, 5)))
rho, u, v, p = SD.vortex(md.xq, md.yq)
Q = SD.primitive_to_conservative(rho, (u, v), p)
s = api.Solver(md, ops)
print(s.kernel_path())
s.upload(Q); s.lsrk45(1e-4, 2); s.sync()
import torch
best = 1e9
for rep in range(4):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    s.sync(); t0 = time.perf_counter()
    s.lsrk45(1e-4, 4); s.sync()
    best = min(best, (time.perf_counter() - t0) * 1e3 / 20)
dof = 4 * (N + 1) ** 2 * md.K
print(f"quad N={N} {Kx}^2: {best:.4f} ms/stage  {dof / best / 1e6:.2f} GDOF*stage/s")
