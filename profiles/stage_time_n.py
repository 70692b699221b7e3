#!/usr/bin/env python
"""Stage-kernel time for one polynomial degree: python profiles/stage_time_n.py N Kx Ky Kz  (ESDG_B200_LIB selects the build)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esdg_cns_b200 import api, setup_dg as SD
N = int(sys.argv[1]); mesh = tuple(int(a) for a in sys.argv[2:5])
prd, pops = SD.reference_hex(N)
pmd = SD.uniform_hex_mesh_md(prd, mesh, domain=((-np.pi, np.pi),) * 3)
x, y, z = SD.hex_node_coords(prd, pmd)
rho, u, v, w, p = SD.taylor_green(x, y, z)
Q = SD.primitive_to_conservative(rho, (u, v, w), p)
s = api.Solver(pmd, pops)
s.upload(Q)
s.lsrk45(1e-4, 2)
s.sync()
s.enable_timing(True)
s.lsrk45(1e-4, 4)
s.sync()
ms, nl = s.kernel_time_ms("hex_stage")
Q1 = s.download()
K = pmd.K
print(f"{os.environ.get('ESDG_B200_LIB', 'default')} N={N} K={K}: {ms/nl:.3f} ms/stage {ms/nl*1e6/K:.2f} ns/element-stage "
      f"{K*(N+1)**3*5/(ms/nl)*1e-6:.2f} GDOF*stage/s chk={sum(float(q.sum()) for q in Q1):.12e}")
