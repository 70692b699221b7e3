#!/usr/bin/env python
"""Stage-kernel time of the bench workload (hex N=4, 128x128x16) for whichever library ESDG_B200_LIB points to.
LF=0 in the environment switches the Lax-Friedrichs penalty off (lf_scale = 0)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esdg_cns_b200 import api, setup_dg as SD
N, mesh = 4, (128, 128, 16)
prd, pops = SD.reference_hex(N, lf_scale=float(os.environ.get("LF", ".25")))
pmd = SD.uniform_hex_mesh_md(prd, mesh, domain=((-np.pi, np.pi),) * 3)
x, y, z = SD.hex_node_coords(prd, pmd)
rho, u, v, w, p = SD.taylor_green(x, y, z)
rho = rho + .05 * np.sin(x) * np.cos(y)      # keep states apart so that experiment builds without the Taylor branch stay finite
Q = SD.primitive_to_conservative(rho, (u, v, w), p)
s = api.Solver(pmd, pops)
s.upload(Q)
s.lsrk45(1e-4, 3)
s.sync()
s.enable_timing(True)
best, prev = 1e9, (0.0, 0)
for rep in range(4):
    s.lsrk45(1e-4, 4)
    s.sync()
    ms, nl = s.kernel_time_ms("hex_stage")
    best = min(best, (ms - prev[0]) / (nl - prev[1]))
    prev = (ms, nl)
ms, nl = best, 1
Q1 = s.download()
print(f"{os.environ.get('ESDG_B200_LIB', 'default')} LF={os.environ.get('LF', '.25')}: {ms/nl:.3f} ms/stage (best of 4 x 20 stages)  finite={all(np.isfinite(q).all() for q in Q1)}  chk={sum(float(q.sum()) for q in Q1):.12e}")
