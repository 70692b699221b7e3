#!/usr/bin/env python
"""Stage-kernel time per element for the hex path, per polynomial degree.
usage (GPU box): [ESDG_HEX_KERNEL=node] python profiles/compare_kernels.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esdg_cns_b200 import api, setup_dg as SD

for N, mesh in ((1, (128, 128, 64)), (2, (128, 128, 32)), (3, (64, 64, 64)), (4, (128, 128, 16))):
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
    s.lsrk45(1e-4, 3)
    s.sync()
    ms, nl = s.kernel_time_ms("hex_stage")
    K = pmd.K
    dof = K * (N + 1) ** 3 * 5
    print(f"N={N} K={K}: {ms/nl:.3f} ms/stage, {ms/nl*1e6/K:.2f} ns/element-stage, {dof/(ms/nl)*1e-6:.2f} GDOF*stage/s "
          f"[{os.environ.get('ESDG_HEX_KERNEL', 'line')}]")
    s.close()
