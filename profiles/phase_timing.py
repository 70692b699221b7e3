#!/usr/bin/env python
"""Per-phase cycle breakdown of k_hex_stage_line (debug build `make -C esdg_cns_b200/csrc timing`).
usage (GPU box): ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_timing.so python profiles/phase_timing.py"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esdg_cns_b200 import api, setup_dg as SD, _lib

N, mesh = 4, (128, 128, 16)
prd, pops = SD.reference_hex(N)
pmd = SD.uniform_hex_mesh_md(prd, mesh, domain=((-np.pi, np.pi),) * 3)
x, y, z = SD.hex_node_coords(prd, pmd)
Q = SD.primitive_to_conservative(*(lambda r, u, v, w, p: (r, (u, v, w), p))(*SD.taylor_green(x, y, z)))
s = api.Solver(pmd, pops)
s.upload(Q)
s.lsrk45(1e-4, 2)
s.sync()
lib = _lib.load()
buf = (ctypes.c_ulonglong * 16)()
lib.esdg_debug_phase_cycles(buf)
s.lsrk45(1e-4, 2)
s.sync()
lib.esdg_debug_phase_cycles(buf)
names = ["0 issue loads", "1 wait loads+sync", "2 A vol prims", "3 B surface", "4 sync", "5 C lines", "6 sync",
         "7 E update", "8 sync", "9 F traces"]
names += ["10 (bulk issue)", "11 (gather issue)", "12 (mapP ld issue)", "13 (mbar init)", "14 (tables+sync)"]
tot = sum(buf[:15])
ncta = 10 * (pmd.K // 2)
for k, nm in enumerate(names):  # phase 0 = prefetch issue only when 10/11 are present
    print(f"{nm:20s} {buf[k]/ncta:9.0f} cycles/CTA {100*buf[k]/tot:5.1f}%")
print(f"total {tot/ncta:.0f} cycles per CTA (thread 0 timeline)")
