#!/usr/bin/env python
"""Per-activity cycle breakdown of the warp-specialised hex stage kernel (debug build with -DESDG_PHASE_TIMING:
`profiles/build_variant.sh timing -DESDG_PHASE_TIMING`).
usage (GPU box): ESDG_B200_LIB=scratch/v/libesdg_b200_timing.so python profiles/phase_timing.py"""
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
nst = 10
s.enable_timing(True)
s.lsrk45(1e-4, nst // 5)
s.sync()
lib.esdg_debug_phase_cycles(buf)
ms, nl = s.kernel_time_ms("hex_stage")
names = {0: "L loop overhead", 1: "L wait full_wv (A done)", 2: "L wait full_tp (traces landed)", 3: "L line_compute", 4: "L wait empty_a3 (E done)",
         5: "L line_store", 6: "P loop / lambda overhead", 7: "P wait full_a3 (lines done)", 8: "P E update", 9: "P bar", 10: "P F traces",
         11: "P rhstest + bar", 12: "P wait empty_in", 13: "P issue loads + next mapP", 14: "P geo + A prims"}
nbatch = nst * ((pmd.K + 2) // 3)
L = sum(buf[k] for k in range(0, 6))
P = sum(buf[k] for k in range(6, 15))
for k in range(15):
    tot = L if k < 6 else P
    print(f"{k:2d} {names[k]:32s} {buf[k]/nbatch:9.0f} cycles/batch {100*buf[k]/tot:5.1f}% of the team's time")
print(f"line team {L/nbatch:.0f}, point team {P/nbatch:.0f} cycles per batch of 3 elements (thread 0 of each team); {ms/nl:.3f} ms per stage (timing build)")
