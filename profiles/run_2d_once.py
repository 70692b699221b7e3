#!/usr/bin/env python
"""One short run of the 2D paths for ncu captures: MODE=euler|cns (default both), K1D elements per side."""
import os, sys
from types import SimpleNamespace
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esdg_cns_b200 import api, setup_dg as SD
N, K1D, mode = 3, int(os.environ.get("K1D", "256")), os.environ.get("MODE", "both")
if mode in ("both", "euler"):
    rd, ops = SD.reference_quad(N)
    md = SD.uniform_quad_mesh_md(rd, 2 * K1D, 2 * K1D, domain=((0, 20), (-5, 5)))
    rho, u, v, p = SD.vortex(md.xq, md.yq)
    s = api.Solver(md, ops)
    s.upload(SD.primitive_to_conservative(rho, (u, v), p))
    s.lsrk45(1e-4, 2)
    s.sync()
if mode in ("both", "cns"):
    rd, ops = SD.reference_quad_modal(N)
    md = SD.uniform_quad_mesh_md(rd, K1D, K1D, periodic=(False, False))
    bc = SD.cavity_bc_sets(md)
    st = SimpleNamespace(Re=1000.0, mu=1e-3, lam=-2e-3 / 3, Pr=.71, bctype=1, inviscid_dissp=True, viscous_dissp=True,
                         wall=bc.wall, lid=bc.lid, vlid=bc.vlid, mapB=bc.mapB)
    one = np.ones_like(md.x)
    s = api.Solver(md, ops, cns=st)
    s.upload(SD.primitive_to_conservative(one, (0 * one, 0 * one), one / (.3 ** 2 * 1.4)))
    s.lsrk45(1e-5, 2)
    s.sync()
