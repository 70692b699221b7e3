#!/usr/bin/env python
"""Convergence sweep of the lid-driven cavity, the harness of
examples/CompressibleNS/dg2D_CNS_convergence_test.jl:841-1088 on the GPU path.

For every (K1D, N, inviscid_dissp, viscous_dissp, Re) the cavity (rho = 1, u = v = 0, p = 1/(Ma^2 gamma), Ma = .3,
smooth lid velocity (1 + cos(pi x))/2, :1066) is advanced to T with the script's DOPRI45 loop and the weak imposition of
the no-slip condition is measured in the boundary L2 norm (:1068-1078):

    err^2 = sum_b Jf wf u_2^2 + sum_wall Jf wf u_1^2 + sum_lid Jf wf (u_1 - vlid)^2

The reference sweeps triangles; the package ships structured quad setup, so the sweep runs the same rhs on the quad
element (BASELINE config 3).  Usage:  python examples/cns_convergence_sweep.py [--K1D 8 16] [--N 1 2 3] [--Re 100] [--T 1.0]
"""
import argparse
import os
import sys
from types import SimpleNamespace

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def boundary_error(Q, rd, ops, bc, vlid, K1D):
    """:1068-1078 -- boundary L2 norm of the velocity mismatch."""
    u1 = ops.Vf @ (Q[1] / Q[0])
    u2 = ops.Vf @ (Q[2] / Q[0])
    Jf = 2.0 / K1D
    WF = np.repeat(np.asarray(rd.wf).reshape(-1, 1), Q[0].shape[1], axis=1).reshape(-1, order="F")
    u1, u2 = u1.reshape(-1, order="F"), u2.reshape(-1, order="F")
    boundary = np.concatenate([bc.lid, bc.wall])
    err = np.sum(Jf * WF[boundary] * u2[boundary] ** 2) + np.sum(Jf * WF[bc.wall] * u1[bc.wall] ** 2) + \
        np.sum(Jf * WF[bc.lid] * (u1[bc.lid] - vlid) ** 2)
    return float(np.sqrt(err))


def run_case(N, K1D, Re, inviscid_dissp=True, viscous_dissp=True, T=1.0, CFL=.01, bctype=1, err_tol=1e-5, device=0):
    from esdg_cns_b200 import api, setup_dg as SD
    rd, ops = SD.reference_quad_modal(N)
    md = SD.uniform_quad_mesh_md(rd, K1D, K1D, periodic=(False, False))
    bc = SD.cavity_bc_sets(md)
    xlid = md.xf.reshape(-1, order="F")[bc.lid]
    vlid = (1 + np.cos(np.pi * xlid)) / 2                                   # :1066
    mu = 1.0 / Re
    st = SimpleNamespace(Re=Re, mu=mu, lam=-2 / 3 * mu, Pr=.71, bctype=bctype, inviscid_dissp=inviscid_dissp,
                         viscous_dissp=viscous_dissp, wall=bc.wall, lid=bc.lid, vlid=vlid, mapB=bc.mapB)
    one = np.ones_like(md.x)
    Q = SD.primitive_to_conservative(one, (0 * one, 0 * one), one / (.3 ** 2 * 1.4))
    CN = (N + 1) * (N + 2) / 2                                              # :874-879
    dt = CFL * (2 / K1D) / CN
    dt = T / np.ceil(T / dt)
    s = api.Solver(md, ops, device=device, cns=st)
    s.upload(Q)
    hist = s.dopri45(T, dt, err_tol=err_tol)
    Q1 = s.download()
    s.close()
    return boundary_error(Q1, rd, ops, bc, vlid, K1D), hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K1D", type=int, nargs="+", default=[8, 16])
    ap.add_argument("--N", type=int, nargs="+", default=[1, 2, 3, 4])
    ap.add_argument("--Re", type=float, nargs="+", default=[100.0])
    ap.add_argument("--T", type=float, default=1.0)
    a = ap.parse_args()
    err = np.zeros((len(a.K1D), len(a.N), len(a.Re)))
    for ki, K1D in enumerate(a.K1D):
        for ni, N in enumerate(a.N):
            for ri, Re in enumerate(a.Re):
                e, hist = run_case(N, K1D, Re, T=a.T)
                err[ki, ni, ri] = e
                print(f"K1D={K1D} N={N} Re={Re:g}: boundary error {e:.4e}  (t = {hist['t'][-1]:.3f}, {len(hist['accepted'])} attempts, "
                      f"{sum(hist['accepted'])} accepted, rhstest {hist['rhstest'][-1]:.2e})")
    np.savetxt("err_arr.txt", err.reshape(len(a.K1D), -1))                 # :1087-1088
    return err


if __name__ == "__main__":
    main()
