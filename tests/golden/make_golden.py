"""Generate the committed golden fixtures from the numpy oracle (run in the build container).

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference ships no golden vectors for rhs and Julia is not available, so
these fixtures pin the ORACLE (a restatement) against regressions; they are not Julia output.
Each .npz holds the inputs (state) and the oracle's rhs / post-step state for a small case.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cns, physics as P, solvers as SV  # noqa: E402


def quad_case():
    rd, md, ops, Q = SV.setup_quad_vortex(3, 4, 4)
    rhs, rt = SV.rhs_collocated(Q, md, ops, rd, .5, True)
    dt = 2 * (2 / 16) / 10
    Q1, _, rt1 = SV.lsrk45_step(Q, [np.zeros_like(q) for q in Q], dt, lambda q, t: SV.rhs_collocated(q, md, ops, rd, .5, t))
    np.savez_compressed(os.path.join(HERE, "quad_N3_4x4_vortex.npz"), Q=np.stack(Q), rhs=np.stack(rhs), rhstest=rt,
                        dt=dt, Q1=np.stack(Q1), rhstest1=rt1)


def hex_case():
    rd, md, ops = SV.setup_hex(2, 3, scale=(np.pi,) * 3)
    rho, u, v, w, p = P.taylor_green(md.xq, md.yq, md.zq)
    Q = list(P.primitive_to_conservative(rho, (u, v, w), p))
    rhs, rt = SV.rhs_collocated(Q, md, ops, rd, .25, True)
    dt = 1e-3
    Q1, _, rt1 = SV.lsrk45_step(Q, [np.zeros_like(q) for q in Q], dt, lambda q, t: SV.rhs_collocated(q, md, ops, rd, .25, t))
    np.savez_compressed(os.path.join(HERE, "hex_N2_3x3x3_tgv.npz"), Q=np.stack(Q), rhs=np.stack(rhs), rhstest=rt,
                        dt=dt, Q1=np.stack(Q1), rhstest1=rt1)


def tri_case():
    rd, md, ops, Q = SV.setup_tri_vortex(3, 3, 2)
    rhs, rt = SV.rhs_tri(Q, md, ops, rd, .25, True)
    np.savez_compressed(os.path.join(HERE, "tri_N3_3x2_vortex.npz"), Q=np.stack(Q), rhs=np.stack(rhs), rhstest=rt)


def flux_kat():
    """Known-answer vector for the two-point flux at the states of runtests.jl:133-160."""
    UL = P.primitive_to_conservative(1.0, (.1, .2, .3), 2.0)
    UR = P.primitive_to_conservative(1.1, (.2, .3, .4), 2.1)
    QL = (UL[0], UL[1] / UL[0], UL[2] / UL[0], UL[3] / UL[0], P.beta_of_u(UL))
    QR = (UR[0], UR[1] / UR[0], UR[2] / UR[0], UR[3] / UR[0], P.beta_of_u(UR))
    F = P.euler_fluxes_3d(QL, QR)
    np.savez_compressed(os.path.join(HERE, "flux3d_kat.npz"), QL=np.array(QL, dtype=float), QR=np.array(QR, dtype=float),
                        F=np.array(F, dtype=float))


def cns_state(md):
    """Smooth non-trivial state on the cavity domain (walls at +-1) used by the CNS parity tests."""
    x, y = md.x, md.y
    rho = 1 + .1 * np.sin(np.pi * x) * np.cos(np.pi * y)
    u = .2 * np.sin(np.pi * x) * np.sin(np.pi * y)
    v = .1 * np.cos(.5 * np.pi * x) * np.cos(.5 * np.pi * y) * (1 - y ** 2)
    p = 1 / (.09 * 1.4) * (1 + .05 * np.cos(np.pi * x))
    return list(P.primitive_to_conservative(rho, (u, v), p))


def cns_case():
    """BASELINE config 3: CNS on quads N=3, lid-driven cavity BCs (adiabatic no-slip), both dissipations on."""
    rd, md, ops, bc, _ = cns.setup_cavity("quad", 3, 4)
    Q = cns_state(md)
    par = cns.cavity_params()
    rhs, rt, rtv = cns.rhsRK(Q, md, ops, bc, par)
    np.savez_compressed(os.path.join(HERE, "cns_quad_N3_4x4_cavity.npz"), Q=np.stack(Q), rhs=np.stack(rhs), rhstest=rt,
                        rhstest_visc=rtv)


if __name__ == "__main__":
    cns_case()
    quad_case()
    hex_case()
    tri_case()
    flux_kat()
    print(sorted(os.listdir(HERE)))
