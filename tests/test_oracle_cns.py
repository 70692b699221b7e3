"""Self-checks of the CNS oracle (oracle/cns.py, restating dg2D_CNS_cavity_optimized.jl).

PARITY UNPINNED (no Julia, no reference fixtures): what pins the restatement is physics --
the viscous matrices must reproduce the Navier-Stokes stress tensor / heat flux through the chain
rule, the BR1 gradient must differentiate polynomials exactly, the wall BCs must be entropy stable
(rhstest_visc <= 0, rhstest <= 0) -- plus the committed golden fixture.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import cns, physics as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_viscous_matrices_reproduce_navier_stokes_stress():
    """sigma = K(v) grad v equals (tau_xx, tau_xy, u.tau + kappa grad T) with
    lambda' = -lambda_in (the `let lambda=-lambda` of :614), kappa = gamma*mu*cv/Pr."""
    rng = np.random.default_rng(3)
    n = 50
    rho = 1 + .3 * rng.random(n)
    u, v = rng.standard_normal(n), rng.standard_normal(n)
    p = 1 + rng.random(n)
    # random gradients of primitive variables (u, v, e)
    ux, uy, vx, vy, ex, ey = (rng.standard_normal(n) for _ in range(6))
    e = p / (.4 * rho)
    # entropy variables v2 = rho*u/rhoe = u/e, v3 = v/e, v4 = -1/e  (euler_variables.jl:79-86)
    v2, v3, v4 = u / e, v / e, -1 / e
    d = lambda q_x, q, e_x: (q_x * e - q * e_x) / e ** 2          # noqa: E731  d(q/e)
    v2x, v2y, v3x, v3y = d(ux, u, ex), d(uy, u, ey), d(vx, v, ex), d(vy, v, ey)
    v4x, v4y = ex / e ** 2, ey / e ** 2
    lam_in, mu, Pr = -2 / 3 * 1e-2, 1e-2, .71
    Kxx, Kxy, Kyy = cns.viscous_matrices((None, v2, v3, v4), lam_in, mu, Pr)
    tx = [None, v2x, v3x, v4x]
    ty = [None, v2y, v3y, v4y]
    sx = [0, 0, 0, 0]
    sy = [0, 0, 0, 0]
    for col in range(1, 4):
        for row in range(1, 4):
            sx[row] = sx[row] + Kxx[row][col] * tx[col] + Kxy[row][col] * ty[col]
            sy[row] = sy[row] + Kxy[col][row] * tx[col] + Kyy[row][col] * ty[col]
    lam = -lam_in
    txx = (lam + 2 * mu) * ux + lam * vy
    tyy = (lam + 2 * mu) * vy + lam * ux
    txy = mu * (uy + vx)
    kap = 1.4 * mu / Pr
    for got, want in ((sx[1], txx), (sx[2], txy), (sy[1], txy), (sy[2], tyy),
                      (sx[3], u * txx + v * txy + kap * ex), (sy[3], u * txy + v * tyy + kap * ey)):
        assert np.abs(got - want).max() < 1e-12 * max(1, np.abs(want).max())


@pytest.mark.parametrize("elem", ["tri", "quad"])
def test_br1_gradient_is_exact_for_polynomials(elem):
    N = 3
    rd, md, ops, bc, _ = cns.setup_cavity(elem, N, 3)
    f = md.x ** 2 * md.y - .5 * md.y ** 3 + md.x                 # degree 3: in the space
    fx = 2 * md.x * md.y + 1
    fy = md.x ** 2 - 1.5 * md.y ** 2
    ff = ops.Vf @ f
    fP = ff.flatten(order="F")[md.mapP.flatten(order="F")].reshape(ff.shape, order="F")
    gx, gy = cns.dg_grad([f], [ff], [fP], md, ops)
    assert np.abs(gx[0] - fx).max() < 1e-11
    assert np.abs(gy[0] - fy).max() < 1e-11


@pytest.mark.parametrize("elem", ["tri", "quad"])
@pytest.mark.parametrize("bctype", [1, 2, 3])
def test_wall_bcs_are_entropy_stable(elem, bctype):
    rd, md, ops, bc, Q = cns.setup_cavity(elem, 3, 4)
    x, y = md.x, md.y
    rho = 1 + .1 * np.sin(np.pi * x) * np.cos(np.pi * y)
    u = .2 * np.sin(np.pi * x) * np.sin(np.pi * y)
    v = .1 * np.cos(.5 * np.pi * x) * np.cos(.5 * np.pi * y) * (1 - y ** 2)
    p = 1 / (.09 * 1.4) * (1 + .05 * np.cos(np.pi * x))
    Q = list(P.primitive_to_conservative(rho, (u, v), p))
    par = cns.cavity_params(bctype=bctype)
    r, rt, rtv = cns.rhsRK(Q, md, ops, bc, par)
    assert all(np.isfinite(a).all() for a in r)
    assert rtv <= 1e-12           # viscous terms dissipate entropy
    assert rt <= 1e-12            # total residual (LF + viscous + BCs)
    # mass conservation with closed walls: sum of wJq * Vq * rhs_rho vanishes
    mass = np.sum(md.wJq * (ops.Vq @ r[0]))
    assert abs(mass) < 1e-12


def test_cavity_at_rest_only_feels_the_lid():
    rd, md, ops, bc, Q = cns.setup_cavity("quad", 3, 4)
    par = cns.cavity_params()
    r, rt, rtv = cns.rhsRK(Q, md, ops, bc, par)
    # the lid acts on the top row through the BCs and on the row below through the sigma jump; the
    # lower half of the cavity sees a uniform state with u = 0 at the walls: rhs = 0
    lid_elems = np.unique(bc.lid // md.nxJ.shape[0])
    others = np.where(md.y.mean(axis=0) < 0)[0]
    assert max(np.abs(a[:, others]).max() for a in r) < 1e-10
    assert max(np.abs(a[:, lid_elems]).max() for a in r[1:]) > 1e-2


def test_inviscid_part_matches_euler_tri_rhs_on_periodic_mesh():
    """With no boundary nodes and the viscous part off, rhs_inviscid is the rhs of dg2D_euler_tri.jl."""
    from oracle import solvers as SV
    rd, md, ops_e, Q = SV.setup_tri_vortex(3, 4, 4)
    ops = cns.build_cns_ops(rd)
    bc = SimpleNamespace(boundary=np.zeros(0, dtype=int), nx=np.zeros(0), ny=np.zeros(0))
    a = cns.rhs_inviscid(Q, md, ops, bc, True)
    b, _ = SV.rhs_tri(Q, md, ops_e, rd, .25, False)
    scale = max(np.abs(x).max() for x in b)
    assert max(np.abs(x - y).max() for x, y in zip(a, b)) < 1e-12 * scale


def test_golden_cns_fixture():
    g = np.load(os.path.join(GOLD, "cns_quad_N3_4x4_cavity.npz"))
    rd, md, ops, bc, _ = cns.setup_cavity("quad", 3, 4)
    par = cns.cavity_params()
    r, rt, rtv = cns.rhsRK(list(g["Q"]), md, ops, bc, par)
    scale = np.abs(g["rhs"]).max()
    assert np.abs(np.stack(r) - g["rhs"]).max() < 1e-12 * scale
    assert abs(rt - g["rhstest"]) < 1e-13 and abs(rtv - g["rhstest_visc"]) < 1e-13


def test_dopri45_coefficients_and_controller():
    """Butcher-table consistency (rows sum to c, 5th-order weights sum to 1, error weights to 0) and a short
    adaptive run of the cavity: steps are accepted, dt grows from the conservative start, entropy decays."""
    a, E, c = cns.dopri45_coeffs()
    assert np.abs(a.sum(axis=1) - c).max() < 1e-14
    assert abs(a[6].sum() - 1) < 1e-14 and abs(E.sum()) < 1e-15
    rd, md, ops, bc, Q = cns.setup_cavity("quad", 2, 3)
    par = cns.cavity_params()
    Q1, hist = cns.dopri45(Q, lambda q: cns.rhsRK(q, md, ops, bc, par), T=1e9, dt=1e-3, max_attempts=6)
    assert all(np.isfinite(q).all() for q in Q1)
    assert any(hist["accepted"]) and hist["t"][-1] > 0
    assert hist["dt"][-1] > hist["dt"][0]
    assert hist["rhstest"][-1] <= 1e-12
    # against a fine-step LSRK45 integration to the same time
    from oracle import solvers as SV
    T = hist["t"][-1]
    n = 40
    Qr, res = [q.copy() for q in Q], [np.zeros_like(q) for q in Q]
    for _ in range(n):
        Qr, res, _ = SV.lsrk45_step(Qr, res, T / n, lambda q, w: cns.rhsRK(q, md, ops, bc, par)[:2])
    assert max(np.abs(x - y).max() for x, y in zip(Q1, Qr)) < 1e-5


def test_shocktube_becker_solution_is_consistent_with_the_rhs():
    """Viscous shock tube (dg2D_CNS_modalESDG.jl): the Becker profile travels with v_inf, so the CNS rhs of the exact
    state must approach its time derivative as the mesh resolves the shock (mu = .02: thickness ~ .04)."""
    errs = []
    for K1D in (8, 16, 32):
        rd, md, ops, bc, Q, pr = cns.setup_shocktube("tri", 3, K1D)
        pr.mu, pr.lam = .02, 2 / 3 * .02
        pr.kappa = pr.mu * (1.4 / .4) / pr.Pr
        Q = cns.becker_exact(md.x, 0.0, pr)
        r, rt, _ = cns.rhsRK(Q, md, ops, bc, pr)
        d = 1e-5
        dQ = [(a - b) / (2 * d) for a, b in zip(cns.becker_exact(md.x, d, pr), cns.becker_exact(md.x, -d, pr))]
        num = np.sqrt(sum(np.sum(md.wJq * (ops.Vq @ (a - b)) ** 2) for a, b in zip(r, dQ)))
        den = np.sqrt(sum(np.sum(md.wJq * (ops.Vq @ b) ** 2) for b in dQ))
        errs.append(num / den)
        assert rt < 0                                       # entropy is dissipated
    assert errs[2] < .15 and errs[2] < .5 * errs[1] < .25 * errs[0]


def test_shocktube_states_and_node_sets():
    rd, md, ops, bc, Q, pr = cns.setup_shocktube("tri", 2, 4)
    # Rankine-Hugoniot: mass, momentum and energy fluxes of the two end states agree in the shock frame
    (rL, uL, _, pL), (rR, uR, _, pR) = pr.left, pr.right
    wL, wR = uL - pr.v_inf, uR - pr.v_inf
    assert abs(rL * wL - rR * wR) < 1e-14
    assert abs(rL * wL ** 2 + pL - rR * wR ** 2 - pR) < 1e-14
    hL, hR = 3.5 * pL / rL + .5 * wL ** 2, 3.5 * pR / rR + .5 * wR ** 2
    assert abs(hL - hR) < 1e-14
    xf = md.xf.flatten(order="F")
    assert np.allclose(xf[bc.inflow], -.5) and np.allclose(xf[bc.outflow], 1.0)
    assert bc.inflow.size == bc.outflow.size == 4 * 3        # K1D faces x (N+1) points
    # the exact profile connects the two states
    Ql = cns.becker_exact(np.array([-.5]), 0.0, pr)
    Qr = cns.becker_exact(np.array([1.0]), 0.0, pr)
    assert abs(Ql[0][0] - rL) < 1e-10 and abs(Qr[0][0] - rR) < 1e-10
