"""Oracle self-checks: SBP identities, map involution, entropy residual at round-off,
free-stream preservation, vortex convergence, agreement of the two independent CPU codings, and
regression against the committed golden fixtures."""
import os

import numpy as np
import pytest

from oracle import physics as P, solvers as SV, setupdg as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("elem,N", [("quad", 2), ("quad", 3), ("hex", 1), ("hex", 2), ("tri", 2), ("tri", 4)])
def test_hybridized_sbp_properties(elem, N):
    rd = {"quad": S.init_reference_quad, "hex": S.init_reference_hex, "tri": S.init_reference_tri}[elem](N)
    ops = SV.build_sbp_ops(rd)
    Nq = rd.wq.size
    for Qk in ops.Qskew:
        assert np.abs(Qk + Qk.T).max() < 1e-13                 # skew symmetric
        assert np.abs(Qk[Nq:, Nq:]).max() < 1e-13              # face-face block vanishes
    # Q_h 1 = 0 for the un-skewed hybridized operator  <=>  skew part rows sum to -(1/2) B 1 ... check via Qh:
    Ds = [rd.Dr, rd.Ds] + ([rd.Dt] if elem == "hex" else [])
    ns = [rd.nrJ, rd.nsJ] + ([rd.ntJ] if elem == "hex" else [])
    for D, n in zip(Ds, ns):
        Q = rd.Pq.T @ rd.M @ D @ rd.Pq
        B = np.diag(rd.wf * n)
        Qh = .5 * np.block([[Q - Q.T, ops.Ef.T @ B], [-B @ ops.Ef, B]])
        assert np.abs(Qh @ np.ones(Qh.shape[0])).max() < 1e-12     # exact differentiation of constants
        Bh = np.zeros_like(Qh)
        Bh[Nq:, Nq:] = B
        assert np.abs(Qh + Qh.T - Bh).max() < 1e-12                # SBP: Q + Q' = B


def test_sizes_match_survey_table():
    rd = S.init_reference_quad(3)
    ops = SV.build_sbp_ops(rd)
    assert sum(len(r) for r in ops.Qnzids) == 224
    rd = S.init_reference_hex(4)
    ops = SV.build_sbp_ops(rd)
    assert sum(len(r) for r in ops.Qnzids) == 3000
    assert sorted(set(len(r) for r in ops.Qnzids)) == [5, 18]


def test_mapP_is_an_involution_without_fixed_points():
    for md in (SV.setup_quad_vortex(2, 4, 3)[1], SV.setup_hex(1, 3)[1], SV.setup_tri_vortex(2, 3, 3)[1]):
        mp = md.mapP.flatten(order="F")
        assert np.array_equal(mp[mp], np.arange(mp.size))
        assert not np.any(mp == np.arange(mp.size))
        assert md.J.min() > 0


def test_quad_config1_entropy_and_freestream():
    rd, md, ops, Q = SV.setup_quad_vortex(3, 16, 16)
    _, rt_ec = SV.rhs_collocated(Q, md, ops, rd, 0.0, True)
    assert abs(rt_ec) < 1e-12                                   # entropy conservative without LF
    _, rt = SV.rhs_collocated(Q, md, ops, rd, .5, True)
    assert rt < 0                                               # dissipative with LF
    Qc = [np.full_like(Q[0], v) for v in (1.0, .3, -.2, 2.5)]
    r, _ = SV.rhs_collocated(Qc, md, ops, rd, .5)
    assert max(np.abs(x).max() for x in r) < 1e-11


def test_hex_entropy_conservation_random_state():
    """The reference's own EC experiment (dg3D_euler_hex.jl:105-110,224-226) with the H1/H2 fix."""
    rd, md, ops = SV.setup_hex(2, 3)
    rng = np.random.default_rng(0)
    rho = 2 + .1 * rng.random(md.xq.shape)
    p = 1 + .1 * rng.random(rho.shape)
    Q = list(P.primitive_to_conservative(rho, (np.zeros_like(rho), np.ones_like(rho), np.zeros_like(rho)), p))
    _, rt = SV.rhs_collocated(Q, md, ops, rd, 0.0, True)
    assert abs(rt) < 1e-13
    _, rt = SV.rhs_collocated(Q, md, ops, rd, .25, True)
    assert rt < 0


def test_curved_hex_entropy_conservation_and_freestream():
    """Curved (trilinear) hexes, the per-point geometric-factor branch of dg3D_euler_hex.jl:69-101,145-151."""
    rd, md, ops = SV.setup_hex_curved(2, 3)
    assert np.abs(md.rxJ - md.rxJ[:1]).max() > 1e-3 and md.J.min() > 0
    one = np.ones_like(md.xq)
    Qc = list(P.primitive_to_conservative(1.2 * one, (.3 * one, -.2 * one, .1 * one), .9 * one))
    r, _ = SV.rhs_collocated(Qc, md, ops, rd, .25)
    assert max(np.abs(x).max() for x in r) < 1e-12
    rng = np.random.default_rng(0)
    rho = 2 + .1 * rng.random(md.xq.shape)
    p = 1 + .1 * rng.random(rho.shape)
    Q = list(P.primitive_to_conservative(rho, (0 * one, one, 0 * one), p))
    _, rt = SV.rhs_collocated(Q, md, ops, rd, 0.0, True)
    assert abs(rt) < 1e-13
    _, rt = SV.rhs_collocated(Q, md, ops, rd, .25, True)
    assert rt < 0


def test_tri_entropy_and_freestream():
    rd, md, ops, Q = SV.setup_tri_vortex(3, 4, 4)
    _, rt = SV.rhs_tri(Q, md, ops, rd, 0.0, True)
    assert abs(rt) < 1e-11
    Qc = [np.full_like(Q[0], v) for v in (1.0, .3, -.2, 2.5)]
    r, _ = SV.rhs_tri(Qc, md, ops, rd, .25)
    assert max(np.abs(x).max() for x in r) < 1e-11


def test_vortex_convergence_quad():
    """L2 error of the vortex after a short time decreases at ~N+1 under refinement."""
    errs = []
    N = 2
    for K1D in (8, 16):
        rd, md, ops, Q = SV.setup_quad_vortex(N, K1D, K1D)
        T = .1
        dt = .5 * (2 / K1D) / ((N + 1) * (N + 2) / 2)
        nsteps = int(np.ceil(T / dt))
        dt = T / nsteps
        res = [np.zeros_like(q) for q in Q]
        for _ in range(nsteps):
            Q, res, _ = SV.lsrk45_step(Q, res, dt, lambda q, t: SV.rhs_collocated(q, md, ops, rd, .5, t))
        ex = P.primitive_to_conservative(*(lambda r, u, v, p: (r, (u, v), p))(*P.vortex(md.xq, md.yq, T)))
        errs.append(np.sqrt(sum(np.sum(md.wJq * (a - b) ** 2) for a, b in zip(Q, ex))))
    rate = np.log2(errs[0] / errs[1])
    assert rate > N + .3, (errs, rate)


def test_c_and_numpy_oracles_agree():
    from helpers import relerr, parity_tol
    from oracle import c_oracle
    rd, md, ops, Q = SV.setup_quad_vortex(3, 8, 8)
    ref, rt = SV.rhs_collocated(Q, md, ops, rd, .5, True)
    out, rtc = c_oracle.COracle(md, ops).rhs(Q, True)
    assert relerr(out, ref) < parity_tol(lambda: SV.rhs_collocated(Q, md, ops, rd, .5)[0], ref)
    assert abs(rt - rtc) < 1e-13
    rd, md, ops = SV.setup_hex(2, 3, scale=(np.pi,) * 3)
    Q = list(P.primitive_to_conservative(*(lambda r, u, v, w, p: (r, (u, v, w), p))(*P.taylor_green(md.xq, md.yq, md.zq))))
    ref, rt = SV.rhs_collocated(Q, md, ops, rd, .25, True)
    co = c_oracle.COracle(md, ops)
    out, rtc = co.rhs(Q, True)
    assert relerr(out, ref) < parity_tol(lambda: SV.rhs_collocated(Q, md, ops, rd, .25)[0], ref)
    Q1, _, _ = SV.lsrk45_step(Q, [np.zeros_like(q) for q in Q], 1e-3, lambda q, t: SV.rhs_collocated(q, md, ops, rd, .25, t))
    Qc, _ = co.lsrk45(Q, 1e-3, 1)
    assert max(np.abs(a - b).max() / max(1.0, np.abs(b).max()) for a, b in zip(Qc, Q1)) < 1e-12


@pytest.mark.parametrize("name", ["quad_N3_4x4_vortex", "hex_N2_3x3x3_tgv", "tri_N3_3x2_vortex"])
def test_oracle_reproduces_golden_fixtures(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    Q = list(d["Q"])
    if name.startswith("quad"):
        rd, md, ops, _ = SV.setup_quad_vortex(3, 4, 4)
        f = lambda q, t=False: SV.rhs_collocated(q, md, ops, rd, .5, t)
    elif name.startswith("hex"):
        rd, md, ops = SV.setup_hex(2, 3, scale=(np.pi,) * 3)
        f = lambda q, t=False: SV.rhs_collocated(q, md, ops, rd, .25, t)
    else:
        rd, md, ops, _ = SV.setup_tri_vortex(3, 3, 2)
        f = lambda q, t=False: SV.rhs_tri(q, md, ops, rd, .25, t)
    rhs, rt = f(Q, True)
    assert np.allclose(np.stack(rhs), d["rhs"], rtol=1e-11, atol=1e-12)
    assert abs(rt - float(d["rhstest"])) < 1e-12
    if "Q1" in d:
        Q1, _, rt1 = SV.lsrk45_step(Q, [np.zeros_like(q) for q in Q], float(d["dt"]), f)
        assert np.allclose(np.stack(Q1), d["Q1"], rtol=1e-13, atol=1e-14)
