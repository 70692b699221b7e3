"""Port of the reference's only test suite, examples/EntropyStableEuler.jl/test/runtests.jl,
applied to the oracle's restatement of the LOCAL physics module (the one the scripts import).

runtests.jl:4-9   logmean symmetry / consistency
runtests.jl:55-127 (2D) and :130-194 (3D): entropy variables = dS/dU (central differences, h=1e-7),
u_vfun(v_ufun(U)) = U, flux symmetry, consistency F(U,U)=f(U), entropy conservation
(vL-vR).F = psiL-psiR with psi = (gamma-1) rho u_i  [entropy_scaling = 1 for the local module].
"""
import numpy as np
import pytest

from oracle import physics as P

G = P.GAMMA


def test_logmean():  # runtests.jl:4-9
    uL, uR = 1.0, 2.0
    assert P.logmean(uL, uR) == P.logmean(uL, uR, np.log(uL), np.log(uR))
    assert P.logmean(uL, uR) == P.logmean(uR, uL)
    assert np.isclose(P.logmean(uL, uL), uL)
    assert np.isclose(P.logmean(uL, uR), (uR - uL) / np.log(uR / uL), rtol=1e-15)


def test_logmean_taylor_branch_constants():
    """H5: |f|<1e-4 uses aavg*(1+v*(-.2-v*(.0512-v*0.026038857142857))) -- replicated as written."""
    a = 1.0
    b = a * (1 + 5e-5)
    f = (b - a) / (.5 * (a + b))
    v = f * f
    assert P.logmean(a, b) == .5 * (a + b) * (1 + v * (-.2 - v * (.0512 - v * 0.026038857142857)))


def _states(dim):
    UL = P.primitive_to_conservative(1.0, (.1, .2, .3)[:dim], 2.0)
    UR = P.primitive_to_conservative(1.1, (.2, .3, .4)[:dim], 2.1)
    return UL, UR


def _prim_beta(U):
    return (U[0],) + tuple(m / U[0] for m in U[1:-1]) + (P.beta_of_u(U),)


@pytest.mark.parametrize("dim", [2, 3])
def test_entropy_variables_are_gradient_of_entropy(dim):
    U = np.array(_states(dim)[0], dtype=float)
    V = P.v_of_u(tuple(U))
    h = 1e-7
    for k in range(dim + 2):
        Up, Um = U.copy(), U.copy()
        Up[k] += h
        Um[k] -= h
        S = lambda W: P.Sfun(W[0], tuple(W[1:-1]), W[-1])
        # local module: S = -rho*s (euler_variables.jl:71-76), v = dS/dU with no 1/(gamma-1) scaling
        assert abs(V[k] - (S(Up) - S(Um)) / (2 * h)) < 10 * h


@pytest.mark.parametrize("dim", [2, 3])
def test_u_v_roundtrip(dim):
    UL, _ = _states(dim)
    U2 = P.u_of_v(P.v_of_u(UL))
    assert np.allclose(U2, UL, rtol=1e-13)


@pytest.mark.parametrize("dim", [2, 3])
def test_flux_symmetry_consistency_entropy_conservation(dim):
    UL, UR = _states(dim)
    QL, QR = _prim_beta(UL), _prim_beta(UR)
    F = P.euler_fluxes(QL, QR)
    F2 = P.euler_fluxes(QR, QL)
    for a, b in zip(F, F2):
        assert np.array_equal(np.array(a, dtype=float), np.array(b, dtype=float))   # bitwise symmetric
    # consistency
    rho, vel, p = UL[0], [m / UL[0] for m in UL[1:-1]], P.pfun(UL[0], tuple(UL[1:-1]), UL[-1])
    FL = P.euler_fluxes(QL, QL)
    for d in range(dim):
        exact = [rho * vel[d]] + [rho * vel[d] * vel[k] + (p if k == d else 0.0) for k in range(dim)] + [vel[d] * (UL[-1] + p)]
        assert np.allclose(np.array(FL[d], dtype=float), exact, rtol=1e-13)
    # entropy conservation: (vL - vR)^T F_d = psi_d(UL) - psi_d(UR), psi_d = (gamma-1) rho u_d
    VL, VR = P.v_of_u(UL), P.v_of_u(UR)
    for d in range(dim):
        lhs = sum((a - b) * f for a, b, f in zip(VL, VR, F[d]))
        rhs = (G - 1) * (UL[1 + d] - UR[1 + d])
        assert np.isclose(lhs, rhs, rtol=1e-12, atol=1e-14)


def test_flux_known_answer_vector():
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "flux3d_kat.npz"))
    F = np.array(P.euler_fluxes_3d(tuple(d["QL"]), tuple(d["QR"])), dtype=float)
    assert np.allclose(F, d["F"], rtol=1e-14, atol=0)


def test_wavespeed_replicates_reference_quirk():
    """H4: wavespeed = sqrt(|rho u_n / rho|) + c with p from the normal momentum only."""
    rho, rhoun, E = 1.3, .4, 2.2
    p = (G - 1) * (E - .5 * rhoun ** 2 / rho)
    assert np.isclose(P.wavespeed(rho, rhoun, E), np.sqrt(abs(rhoun / rho)) + np.sqrt(G * p / rho), rtol=1e-15)


def test_vortex_is_isentropic():
    x, y = np.meshgrid(np.linspace(0, 10, 7), np.linspace(-4, 4, 5))
    rho, u, v, p = P.vortex(x, y, .3)
    assert np.allclose(p, rho ** G)
