"""The identity the tensor-product CNS path rests on (esdg_cns_b200/csrc/kernels_cnsq.cuh), checked on the CPU with the
oracle: on a quad the Gauss quadrature has as many points as the basis has functions, so with Q~ = Vq Q every operator
of rhsRK! (dg2D_CNS_cavity_optimized.jl:955-972) becomes a Kronecker product of 1-D Gauss-node tables, and evaluating
the oracle with those nodal operators gives Vq * (the oracle's rhs)."""
from types import SimpleNamespace

import numpy as np
import pytest


@pytest.mark.parametrize("N,bctype", [(2, 1), (3, 1), (3, 2), (2, 3)])
def test_rhsRK_in_the_gauss_nodal_basis_equals_the_modal_one(N, bctype):
    from oracle import cns
    from esdg_cns_b200 import setup_dg as SD
    rd, md, ops, bc, Q = cns.setup_cavity("quad", N, 4)
    par = cns.cavity_params(Re=100.0, bctype=bctype)
    rng = np.random.default_rng(3)
    x, y = md.x, md.y
    rho = 1 + .1 * np.sin(2 * x) * np.cos(y) + .01 * rng.random(x.shape)
    u = .3 * np.sin(x + .5) * np.cos(2 * y)
    v = -.2 * np.cos(x) * np.sin(y + .3)
    p = 1 / (.3 ** 2 * 1.4) * (1 + .05 * np.cos(x - y))
    from oracle import physics as P
    Q = list(P.primitive_to_conservative(rho, (u, v), p))
    ref, rt, rtv = cns.rhsRK(Q, md, ops, bc, par)

    Vq, Pq = ops.Vq, ops.Pq
    n = N + 1
    assert Vq.shape == (n * n, n * n)
    assert np.abs(Vq @ Pq - np.eye(n * n)).max() < 1e-12
    # the same operators in the nodal basis ...
    nod = SimpleNamespace(**vars(ops))
    nod.Vq = np.eye(n * n)
    nod.Pq = np.eye(n * n)
    nod.Vf = ops.Vf @ Pq
    nod.VhP = np.vstack([np.eye(n * n), nod.Vf])
    nod.Dr, nod.Ds = Vq @ ops.Dr @ Pq, Vq @ ops.Ds @ Pq
    nod.Lf = Vq @ ops.Lf
    nod.Ph = Vq @ ops.Ph
    # ... are the Kronecker products the kernels hard-wire (node k = i + n j, s = x[i] fastest; faces CCW, last two reversed)
    prd, pops = SD.reference_quad(N)
    assert np.abs(nod.Vf - pops.Ef).max() < 1e-12
    assert np.abs(nod.Lf - pops.Lf).max() < 1e-11
    assert np.abs(nod.Ph - pops.Ph).max() < 1e-11
    xg, _ = SD.gauss_1d(N)
    D1 = SD.lagrange_deriv(xg)
    assert np.abs(nod.Dr - np.kron(D1, np.eye(n))).max() < 1e-11       # d/dr acts on j (slow index)
    assert np.abs(nod.Ds - np.kron(np.eye(n), D1)).max() < 1e-11
    Qn = [Vq @ q for q in Q]
    out, rt_n, rtv_n = cns.rhsRK(Qn, md, nod, bc, par)
    scale = max(np.abs(r).max() for r in ref)
    err = max(np.abs(Pq @ a - b).max() for a, b in zip(out, ref)) / scale
    print(f"N={N} bctype={bctype}: nodal vs modal rhsRK relerr {err:.2e}; rhstest {rt:.3e}/{rt_n:.3e} visc {rtv:.3e}/{rtv_n:.3e}")
    assert err < 1e-12
    assert abs(rt - rt_n) < 1e-11 * max(1.0, abs(rt)) and abs(rtv - rtv_n) < 1e-11 * max(1.0, abs(rtv))
