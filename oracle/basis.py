"""Reference-element bases and quadrature (oracle; test infrastructure only).

Restates ``src/Basis1D.jl``, ``src/Basis2DQuad.jl``, ``src/Basis3DHex.jl`` and
``src/Basis2DTri.jl`` of the reference.  ``meshgrid`` follows the MATLAB
semantics of the un-vendored ``VectorizedRoutines.Matlab.meshgrid``
(``X[i,j]=vx[j], Y[i,j]=vy[i]``; 3-arg likewise with ``Z[i,j,k]=vz[k]``) and
``vec`` is column-major, see SURVEY.md Appendix B.
"""
import json
import os
from math import gamma as _gamma, sqrt

import numpy as np


def meshgrid2(vx, vy=None):
    vy = vx if vy is None else vy
    X, Y = np.meshgrid(np.asarray(vx), np.asarray(vy), indexing="xy")
    return X, Y


def vec(A):
    return np.asarray(A).flatten(order="F")


# ---------------------------------------------------------------- Basis1D.jl
def jacobiP(x, alpha, beta, N):
    """Orthonormal Jacobi polynomial P_N^(alpha,beta) at x (Basis1D.jl:105-138)."""
    xp = np.asarray(x, dtype=np.float64).reshape(-1)
    PL = np.zeros((N + 1, xp.size))
    gamma0 = 2 ** (alpha + beta + 1) / (alpha + beta + 1) * _gamma(alpha + 1) * _gamma(beta + 1) / _gamma(alpha + beta + 1)
    PL[0, :] = 1.0 / sqrt(gamma0)
    if N == 0:
        return PL[0, :].copy()
    gamma1 = (alpha + 1) * (beta + 1) / (alpha + beta + 3) * gamma0
    PL[1, :] = ((alpha + beta + 2) * xp / 2 + (alpha - beta) / 2) / sqrt(gamma1)
    if N == 1:
        return PL[1, :].copy()
    aold = 2 / (2 + alpha + beta) * sqrt((alpha + 1) * (beta + 1) / (alpha + beta + 3))
    for i in range(1, N):
        h1 = 2 * i + alpha + beta
        anew = 2 / (h1 + 2) * sqrt((i + 1) * (i + 1 + alpha + beta) * (i + 1 + alpha) * (i + 1 + beta) / (h1 + 1) / (h1 + 3))
        bnew = -(alpha ** 2 - beta ** 2) / h1 / (h1 + 2)
        PL[i + 1, :] = 1 / anew * (-aold * PL[i - 1, :] + (xp - bnew) * PL[i, :])
        aold = anew
    return PL[N, :].copy()


def grad_jacobiP(r, alpha, beta, N):  # Basis1D.jl:88-94
    r = np.asarray(r, dtype=np.float64).reshape(-1)
    if N == 0:
        return np.zeros(r.size)
    return sqrt(N * (N + alpha + beta + 1)) * jacobiP(r, alpha + 1, beta + 1, N - 1)


def gauss_quad(alpha, beta, N):
    """Golub-Welsch Gauss-Jacobi rule with N+1 points (Basis1D.jl:59-77)."""
    if N == 0:
        return np.array([-(alpha - beta) / (alpha + beta + 2)]), np.array([2.0])
    n = np.arange(0, N + 1, dtype=np.float64)
    h1 = 2 * n + alpha + beta
    with np.errstate(divide="ignore", invalid="ignore"):
        d0 = -1 / 2 * (alpha ** 2 - beta ** 2) / (h1 + 2) / h1
    k = np.arange(1, N + 1, dtype=np.float64)
    d1 = 2 / (h1[:N] + 2) * np.sqrt(k * (k + alpha + beta) * (k + alpha) * (k + beta) / (h1[:N] + 1) / (h1[:N] + 3))
    J = np.diag(d0) + np.diag(d1, 1)
    if alpha + beta < 10 * np.finfo(float).eps:
        J[0, 0] = 0.0
    J = J + J.T
    x, V = np.linalg.eigh(J)
    w = V[0, :] ** 2 * 2 ** (alpha + beta + 1) / (alpha + beta + 1) * _gamma(alpha + 1) * _gamma(beta + 1) / _gamma(alpha + beta + 1)
    return x, w


def vandermonde_1D(N, r):  # Basis1D.jl:148-154
    r = np.asarray(r, dtype=np.float64).reshape(-1)
    V = np.zeros((r.size, N + 1))
    for j in range(N + 1):
        V[:, j] = jacobiP(r, 0, 0, j)
    return V


def grad_vandermonde_1D(N, r):  # Basis1D.jl:164-170
    r = np.asarray(r, dtype=np.float64).reshape(-1)
    V = np.zeros((r.size, N + 1))
    for j in range(N + 1):
        V[:, j] = grad_jacobiP(r, 0, 0, j)
    return V


def gauss_lobatto_quad(alpha, beta, N):  # Basis1D.jl:24-47
    if alpha != 0 and beta != 0:
        raise ValueError("alpha/beta not zero")
    if N == 0:
        return np.array([0.0]), np.array([2.0])
    if N == 1:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    xint, _ = gauss_quad(alpha + 1, beta + 1, N - 2)
    x = np.concatenate([[-1.0], xint, [1.0]])
    V = vandermonde_1D(N, x)
    w = np.sum(np.linalg.inv(V @ V.T), axis=1)
    return x, w


# ------------------------------------------------------------ Basis2DQuad.jl
def quad_vandermonde_2D(N, r, s):  # Basis2DQuad.jl:25-38
    V = np.zeros((len(r), (N + 1) ** 2))
    sk = 0
    for i in range(N + 1):
        for j in range(N + 1):
            V[:, sk] = jacobiP(r, 0, 0, i) * jacobiP(s, 0, 0, j)
            sk += 1
    return V


def quad_grad_vandermonde_2D(N, r, s):  # Basis2DQuad.jl:49-64
    Vr = np.zeros((len(r), (N + 1) ** 2))
    Vs = np.zeros((len(r), (N + 1) ** 2))
    sk = 0
    for i in range(N + 1):
        for j in range(N + 1):
            Vr[:, sk] = grad_jacobiP(r, 0, 0, i) * jacobiP(s, 0, 0, j)
            Vs[:, sk] = jacobiP(r, 0, 0, i) * grad_jacobiP(s, 0, 0, j)
            sk += 1
    return Vr, Vs


def quad_nodes_2D(N):  # Basis2DQuad.jl:77-81  (r fastest)
    r1D, _ = gauss_lobatto_quad(0, 0, N)
    s, r = meshgrid2(r1D)
    return vec(r), vec(s)


def quad_equi_nodes_2D(N):  # Basis2DQuad.jl:93-98
    r1D = np.linspace(-1, 1, N + 1)
    s, r = meshgrid2(r1D)
    return vec(r), vec(s)


def quad_quad_nodes_2D(N):  # Basis2DQuad.jl:110-116
    r1D, w1D = gauss_quad(0, 0, N)
    s, r = meshgrid2(r1D)
    ws, wr = meshgrid2(w1D)
    return vec(r), vec(s), vec(wr * ws)


# ------------------------------------------------------------- Basis3DHex.jl
def hex_vandermonde_3D(N, r, s, t):  # Basis3DHex.jl:25-39
    V = np.zeros((len(r), (N + 1) ** 3))
    sk = 0
    for i in range(N + 1):
        for j in range(N + 1):
            for k in range(N + 1):
                V[:, sk] = jacobiP(r, 0, 0, i) * jacobiP(s, 0, 0, j) * jacobiP(t, 0, 0, k)
                sk += 1
    return V


def hex_grad_vandermonde_3D(N, r, s, t):  # Basis3DHex.jl:47-66
    n = (N + 1) ** 3
    Vr = np.zeros((len(r), n))
    Vs = np.zeros((len(r), n))
    Vt = np.zeros((len(r), n))
    sk = 0
    for i in range(N + 1):
        for j in range(N + 1):
            for k in range(N + 1):
                Pi, Pj, Pk = jacobiP(r, 0, 0, i), jacobiP(s, 0, 0, j), jacobiP(t, 0, 0, k)
                Vr[:, sk] = grad_jacobiP(r, 0, 0, i) * Pj * Pk
                Vs[:, sk] = Pi * grad_jacobiP(s, 0, 0, j) * Pk
                Vt[:, sk] = Pi * Pj * grad_jacobiP(t, 0, 0, k)
                sk += 1
    return Vr, Vs, Vt


def tensor3_r_fastest(v1d):
    """(r,s,t) tensor grid, r fastest then s then t.

    DEVIATION: the reference's ``vec.(meshgrid(r1D,r1D,r1D))``
    (Basis3DHex.jl:77-80, SetupDG.jl:362-364) makes *s* fastest, which is
    inconsistent with the x-fastest vertex ordering of UniformHexMesh.jl:59-74
    and yields J<0; the author marks the 3D script broken.  We order r fastest
    (as Basis2DQuad.nodes_2D does) so that J>0.
    """
    v1d = np.asarray(v1d)
    n = v1d.size
    a, b, c = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    # flatten with a fastest
    a, b, c = a.flatten(order="F"), b.flatten(order="F"), c.flatten(order="F")
    return v1d[a], v1d[b], v1d[c]


def hex_nodes_3D(N):
    r1D, _ = gauss_lobatto_quad(0, 0, N)
    return tensor3_r_fastest(r1D)


# ------------------------------------------------------------- Basis2DTri.jl
def tri_rstoab(r, s):  # Basis2DTri.jl:77-87
    r = np.asarray(r, dtype=np.float64).reshape(-1)
    s = np.asarray(s, dtype=np.float64).reshape(-1)
    a = np.full(r.size, -1.0)
    m = s != 1
    a[m] = 2 * (1 + r[m]) / (1 - s[m]) - 1
    return a, s


def tri_simplex_2D(a, b, i, j):  # Basis2DTri.jl:25-29
    h1 = jacobiP(a, 0, 0, i)
    h2 = jacobiP(b, 2 * i + 1, 0, j)
    return sqrt(2.0) * h1 * h2 * (1 - b) ** i


def tri_grad_simplex_2D(a, b, id, jd):  # Basis2DTri.jl:42-66
    fa = jacobiP(a, 0, 0, id)
    gb = jacobiP(b, 2 * id + 1, 0, jd)
    dfa = grad_jacobiP(a, 0, 0, id)
    dgb = grad_jacobiP(b, 2 * id + 1, 0, jd)
    dmodedr = dfa * gb
    if id > 0:
        dmodedr = dmodedr * ((0.5 * (1 - b)) ** (id - 1))
    dmodeds = dfa * (gb * (0.5 * (1 + a)))
    if id > 0:
        dmodeds = dmodeds * ((0.5 * (1 - b)) ** (id - 1))
    tmp = dgb * ((0.5 * (1 - b)) ** id)
    if id > 0:
        tmp = tmp - 0.5 * id * gb * ((0.5 * (1 - b)) ** (id - 1))
    dmodeds = dmodeds + fa * tmp
    dmodedr = 2 ** (id + 0.5) * dmodedr
    dmodeds = 2 ** (id + 0.5) * dmodeds
    return dmodedr, dmodeds


def tri_vandermonde_2D(N, r, s):  # Basis2DTri.jl:98-110
    Np = (N + 1) * (N + 2) // 2
    V = np.zeros((len(r), Np))
    a, b = tri_rstoab(r, s)
    sk = 0
    for i in range(N + 1):
        for j in range(N - i + 1):
            V[:, sk] = tri_simplex_2D(a, b, i, j)
            sk += 1
    return V


def tri_grad_vandermonde_2D(N, r, s):  # Basis2DTri.jl:122-137
    Np = (N + 1) * (N + 2) // 2
    Vr = np.zeros((len(r), Np))
    Vs = np.zeros((len(r), Np))
    a, b = tri_rstoab(r, s)
    sk = 0
    for i in range(N + 1):
        for j in range(N - i + 1):
            Vr[:, sk], Vs[:, sk] = tri_grad_simplex_2D(a, b, i, j)
            sk += 1
    return Vr, Vs


def tri_xytors(x, y):  # Basis2DTri.jl:151-158
    L1 = (sqrt(3.0) * y + 1.0) / 3.0
    L2 = (-3.0 * x - sqrt(3.0) * y + 2.0) / 6.0
    L3 = (3.0 * x - sqrt(3.0) * y + 2.0) / 6.0
    return -L2 + L3 - L1, -L2 - L3 + L1


def tri_warp_factor(N, rout):  # Basis2DTri.jl:170-186
    LGLr, _ = gauss_lobatto_quad(0, 0, N)
    req = np.linspace(-1, 1, N + 1)
    Veq = vandermonde_1D(N, req)
    Pmat = np.zeros((N + 1, rout.size))
    for i in range(N + 1):
        Pmat[i, :] = jacobiP(rout, 0, 0, i)
    Lmat = np.linalg.solve(Veq.T, Pmat)
    warp = Lmat.T @ (LGLr - req)
    zerof = (np.abs(rout) < 1.0 - 1.0e-10).astype(np.float64)
    sf = 1.0 - (zerof * rout) ** 2
    return warp / sf + warp * (zerof - 1)


_ALPOPT = [0.0, 0.0, 1.4152, 0.1001, 0.2751, 0.98, 1.0999, 1.2832, 1.3648,
           1.4773, 1.4959, 1.5743, 1.577, 1.6223, 1.6258]  # Basis2DTri.jl:201


def tri_nodes_2D(N):  # Basis2DTri.jl:197-233
    Np = (N + 1) * (N + 2) // 2
    alpha = _ALPOPT[N - 1] if N < 16 else 5 / 3
    L1 = np.zeros(Np)
    L3 = np.zeros(Np)
    sk = 0
    for n in range(1, N + 2):
        for m in range(1, N + 3 - n):
            L1[sk] = (n - 1) / N
            L3[sk] = (m - 1) / N
            sk += 1
    L2 = 1.0 - L1 - L3
    x = -L2 + L3
    y = (-L2 - L3 + 2 * L1) / sqrt(3.0)
    blend1 = 4.0 * L2 * L3
    blend2 = 4.0 * L1 * L3
    blend3 = 4.0 * L1 * L2
    warpf1 = tri_warp_factor(N, L3 - L2)
    warpf2 = tri_warp_factor(N, L1 - L3)
    warpf3 = tri_warp_factor(N, L2 - L1)
    warp1 = blend1 * warpf1 * (1 + (alpha * L1) ** 2)
    warp2 = blend2 * warpf2 * (1 + (alpha * L2) ** 2)
    warp3 = blend3 * warpf3 * (1 + (alpha * L3) ** 2)
    x = x + 1 * warp1 + np.cos(2 * np.pi / 3) * warp2 + np.cos(4 * np.pi / 3) * warp3
    y = y + 0 * warp1 + np.sin(2 * np.pi / 3) * warp2 + np.sin(4 * np.pi / 3) * warp3
    return tri_xytors(x, y)


def tri_equi_nodes_2D(N):  # Basis2DTri.jl:245-262
    r1D = np.linspace(-1, 1, N + 1)
    r, s = [], []
    for i in range(N + 1):
        for j in range(N - i + 1):
            r.append(r1D[i])
            s.append(r1D[j])
    return np.array(r), np.array(s)


_TRI_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tri_cubature.json")
_tri_cache = None


def tri_quad_nodes_2D(N):
    """Triangle cubature exact to degree N (Basis2DTri.jl:274-296).

    The reference reads ``src/QuadratureData/quad_nodes_tri_N{N}.txt``; the
    numerical tables (a published cubature rule, data not code) were extracted
    once by ``oracle/tools/extract_tri_cubature.py`` into
    ``oracle/data/tri_cubature.json`` so the oracle runs where
    ``/root/reference`` does not exist.
    """
    global _tri_cache
    if N < 28:
        if _tri_cache is None:
            with open(_TRI_TABLE) as fh:
                _tri_cache = json.load(fh)
        rsw = np.array(_tri_cache[str(N)], dtype=np.float64)
        return rsw[:, 0].copy(), rsw[:, 1].copy(), rsw[:, 2].copy()
    cubNA = int(np.ceil((N + 1) / 2))
    cubNB = int(np.ceil((N + 1) / 2))
    cubA, cubWA = gauss_quad(0, 0, cubNA - 1)
    cubB, cubWB = gauss_quad(1, 0, cubNB - 1)
    A = np.ones((cubNB, 1)) * cubA[None, :]
    B = cubB[:, None] * np.ones((1, cubNA))
    r = 0.5 * (1 + A) * (1 - B) - 1
    s = B
    w = 0.5 * cubWB[:, None] * cubWA[None, :]
    return vec(r), vec(s), vec(w)
