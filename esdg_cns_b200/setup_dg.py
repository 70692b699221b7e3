"""Host-side setup for uniform quad / triangle / hex meshes (not on the hot path).

In a Julia deployment ``rd``/``md``/``ops`` come from the reference's own
``SetupDG.jl`` and script-level operator assembly and are handed to the library
through the C ABI.  This module plays that role for the Python host, tests and
benchmarks.  It is written tensor-product-first (everything derives from the
1D Gauss rule, the 1D Lagrange derivative matrix and the two end-point
interpolation vectors) and builds connectivity analytically, so a 128^3 mesh
sets up in seconds -- the reference's generic ``build_node_maps`` is
O(faces*Nfp^2) with dense distance matrices (src/node_map_functions.jl:35-51).
Array names, shapes and orderings are the reference's:

* quad: volume nodes s-fastest (src/SetupDG.jl:244-246), faces CCW (:235-239)
* hex : volume nodes r-fastest (the H2 fix), faces r-,r+,s-,s+,t-,t+ with the
  on-face ordering of src/SetupDG.jl:345-357
* ``mapP`` is 0-based here; the binding adds 1 (Julia convention) at the ABI.

tests/test_setup_parity.py checks every array against the oracle's literal
restatement of SetupDG.jl on small meshes.
"""
from types import SimpleNamespace

import numpy as np


# ------------------------------------------------------------------ 1D pieces
def gauss_1d(N):
    x, w = np.polynomial.legendre.leggauss(N + 1)
    return x, w


def lagrange_eval(x, y):
    """L[i,j] = l_j(y_i) for the Lagrange basis on nodes x."""
    x = np.asarray(x, dtype=np.float64)
    y = np.atleast_1d(np.asarray(y, dtype=np.float64))
    n = x.size
    L = np.ones((y.size, n))
    for j in range(n):
        for m in range(n):
            if m != j:
                L[:, j] *= (y - x[m]) / (x[j] - x[m])
    return L


def lagrange_deriv(x):
    """D[i,j] = l_j'(x_i) (barycentric form)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    c = np.array([np.prod([x[j] - x[m] for m in range(n) if m != j]) for j in range(n)])
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = c[i] / (c[j] * (x[i] - x[j]))
        D[i, i] = -np.sum(D[i, :])
    return D


# --------------------------------------------------------- reference elements
def _hybridize(Qs, Ef, wf, normals):
    out = []
    for Q, n in zip(Qs, normals):
        B = np.diag(wf * n)
        Qh = .5 * np.block([[Q - Q.T, Ef.T @ B], [-B @ Ef, B]])   # dg2D_euler_quad.jl:53-56
        out.append(.5 * (Qh - Qh.T))                               # :60-61
    return out


def reference_quad(N, lf_scale=.5):
    """Collocated Gauss quad element in the quadrature-node basis (dg2D_euler_quad.jl:47-78)."""
    x, w = gauss_1d(N)
    n = N + 1
    D = lagrange_deriv(x)
    lm, lp = lagrange_eval(x, -1.0)[0], lagrange_eval(x, 1.0)[0]
    I = np.eye(n)
    # k = i + n*j with s = x[i] (fastest), r = x[j]
    rq, sq = np.repeat(x, n), np.tile(x, n)
    wq = np.repeat(w, n) * np.tile(w, n)
    Dr = np.kron(D, I)          # acts on j
    Ds = np.kron(I, D)          # acts on i
    Ef = np.zeros((4 * n, n * n))
    k = lambda i, j: i + n * j
    for m in range(n):
        for a in range(n):
            Ef[m, k(a, m)] = lm[a]                      # f1: s=-1, r=x[m]
            Ef[n + m, k(m, a)] = lp[a]                  # f2: r=+1, s=x[m]
            Ef[2 * n + m, k(a, n - 1 - m)] = lp[a]      # f3: s=+1, r=-x[m]
            Ef[3 * n + m, k(n - 1 - m, a)] = lm[a]      # f4: r=-1, s=-x[m]
    e, z = np.ones(n), np.zeros(n)
    rd = SimpleNamespace(elem="quad", N=N, Nfaces=4, rq=rq, sq=sq, wq=wq, wf=np.tile(w, 4),
                         rf=np.concatenate([x, e, -x, -e]), sf=np.concatenate([-e, x, e, -x]),
                         nrJ=np.concatenate([z, e, z, -e]), nsJ=np.concatenate([-e, z, e, z]))
    W = np.diag(wq)
    Qskew = _hybridize([W @ Dr, W @ Ds], Ef, rd.wf, [rd.nrJ, rd.nsJ])
    Vh = np.vstack([np.eye(n * n), Ef])
    ops = SimpleNamespace(elem="quad", N=N, Nfaces=4, wq=wq, Ef=Ef, Qskew=Qskew, Vh=Vh,
                          Ph=np.diag(1 / wq) @ Vh.T, Lf=np.diag(1 / wq) @ (Ef.T @ np.diag(rd.wf)),
                          volume_scale=2.0, lf_scale=lf_scale)
    return rd, ops


def reference_hex(N, lf_scale=.25):
    """Collocated Gauss hex element, r-fastest nodes (dg3D_euler_hex.jl:34-58,94-98 with the H2 fix).

    ``lf_scale`` is the constant c in ``LFc = c*max(lam, lamP)*sJ``.  NOTE the default: the upstream hex script
    multiplies its penalty by zero (``LFc = 0*.25*...``, dg3D_euler_hex.jl:193), i.e. it runs entropy CONSERVATIVE;
    the default here is the .25 the formula carries (the value of the triangle / CNS scripts and of BASELINE config 5),
    which makes the rhs entropy DISSIPATIVE.  Pass ``lf_scale=0.0`` for the upstream script's behaviour."""
    x, w = gauss_1d(N)
    n = N + 1
    D = lagrange_deriv(x)
    lm, lp = lagrange_eval(x, -1.0)[0], lagrange_eval(x, 1.0)[0]
    I = np.eye(n)
    # k = a + n*b + n^2*c with r=x[a], s=x[b], t=x[c]
    rq = np.tile(x, n * n)
    sq = np.tile(np.repeat(x, n), n)
    tq = np.repeat(x, n * n)
    wq = np.tile(w, n * n) * np.tile(np.repeat(w, n), n) * np.repeat(w, n * n)
    Dr = np.kron(I, np.kron(I, D))
    Ds = np.kron(I, np.kron(D, I))
    Dt = np.kron(D, np.kron(I, I))
    nfp = n * n
    Ef = np.zeros((6 * nfp, n ** 3))
    k = lambda a, b, c: a + n * b + n * n * c
    # face point m = i + n*j carries (rquad, squad) = (x[j], x[i])   (src/SetupDG.jl:346)
    for j in range(n):
        for i in range(n):
            m = i + n * j
            for a in range(n):
                Ef[0 * nfp + m, k(a, j, i)] = lm[a]   # f1 r=-1: (s,t)=(x[j],x[i])
                Ef[1 * nfp + m, k(a, j, i)] = lp[a]   # f2 r=+1
                Ef[2 * nfp + m, k(j, a, i)] = lm[a]   # f3 s=-1: (r,t)=(x[j],x[i])
                Ef[3 * nfp + m, k(j, a, i)] = lp[a]   # f4 s=+1
                Ef[4 * nfp + m, k(j, i, a)] = lm[a]   # f5 t=-1: (r,s)=(x[j],x[i])
                Ef[5 * nfp + m, k(j, i, a)] = lp[a]   # f6 t=+1
    rquad, squad = np.repeat(x, n), np.tile(x, n)
    wquad = np.repeat(w, n) * np.tile(w, n)
    e, z = np.ones(nfp), np.zeros(nfp)
    rd = SimpleNamespace(elem="hex", N=N, Nfaces=6, rq=rq, sq=sq, tq=tq, wq=wq, wf=np.tile(wquad, 6),
                         rf=np.concatenate([-e, e, rquad, rquad, rquad, rquad]),
                         sf=np.concatenate([rquad, rquad, -e, e, squad, squad]),
                         tf=np.concatenate([squad, squad, squad, squad, -e, e]),
                         nrJ=np.concatenate([-e, e, z, z, z, z]), nsJ=np.concatenate([z, z, -e, e, z, z]),
                         ntJ=np.concatenate([z, z, z, z, -e, e]))
    W = np.diag(wq)
    Qskew = _hybridize([W @ Dr, W @ Ds, W @ Dt], Ef, rd.wf, [rd.nrJ, rd.nsJ, rd.ntJ])
    Vh = np.vstack([np.eye(n ** 3), Ef])
    ops = SimpleNamespace(elem="hex", N=N, Nfaces=6, wq=wq, Ef=Ef, Qskew=Qskew, Vh=Vh,
                          Ph=2 * np.diag(1 / wq) @ Vh.T,               # hex folds the 2 into Ph (:96)
                          Lf=np.diag(1 / wq) @ (Ef.T @ np.diag(rd.wf)),
                          volume_scale=1.0, lf_scale=lf_scale)
    return rd, ops


def gll_1d(N):
    """Gauss-Lobatto-Legendre points (the interpolation nodes of Basis2DQuad.nodes_2D, :77-81)."""
    if N == 1:
        return np.array([-1.0, 1.0])
    inner = np.polynomial.legendre.Legendre.basis(N).deriv().roots()
    return np.concatenate([[-1.0], np.sort(inner.real), [1.0]])


def reference_quad_modal(N, lf_scale=.25):
    """GLL-nodal quad element with Gauss quadrature and the Vq/Pq projection: ``init_reference_quad``
    (src/SetupDG.jl:205-268) + the operator tuple of the CNS scripts
    (dg2D_CNS_cavity_optimized.jl:61-89), built from 1D Lagrange pieces.  BASELINE config 3."""
    n = N + 1
    xg, wg = gauss_1d(N)
    xl = gll_1d(N)
    D1 = lagrange_deriv(xl)
    Lq = lagrange_eval(xl, xg)                      # Gauss <- GLL
    Lm, Lp = lagrange_eval(xl, -1.0)[0], lagrange_eval(xl, 1.0)[0]
    I = np.eye(n)
    # nodes k = a + n*b: r = xl[a] (fastest), s = xl[b]
    r, s = np.tile(xl, n), np.repeat(xl, n)
    Dr, Ds = np.kron(I, D1), np.kron(D1, I)
    # volume quadrature q = i + n*j: s = xg[i] (fastest), r = xg[j]   (SetupDG.jl:244-246)
    rq, sq = np.repeat(xg, n), np.tile(xg, n)
    wq = np.repeat(wg, n) * np.tile(wg, n)

    def interp(rr, ss):
        A, B = lagrange_eval(xl, rr), lagrange_eval(xl, ss)     # [pt, a], [pt, b]
        return (B[:, :, None] * A[:, None, :]).reshape(len(rr), n * n)   # column k = a + n*b

    Vq = interp(rq, sq)
    e, z = np.ones(n), np.zeros(n)
    rf, sf = np.concatenate([xg, e, -xg, -e]), np.concatenate([-e, xg, e, -xg])   # CCW faces (:235-239)
    wf = np.tile(wg, 4)
    Vf = interp(rf, sf)
    M = Vq.T @ (wq[:, None] * Vq)
    Pq = np.linalg.solve(M, Vq.T * wq[None, :])
    LIFT = np.linalg.solve(M, Vf.T * wf[None, :])
    rd = SimpleNamespace(elem="quad", N=N, Nfaces=4, r=r, s=s, rq=rq, sq=sq, wq=wq, rf=rf, sf=sf, wf=wf,
                         nrJ=np.concatenate([z, e, z, -e]), nsJ=np.concatenate([-e, z, e, z]),
                         Dr=Dr, Ds=Ds, Vq=Vq, Pq=Pq, Vf=Vf, LIFT=LIFT, M=M)
    Ef = Vf @ Pq
    Qskew = _hybridize([Pq.T @ M @ Dr @ Pq, Pq.T @ M @ Ds @ Pq], Ef, wf, [rd.nrJ, rd.nsJ])
    Vh = np.vstack([Vq, Vf])
    ops = SimpleNamespace(elem="quad", N=N, Nfaces=4, wq=wq, Qskew=Qskew, Vh=Vh, Ph=np.linalg.solve(M, Vh.T),
                          VhP=Vh @ Pq, Lf=LIFT, Vq=Vq, Pq=Pq, Vf=Vf, Dr=Dr, Ds=Ds, modal=True,
                          volume_scale=2.0, lf_scale=lf_scale)
    return rd, ops


# ------------------------------------------------------------------ triangles
def _scaled_legendre(i, xi, eta):
    """Q_i(xi, eta) = eta^i P_i(xi / eta) (a polynomial: no division) and its partial derivatives."""
    Q0, Q1 = np.ones_like(xi), xi.copy()
    dx0, dx1 = np.zeros_like(xi), np.ones_like(xi)
    de0, de1 = np.zeros_like(xi), np.zeros_like(xi)
    if i == 0:
        return Q0, dx0, de0
    for k in range(1, i):
        Q2 = ((2 * k + 1) * xi * Q1 - k * eta * eta * Q0) / (k + 1)
        dx2 = ((2 * k + 1) * (Q1 + xi * dx1) - k * eta * eta * dx0) / (k + 1)
        de2 = ((2 * k + 1) * xi * de1 - k * (2 * eta * Q0 + eta * eta * de0)) / (k + 1)
        Q0, Q1, dx0, dx1, de0, de1 = Q1, Q2, dx1, dx2, de1, de2
    return Q1, dx1, de1


def _jacobi_a0(j, alpha, x):
    """P_j^(alpha, 0)(x) (standard normalisation) and its derivative, by the three-term recurrence."""
    P0, P1 = np.ones_like(x), ((alpha + 2) * x + alpha) / 2
    d0, d1 = np.zeros_like(x), np.full_like(x, (alpha + 2) / 2)
    if j == 0:
        return P0, d0
    for n in range(1, j):
        c = 2 * n + alpha
        a1 = 2 * (n + 1) * (n + alpha + 1) * c
        a2 = (c + 1) * alpha * alpha
        a3 = c * (c + 1) * (c + 2)
        a4 = 2 * (n + alpha) * n * (c + 2)
        P2 = ((a2 + a3 * x) * P1 - a4 * P0) / a1
        d2 = (a3 * P1 + (a2 + a3 * x) * d1 - a4 * d0) / a1
        P0, P1, d0, d1 = P1, P2, d1, d2
    return P1, d1


def _tri_basis(N, r, s):
    """A basis of P_N on the reference triangle (collapsed-coordinate products, orthogonal) and its r/s derivatives at
    the points (r, s): columns ordered (i, j), i + j <= N.  Any basis of P_N gives the same nodal operators."""
    r, s = np.asarray(r, dtype=np.float64), np.asarray(s, dtype=np.float64)
    xi, eta = r + (1 + s) / 2, (1 - s) / 2
    V, Vr, Vs = [], [], []
    for i in range(N + 1):
        Q, Qx, Qe = _scaled_legendre(i, xi, eta)
        for j in range(N - i + 1):
            R, dR = _jacobi_a0(j, 2 * i + 1, s)
            V.append(Q * R)
            Vr.append(Qx * R)
            Vs.append((.5 * Qx - .5 * Qe) * R + Q * dR)
    return np.array(V).T, np.array(Vr).T, np.array(Vs).T


def _warp_factor(N, rout):
    """1-D warp of the equispaced points towards the GLL points, evaluated at rout (Hesthaven & Warburton's warp & blend;
    src/Basis2DTri.jl:170-186)."""
    req = np.linspace(-1, 1, N + 1)
    warp = lagrange_eval(req, rout) @ (gll_1d(N) - req)
    zerof = (np.abs(rout) < 1.0 - 1.0e-10).astype(np.float64)
    sf = 1.0 - (zerof * rout) ** 2
    return warp / sf + warp * (zerof - 1)


def tri_nodes(N):
    """Warp & blend interpolation nodes of degree N (src/Basis2DTri.jl:197-233), reference ordering."""
    alpopt = [0.0, 0.0, 1.4152, 0.1001, 0.2751, 0.98, 1.0999, 1.2832, 1.3648, 1.4773, 1.4959, 1.5743, 1.577, 1.6223, 1.6258]
    alpha = alpopt[N - 1] if N < 16 else 5 / 3
    L1, L3 = [], []
    for n in range(N + 1):
        for m in range(N + 1 - n):
            L1.append(n / N)
            L3.append(m / N)
    L1, L3 = np.array(L1), np.array(L3)
    L2 = 1.0 - L1 - L3
    x, y = -L2 + L3, (-L2 - L3 + 2 * L1) / np.sqrt(3.0)
    w1 = 4 * L2 * L3 * _warp_factor(N, L3 - L2) * (1 + (alpha * L1) ** 2)
    w2 = 4 * L1 * L3 * _warp_factor(N, L1 - L3) * (1 + (alpha * L2) ** 2)
    w3 = 4 * L1 * L2 * _warp_factor(N, L2 - L1) * (1 + (alpha * L3) ** 2)
    x = x + w1 + np.cos(2 * np.pi / 3) * w2 + np.cos(4 * np.pi / 3) * w3
    y = y + np.sin(2 * np.pi / 3) * w2 + np.sin(4 * np.pi / 3) * w3
    l1 = (np.sqrt(3.0) * y + 1.0) / 3.0                       # equilateral -> reference right triangle
    l2 = (-3.0 * x - np.sqrt(3.0) * y + 2.0) / 6.0
    l3 = (3.0 * x - np.sqrt(3.0) * y + 2.0) / 6.0
    return -l2 + l3 - l1, -l2 - l3 + l1


def tri_cubature(degree):
    """The reference's triangle cubature of the given (even) degree: src/QuadratureData/quad_nodes_tri_N*.txt, a
    published rule (data, extracted once by oracle/tools/extract_tri_cubature.py; degrees 2..10 shipped)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tri_cubature.json")) as fh:
        tab = json.load(fh)
    if str(degree) not in tab:
        raise ValueError(f"no triangle cubature of degree {degree} in the package (N <= 5)")
    rsw = np.array(tab[str(degree)], dtype=np.float64)
    return rsw[:, 0].copy(), rsw[:, 1].copy(), rsw[:, 2].copy()


def reference_tri(N, lf_scale=.25):
    """Modal triangle element of the Euler / CNS scripts: ``init_reference_tri`` (src/SetupDG.jl:151-201) + the operator
    tuple of dg2D_euler_tri.jl:45-63 / dg2D_CNS_cavity_optimized.jl:61-89.  BASELINE config 2."""
    r, s = tri_nodes(N)
    V, Vr, Vs = _tri_basis(N, r, s)
    Vinv = np.linalg.inv(V)
    Dr, Ds = Vr @ Vinv, Vs @ Vinv
    x1, w1 = gauss_1d(N)
    e, z = np.ones_like(x1), np.zeros_like(x1)
    rf, sf = np.concatenate([x1, -x1, -e]), np.concatenate([-e, x1, -x1])       # faces s=-1, r+s=0, r=-1 (:176-181)
    wf = np.tile(w1, 3)
    nrJ, nsJ = np.concatenate([z, e, -e]), np.concatenate([-e, e, z])
    rq, sq, wq = tri_cubature(2 * N)
    Vq = _tri_basis(N, rq, sq)[0] @ Vinv
    Vf = _tri_basis(N, rf, sf)[0] @ Vinv
    M = Vq.T @ (wq[:, None] * Vq)
    Pq = np.linalg.solve(M, Vq.T * wq[None, :])
    LIFT = np.linalg.solve(M, Vf.T * wf[None, :])
    rd = SimpleNamespace(elem="tri", N=N, Nfaces=3, r=r, s=s, rq=rq, sq=sq, wq=wq, rf=rf, sf=sf, wf=wf, nrJ=nrJ, nsJ=nsJ,
                         Dr=Dr, Ds=Ds, Vq=Vq, Pq=Pq, Vf=Vf, LIFT=LIFT, M=M)
    Ef = Vf @ Pq
    Qskew = _hybridize([Pq.T @ M @ Dr @ Pq, Pq.T @ M @ Ds @ Pq], Ef, wf, [nrJ, nsJ])
    Vh = np.vstack([Vq, Vf])
    ops = SimpleNamespace(elem="tri", N=N, Nfaces=3, wq=wq, Ef=Ef, Qskew=Qskew, Vh=Vh, Ph=np.linalg.solve(M, Vh.T),
                          VhP=Vh @ Pq, Lf=LIFT, Vq=Vq, Pq=Pq, Vf=Vf, Dr=Dr, Ds=Ds, modal=True,
                          volume_scale=2.0, lf_scale=lf_scale)
    return rd, ops


def uniform_tri_mesh_md(rd, Kx, Ky, domain=((-1.0, 1.0), (-1.0, 1.0)), periodic=(True, True)):
    """MeshData of the uniform triangle mesh of src/UniformTriMesh.jl:25-42: every cell (i, j), x-fastest, is split along
    its SW-NE diagonal into element 2c = (SW, SE, NE) and 2c + 1 = (NE, NW, SW).  Connectivity is analytic: a face always
    meets the SAME face number of its neighbour (bottom/top edges are faces 1, right/left edges faces 2, the diagonal faces
    3) with the point order reversed."""
    n = rd.N + 1
    Nfq = 3 * n
    (x0, x1), (y0, y1) = domain
    hx, hy = (x1 - x0) / Kx, (y1 - y0) / Ky
    K = 2 * Kx * Ky
    c = np.arange(Kx * Ky)
    ci, cj = c % Kx, c // Kx
    X0, Y0 = x0 + hx * ci, y0 + hy * cj
    VX = np.empty((3, K))
    VY = np.empty((3, K))
    VX[:, 0::2] = [X0, X0 + hx, X0 + hx]
    VY[:, 0::2] = [Y0, Y0, Y0 + hy]
    VX[:, 1::2] = [X0 + hx, X0, X0]
    VY[:, 1::2] = [Y0 + hy, Y0 + hy, Y0]

    def affine(rr, ss):                        # vertex interpolation: (-(r+s)/2, (1+r)/2, (1+s)/2)
        W = np.stack([-(rr + ss) / 2, (1 + rr) / 2, (1 + ss) / 2], axis=1)
        return W @ VX, W @ VY

    md = SimpleNamespace(K=K, dim=2)
    md.x, md.y = affine(rd.r, rd.s)
    md.xq, md.yq = affine(rd.rq, rd.sq)
    md.xf, md.yf = affine(rd.rf, rd.sf)
    xr, xs = (VX[1] - VX[0]) / 2, (VX[2] - VX[0]) / 2
    yr, ys = (VY[1] - VY[0]) / 2, (VY[2] - VY[0]) / 2
    one = np.ones((1, K))
    md.rxJ, md.sxJ, md.ryJ, md.syJ = ys * one, -yr * one, -xs * one, xr * one
    md.J = (xr * ys - xs * yr) * one
    nr, ns = rd.nrJ[::n][:, None], rd.nsJ[::n][:, None]       # one row per face
    md.nxJ = nr * md.rxJ + ns * md.sxJ
    md.nyJ = nr * md.ryJ + ns * md.syJ
    md.sJ = np.sqrt(md.nxJ ** 2 + md.nyJ ** 2)
    q = np.arange(n)
    mapP = np.empty((Nfq, K), dtype=np.int64)
    self_idx = np.arange(Nfq)[:, None] + Nfq * np.arange(K)[None, :]
    e0, e1 = 2 * c, 2 * c + 1

    def link(face, ea, eb, ok):                # face `face` of elements ea <-> the same face of eb, points reversed
        rows = face * n + q
        tgt = (face * n + (n - 1 - q))[:, None] + Nfq * eb[None, :]
        mapP[np.ix_(rows, ea)] = np.where(ok[None, :], tgt, self_idx[np.ix_(rows, ea)])

    below = ci + Kx * ((cj - 1) % Ky)
    above = ci + Kx * ((cj + 1) % Ky)
    right = (ci + 1) % Kx + Kx * cj
    left = (ci - 1) % Kx + Kx * cj
    yes = np.ones(Kx * Ky, dtype=bool)
    link(0, e0, 2 * below + 1, yes if periodic[1] else cj > 0)              # my bottom edge = top edge of the cell below
    link(0, e1, 2 * above, yes if periodic[1] else cj < Ky - 1)
    link(1, e0, 2 * right + 1, yes if periodic[0] else ci < Kx - 1)         # my right edge = left edge of the cell to the right
    link(1, e1, 2 * left, yes if periodic[0] else ci > 0)
    link(2, e0, e1, yes)                                                    # the diagonal
    link(2, e1, e0, yes)
    md.mapP = mapP
    md.n_halo = 0
    return md


def cavity_bc_sets(md):
    """``lid`` / ``wall`` node sets of the lid-driven cavity on [-1,1]^2
    (dg2D_CNS_cavity_optimized.jl:139-147); 0-based linear indices into the Nfq x K face arrays."""
    mp = md.mapP.reshape(-1, order="F")
    mapB = np.nonzero(mp == np.arange(mp.size))[0]
    yb = md.yf.reshape(-1, order="F")[mapB]
    lid = mapB[np.abs(yb - 1) < 1e-12]
    wall = mapB[np.abs(yb - 1) >= 1e-12]
    return SimpleNamespace(lid=lid, wall=wall, vlid=np.ones(lid.size), mapB=mapB)


# --------------------------------------------------------------------- meshes
def uniform_quad_mesh_md(rd, Kx, Ky, domain=((-1.0, 1.0), (-1.0, 1.0)), periodic=(True, True), yslab=None):
    """MeshData of a uniform Kx x Ky quad mesh (elements x-fastest, UniformQuadMesh.jl:37-44).

    ``yslab=(rank, nranks)`` builds only the rank's contiguous range of element rows: trace
    points owned by the slab below / above are mapped to halo slots appended after ``Nfq*K`` (lower
    neighbour's s=+1 faces first, then the upper neighbour's s=-1 faces); ``md.send_lower`` /
    ``md.send_upper`` list the local trace points each neighbour needs, in the order of its halo slots.
    """
    n = rd.N + 1
    Nfq = 4 * n
    (x0, x1), (y0, y1) = domain
    hx, hy = (x1 - x0) / Kx, (y1 - y0) / Ky
    if yslab is None:
        j0, j1, nranks = 0, Ky, 1
    else:
        rank, nranks = yslab
        if Ky % nranks:
            raise ValueError("Ky must be divisible by the number of ranks")
        j0, j1 = rank * (Ky // nranks), (rank + 1) * (Ky // nranks)
    split = yslab is not None and nranks > 1
    K = Kx * (j1 - j0)
    ex = np.arange(K) % Kx
    ey = np.arange(K) // Kx + j0
    md = SimpleNamespace(K=K, dim=2, Kxy=(Kx, Ky), yrange=(j0, j1))
    md.xq = (x0 + hx * ex)[None, :] + hx / 2 * (1 + rd.rq)[:, None]
    md.yq = (y0 + hy * ey)[None, :] + hy / 2 * (1 + rd.sq)[:, None]
    md.xf = (x0 + hx * ex)[None, :] + hx / 2 * (1 + rd.rf)[:, None]
    md.yf = (y0 + hy * ey)[None, :] + hy / 2 * (1 + rd.sf)[:, None]
    if hasattr(rd, "r"):                     # interpolation nodes of a modal/nodal element
        md.x = (x0 + hx * ex)[None, :] + hx / 2 * (1 + rd.r)[:, None]
        md.y = (y0 + hy * ey)[None, :] + hy / 2 * (1 + rd.s)[:, None]
    one = np.ones((1, K))
    md.rxJ, md.sxJ, md.ryJ, md.syJ = hy / 2 * one, 0 * one, 0 * one, hx / 2 * one
    md.J = hx * hy / 4 * one
    nr = rd.nrJ[::n][:, None]
    ns = rd.nsJ[::n][:, None]
    md.nxJ = md.rxJ * nr + md.sxJ * ns          # Nfaces x K (constant per face)
    md.nyJ = md.ryJ * nr + md.syJ * ns
    md.sJ = np.sqrt(md.nxJ ** 2 + md.nyJ ** 2)
    # connectivity: f1(s=-1)<->f3 of (ex,ey-1); f2(r=+1)<->f4 of (ex+1,ey); points reversed
    mapP = np.arange(Nfq * K, dtype=np.int64).reshape((Nfq, K), order="F")
    m = np.arange(n)
    for f, (dx, dy, fn) in enumerate([(0, -1, 2), (1, 0, 3), (0, 1, 0), (-1, 0, 1)]):
        nx_, ny_ = ex + dx, ey + dy
        ok = np.ones(K, dtype=bool)
        if periodic[0]:
            nx_ = nx_ % Kx
        else:
            ok &= (nx_ >= 0) & (nx_ < Kx)
        if periodic[1]:
            ny_ = ny_ % Ky
        else:
            ok &= (ny_ >= 0) & (ny_ < Ky)
        local = (ny_ >= j0) & (ny_ < j1)
        en = np.where(ok & local, nx_ + Kx * (ny_ - j0), 0)
        tgt = (fn * n + (n - 1 - m))[:, None] + Nfq * en[None, :]
        if split and dy != 0:
            # the neighbour row lives on another rank: slot = its elements in x order, its face points in order
            slot = nx_[None, :] * n + (n - 1 - m)[:, None]
            tgt = np.where(local[None, :], tgt, Nfq * K + slot + (0 if dy < 0 else Kx * n))
        rows = f * n + m
        mapP[rows[:, None], np.arange(K)[None, :]] = np.where(ok[None, :], tgt, mapP[rows, :])
    md.mapP = mapP
    md.n_halo = 0
    if split:
        md.n_halo = 2 * Kx * n
        bottom = np.nonzero(ey == j0)[0]
        top = np.nonzero(ey == j1 - 1)[0]
        md.send_lower = ((0 * n + m)[None, :] + Nfq * bottom[:, None]).reshape(-1)   # my s=-1 faces (face 1)
        md.send_upper = ((2 * n + m)[None, :] + Nfq * top[:, None]).reshape(-1)      # my s=+1 faces (face 3)
        has_lower = periodic[1] or j0 > 0
        has_upper = periodic[1] or j1 < Ky
        md.lower_rank = (rank - 1) % nranks if has_lower else None
        md.upper_rank = (rank + 1) % nranks if has_upper else None
    return md


def uniform_hex_mesh_md(rd, Kxyz, domain=((-1.0, 1.0),) * 3, periodic=(True, True, True), zslab=None):
    """MeshData of a uniform hex mesh, elements x-fastest (UniformHexMesh.jl:59-63).

    ``zslab=(rank, nranks)`` builds only the rank's contiguous z-slab: trace
    points owned by the slab below / above are mapped to halo slots appended after
    ``Nfq*K`` (lower neighbour's t=+1 faces first, then the upper neighbour's t=-1 faces);
    ``md.send_lower`` / ``md.send_upper`` list the local trace points each neighbour needs.
    """
    n = rd.N + 1
    nfp = n * n
    Nfq = 6 * nfp
    Kx, Ky, Kz = Kxyz
    (x0, x1), (y0, y1), (z0, z1) = domain
    hx, hy, hz = (x1 - x0) / Kx, (y1 - y0) / Ky, (z1 - z0) / Kz
    if zslab is None:
        k0, k1, nranks = 0, Kz, 1
    else:
        rank, nranks = zslab
        if Kz % nranks:
            raise ValueError("Kz must be divisible by the number of ranks")
        k0, k1 = rank * (Kz // nranks), (rank + 1) * (Kz // nranks)
    Kzl = k1 - k0
    K = Kx * Ky * Kzl
    e = np.arange(K)
    ex, ey, ez = e % Kx, (e // Kx) % Ky, e // (Kx * Ky) + k0
    md = SimpleNamespace(K=K, dim=3, Kxyz=(Kx, Ky, Kz), zrange=(k0, k1))
    md.elem_origin = (x0 + hx * ex, y0 + hy * ey, z0 + hz * ez)
    md.h = (hx, hy, hz)
    one = np.ones((1, K))
    zero = 0 * one
    md.rxJ, md.syJ, md.tzJ = hy * hz / 4 * one, hx * hz / 4 * one, hx * hy / 4 * one
    md.sxJ = md.txJ = md.ryJ = md.tyJ = md.rzJ = md.szJ = zero
    md.J = hx * hy * hz / 8 * one
    nr, ns, nt = rd.nrJ[::nfp][:, None], rd.nsJ[::nfp][:, None], rd.ntJ[::nfp][:, None]
    md.nxJ = nr * md.rxJ + ns * md.sxJ + nt * md.txJ
    md.nyJ = nr * md.ryJ + ns * md.syJ + nt * md.tyJ
    md.nzJ = nr * md.rzJ + ns * md.szJ + nt * md.tzJ
    md.sJ = np.sqrt(md.nxJ ** 2 + md.nyJ ** 2 + md.nzJ ** 2)
    # connectivity: opposite faces share the on-face parametrisation, so point m <-> point m
    mapP = np.empty((Nfq, K), dtype=np.int64)
    m = np.arange(nfp)
    split = zslab is not None and (nranks > 1)
    n_layer = Kx * Ky * nfp
    for f, (d, sgn, fn) in enumerate([(0, -1, 1), (0, 1, 0), (1, -1, 3), (1, 1, 2), (2, -1, 5), (2, 1, 4)]):
        c = [ex.copy(), ey.copy(), ez.copy()]
        c[d] = c[d] + sgn
        ok = np.ones(K, dtype=bool)
        for q, Kq in enumerate((Kx, Ky, Kz)):
            if periodic[q]:
                c[q] = c[q] % Kq
            else:
                ok &= (c[q] >= 0) & (c[q] < Kq)
        self_idx = (f * nfp + m)[:, None] + Nfq * e[None, :]
        local = (c[2] >= k0) & (c[2] < k1) if split else np.ones(K, dtype=bool)
        en = c[0] + Kx * c[1] + Kx * Ky * (c[2] - k0)
        tgt = (fn * nfp + m)[:, None] + Nfq * np.where(local, en, 0)[None, :]
        if split and d == 2:
            slot = (c[0] + Kx * c[1])[None, :] * nfp + m[:, None]
            halo = Nfq * K + slot + (0 if sgn < 0 else n_layer)
            tgt = np.where(local[None, :], tgt, halo)
        mapP[f * nfp:(f + 1) * nfp, :] = np.where(ok[None, :], tgt, self_idx)
    md.mapP = mapP
    md.n_halo = 0
    if split:
        md.n_halo = 2 * n_layer
        bottom = np.nonzero(ez == k0)[0]
        top = np.nonzero(ez == k1 - 1)[0]
        md.send_lower = ((4 * nfp + m)[None, :] + Nfq * bottom[:, None]).reshape(-1)   # my t=-1 faces
        md.send_upper = ((5 * nfp + m)[None, :] + Nfq * top[:, None]).reshape(-1)      # my t=+1 faces
        has_lower = periodic[2] or k0 > 0
        has_upper = periodic[2] or k1 < Kz
        md.lower_rank = (rank - 1) % nranks if has_lower else None
        md.upper_rank = (rank + 1) % nranks if has_upper else None
    return md


def curved_hex_mesh_md(rd, Kxyz, domain=((-1.0, 1.0),) * 3, amp=.08):
    """A periodic hex mesh with CURVED (trilinear, non-affine) elements: the vertices of the uniform mesh are displaced by a
    smooth periodic field of amplitude ``amp`` element widths, so every element is a general hexahedron, the mesh stays
    watertight (a trilinear map is reproduced exactly by any N >= 1) and the connectivity is that of the uniform mesh.
    Geometric factors in the curl form of src/geometric_factors.jl:34-67 at the Gauss nodes, interpolated to the face
    points as the hex script does (dg3D_euler_hex.jl:69-101): ``rxJ.. : Nh x K``, ``J : Nq x K``, ``nxJ.., sJ : Nfq x K``."""
    md = uniform_hex_mesh_md(rd, Kxyz, domain=domain)
    n = rd.N + 1
    Kx, Ky, Kz = Kxyz
    (x0, x1), (y0, y1), (z0, z1) = domain
    hx, hy, hz = md.h
    x1d, _ = gauss_1d(rd.N)
    D = lagrange_deriv(x1d)
    I = np.eye(n)
    Dr, Ds, Dt = np.kron(I, np.kron(I, D)), np.kron(I, np.kron(D, I)), np.kron(D, np.kron(I, I))
    lm, lp = lagrange_eval(x1d, -1.0)[0], lagrange_eval(x1d, 1.0)[0]
    e = np.arange(md.K)
    ex, ey, ez = e % Kx, (e // Kx) % Ky, e // (Kx * Ky)

    def vertex(ix, iy, iz):                       # displaced vertex (ix, iy, iz) of the lattice, periodic in the displacement
        X, Y, Z = x0 + hx * ix, y0 + hy * iy, z0 + hz * iz
        a, b, c = 2 * np.pi * ix / Kx, 2 * np.pi * iy / Ky, 2 * np.pi * iz / Kz
        return (X + amp * hx * np.sin(b + .3) * np.cos(c) * np.cos(a),
                Y + amp * hy * np.sin(c - .2) * np.cos(a) * np.cos(b),
                Z + amp * hz * np.sin(a + .5) * np.cos(b) * np.cos(c))

    xyz = [np.zeros((n ** 3, md.K)) for _ in range(3)]
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                shape = ((1 + (2 * dx - 1) * rd.rq) * (1 + (2 * dy - 1) * rd.sq) * (1 + (2 * dz - 1) * rd.tq) / 8)[:, None]
                V = vertex(ex + dx, ey + dy, ez + dz)
                for q in range(3):
                    xyz[q] += shape * V[q][None, :]
    x, y, z = xyz
    xr, xs, xt = Dr @ x, Ds @ x, Dt @ x
    yr, ys, yt = Dr @ y, Ds @ y, Dt @ y
    zr, zs, zt = Dr @ z, Ds @ z, Dt @ z
    Fr, Fs, Ft = yr * z, ys * z, yt * z
    G = {}
    G["rxJ"], G["sxJ"], G["txJ"] = Dt @ Fs - Ds @ Ft, Dr @ Ft - Dt @ Fr, Ds @ Fr - Dr @ Fs
    Fr, Fs, Ft = xr * z, xs * z, xt * z
    G["ryJ"], G["syJ"], G["tyJ"] = -(Dt @ Fs - Ds @ Ft), -(Dr @ Ft - Dt @ Fr), -(Ds @ Fr - Dr @ Fs)
    Fr, Fs, Ft = yr * x, ys * x, yt * x
    G["rzJ"], G["szJ"], G["tzJ"] = -(Dt @ Fs - Ds @ Ft), -(Dr @ Ft - Dt @ Fr), -(Ds @ Fr - Dr @ Fs)
    md.J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)
    _, ops = reference_hex(rd.N)
    Ef = ops.Ef
    for k, v in G.items():
        setattr(md, k, np.vstack([v, Ef @ v]))
    Nq = n ** 3
    nr, ns, nt = rd.nrJ[:, None], rd.nsJ[:, None], rd.ntJ[:, None]
    md.nxJ = nr * md.rxJ[Nq:] + ns * md.sxJ[Nq:] + nt * md.txJ[Nq:]
    md.nyJ = nr * md.ryJ[Nq:] + ns * md.syJ[Nq:] + nt * md.tyJ[Nq:]
    md.nzJ = nr * md.rzJ[Nq:] + ns * md.szJ[Nq:] + nt * md.tzJ[Nq:]
    md.sJ = np.sqrt(md.nxJ ** 2 + md.nyJ ** 2 + md.nzJ ** 2)
    md.xq, md.yq, md.zq = x, y, z
    md.curved = True
    return md


def structured_hex_md(rd, Kxyz, domain=((-1.0, 1.0),) * 3, periodic=(True, True, True), zslab=None):
    """The same mesh as ``uniform_hex_mesh_md`` WITHOUT its arrays: ``api.Solver`` hands the six integers and the box to
    ``esdg_set_structured_hex`` and the library generates mapP and the geometric factors on the device (SURVEY 8f
    rank 3).  Send lists and halo size of a z-slab are the same as there."""
    n = rd.N + 1
    nfp = n * n
    Nfq = 6 * nfp
    Kx, Ky, Kz = Kxyz
    (x0, x1), (y0, y1), (z0, z1) = domain
    hx, hy, hz = (x1 - x0) / Kx, (y1 - y0) / Ky, (z1 - z0) / Kz
    if zslab is None:
        rank, nranks = 0, 1
    else:
        rank, nranks = zslab
        if Kz % nranks:
            raise ValueError("Kz must be divisible by the number of ranks")
    k0, k1 = rank * (Kz // nranks), (rank + 1) * (Kz // nranks)
    K = Kx * Ky * (k1 - k0)
    e = np.arange(K)
    ex, ey, ez = e % Kx, (e // Kx) % Ky, e // (Kx * Ky) + k0
    md = SimpleNamespace(K=K, dim=3, Kxyz=(Kx, Ky, Kz), zrange=(k0, k1), structured=True, domain=domain, periodic=periodic,
                         elem_origin=(x0 + hx * ex, y0 + hy * ey, z0 + hz * ez), h=(hx, hy, hz), n_halo=0)
    if nranks > 1:
        m = np.arange(nfp)
        md.n_halo = 2 * Kx * Ky * nfp
        bottom = np.nonzero(ez == k0)[0]
        top = np.nonzero(ez == k1 - 1)[0]
        md.send_lower = ((4 * nfp + m)[None, :] + Nfq * bottom[:, None]).reshape(-1)
        md.send_upper = ((5 * nfp + m)[None, :] + Nfq * top[:, None]).reshape(-1)
        md.lower_rank = (rank - 1) % nranks if (periodic[2] or k0 > 0) else None
        md.upper_rank = (rank + 1) % nranks if (periodic[2] or k1 < Kz) else None
    return md


def hex_node_coords(rd, md):
    """Physical coordinates of the volume quadrature nodes (Nq x K each)."""
    ox, oy, oz = md.elem_origin
    hx, hy, hz = md.h
    return (ox[None, :] + hx / 2 * (1 + rd.rq)[:, None],
            oy[None, :] + hy / 2 * (1 + rd.sq)[:, None],
            oz[None, :] + hz / 2 * (1 + rd.tq)[:, None])


# --------------------------------------------------------- initial conditions
GAMMA = 1.4


def primitive_to_conservative(rho, U, p):
    """examples/EntropyStableEuler/euler_variables.jl:15-20"""
    un = sum(u * u for u in U)
    return [rho] + [rho * u for u in U] + [p / (GAMMA - 1) + .5 * rho * un]


def vortex(x, y, t=0.0):
    """2D isentropic vortex, examples/EntropyStableEuler/EntropyStableEuler.jl:21-35"""
    r2 = (x - 5 - t) ** 2 + y ** 2
    b = 5 * np.exp(1 - r2)
    u = 1 - b * y / (2 * np.pi)
    v = b * (x - 5 - t) / (2 * np.pi)
    rho = (1 - (1 / (8 * GAMMA * np.pi ** 2)) * (GAMMA - 1) / 2 * b ** 2) ** (1 / (GAMMA - 1))
    return rho, u, v, rho ** GAMMA


def taylor_green(x, y, z, mach=0.1):
    """3D Taylor-Green vortex (project-defined IC for config 5)."""
    rho = np.ones_like(x)
    u = np.sin(x) * np.cos(y) * np.cos(z)
    v = -np.cos(x) * np.sin(y) * np.cos(z)
    w = np.zeros_like(x)
    p = 1.0 / (GAMMA * mach ** 2) + (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2) / 16
    return rho, u, v, w, p
