"""Hybridized-SBP operator assembly, ``rhs`` variants and LSRK45 (oracle).

Restates the script-level hot path of the reference:

* quad  : ``examples/dg2D_euler_quad.jl``  ops :47-78, hadamard :102-138, rhs :141-194, LSRK :196-207
* tri   : ``examples/dg2D_euler_tri.jl``   ops :45-63,77, hadamard :88-126, rhs :130-186
* hex   : ``examples/dg3D_euler_hex.jl``   ops :34-58,94-98, hadamard :122-164, rhs :167-222
  (with the H1/H2 ordering fix; LF factor is a parameter because the
  reference multiplies it by 0, :193)
* LSRK45 coefficients: ``src/CommonUtils.jl:29-49``

The Hadamard sums are vectorised over elements (the reference loops ``for e =
1:K``); the (i, j) loops and their accumulation order are the reference's.
"""
from types import SimpleNamespace

import numpy as np

from . import physics as P
from . import setupdg as S
from . import mesh as M


# ------------------------------------------------------------------ LSRK45
def rk45_coeffs():  # CommonUtils.jl:29-49
    rk4a = np.array([0.0,
                     -567301805773.0 / 1357537059087.0,
                     -2404267990393.0 / 2016746695238.0,
                     -3550918686646.0 / 2091501179385.0,
                     -1275806237668.0 / 842570457699.0])
    rk4b = np.array([1432997174477.0 / 9575080441755.0,
                     5161836677717.0 / 13612068292357.0,
                     1720146321549.0 / 2090206949498.0,
                     3134564353537.0 / 4481467310338.0,
                     2277821191437.0 / 14882151754819.0])
    rk4c = np.array([0.0,
                     1432997174477.0 / 9575080441755.0,
                     2526269341429.0 / 6820363962896.0,
                     2006345519317.0 / 3224310063776.0,
                     2802321613138.0 / 2924317926251.0,
                     1.0])
    return rk4a, rk4b, rk4c


# ------------------------------------------------------- operator assembly
def _droptol(A, tol):
    A = np.array(A, copy=True)
    A[np.abs(A) <= tol] = 0.0
    return A


def _union_ids(mats):
    """``unique([A[i,:].nzind; B[i,:].nzind; ...])`` (first-seen order)."""
    n = mats[0].shape[0]
    out = []
    for i in range(n):
        seen = []
        for A in mats:
            for j in np.nonzero(A[i, :])[0]:
                if j not in seen:
                    seen.append(int(j))
        out.append(seen)
    return out


def build_sbp_ops(rd):
    """Hybridized skew SBP operators + quadrature-basis Vh/Ph/Lf.

    quad: dg2D_euler_quad.jl:47-78 ; hex: dg3D_euler_hex.jl:34-58,94-98 ;
    tri: dg2D_euler_tri.jl:45-63.
    """
    ops = SimpleNamespace()
    ops.elem, ops.N, ops.Nfaces, ops.wq = rd.elem, rd.N, rd.Nfaces, rd.wq
    ops.lf_scale = {"quad": .5, "tri": .25, "hex": .25}[rd.elem]   # quad :165, tri :156, (hex :193 has 0*.25)
    dim = 3 if rd.elem == "hex" else 2
    Ds = [rd.Dr, rd.Ds] + ([rd.Dt] if dim == 3 else [])
    ns = [rd.nrJ, rd.nsJ] + ([rd.ntJ] if dim == 3 else [])
    Ef = rd.Vf @ rd.Pq
    ops.Ef = Ef
    Qh = []
    for D, n in zip(Ds, ns):
        Q = rd.Pq.T @ rd.M @ D @ rd.Pq
        Bm = np.diag(rd.wf * n)
        Qh.append(.5 * np.block([[Q - Q.T, Ef.T @ Bm], [-Bm @ Ef, Bm]]))
    skew = [.5 * (A - A.T) for A in Qh]
    ops.Qskew = skew                       # Qrhskew, Qshskew[, Qthskew]
    Nq = rd.wq.size
    if rd.elem in ("quad", "hex"):
        ops.Qsparse = [_droptol(A, 1e-12) for A in skew]
        ops.Qnzids = _union_ids(ops.Qsparse)
        Vh = _droptol(np.vstack([np.eye(Nq), Ef]), 1e-12)
        two = 2.0 if rd.elem == "hex" else 1.0          # hex folds the 2 into Ph (dg3D_euler_hex.jl:96)
        ops.Vh = Vh
        ops.Ph = _droptol(two * np.diag(1 / rd.wq) @ Vh.T, 1e-12)
        ops.Lf = _droptol(np.diag(1 / rd.wq) @ (Ef.T @ np.diag(rd.wf)), 1e-12)
        ops.volume_scale = 1.0 if rd.elem == "hex" else 2.0
    else:
        Vh = np.vstack([rd.Vq, rd.Vf])
        ops.Vh = Vh
        ops.Ph = np.linalg.solve(rd.M, Vh.T)
        ops.VhP = Vh @ rd.Pq
        ops.Lf = rd.LIFT
        ops.Vq = rd.Vq
        ops.volume_scale = 2.0
    return ops


def interp_geofacs(md, rd, ops):
    """``rxJ.. = Vh*rxJ..`` (quad :86-88, tri :71-73) / ``[Vq;Vf]*..`` (hex :90-91), J=Vq*J (hex :99-101)."""
    if rd.elem == "hex":
        VqVf = np.vstack([rd.Vq, rd.Vf])
        for name in ("rxJ", "sxJ", "txJ", "ryJ", "syJ", "tyJ", "rzJ", "szJ", "tzJ"):
            setattr(md, name, VqVf @ getattr(md, name))
        md.J = rd.Vq @ md.J
        md.wJq = np.diag(rd.wq) @ md.J
    else:
        for name in ("rxJ", "sxJ", "ryJ", "syJ"):
            setattr(md, name, ops.Vh @ getattr(md, name))
    return md


# --------------------------------------------------------------- Hadamard
def sparse_hadamard_sum(Qh, ops, vgeo, dim):
    """Vectorised-over-elements ``sparse_hadamard_sum``.

    2D (dg2D_euler_quad.jl:102-138): vgeo = (rxJ,sxJ,ryJ,syJ) rows ``[1,e]`` (affine).
    3D (dg3D_euler_hex.jl:122-164): vgeo = 9 arrays Nh x K, averaged per pair.
    """
    nf = len(Qh)
    Nh, K = Qh[0].shape
    Qlog = (P._log(Qh[0]), P._log(Qh[-1]))
    dt = Qh[0].dtype
    out = [np.zeros((Nh, K), dtype=dt) for _ in range(nf)]
    Qs = ops.Qsparse
    for i in range(Nh):
        Qi = tuple(q[i] for q in Qh)
        Li = tuple(q[i] for q in Qlog)
        acc = [np.zeros(K, dtype=dt) for _ in range(nf)]
        for j in ops.Qnzids[i]:
            Qj = tuple(q[j] for q in Qh)
            Lj = tuple(q[j] for q in Qlog)
            F = P.euler_fluxes(Qi, Qj, Li, Lj)
            if dim == 2:
                rxJ, sxJ, ryJ, syJ = vgeo
                Fx, Fy = F
                for f in range(nf):
                    Fr = rxJ * Fx[f] + ryJ * Fy[f]
                    Fs = sxJ * Fx[f] + syJ * Fy[f]
                    acc[f] = acc[f] + (Qs[0][i, j] * Fr + Qs[1][i, j] * Fs)
            else:
                g = [.5 * (a[i] + a[j]) for a in vgeo]
                rxJ, sxJ, txJ, ryJ, syJ, tyJ, rzJ, szJ, tzJ = g
                Fx, Fy, Fz = F
                for f in range(nf):
                    Fr = rxJ * Fx[f] + ryJ * Fy[f] + rzJ * Fz[f]
                    Fs = sxJ * Fx[f] + syJ * Fy[f] + szJ * Fz[f]
                    Ft = txJ * Fx[f] + tyJ * Fy[f] + tzJ * Fz[f]
                    acc[f] = acc[f] + (Qs[0][i, j] * Fr + Qs[1][i, j] * Fs + Qs[2][i, j] * Ft)
        for f in range(nf):
            out[f][i] = acc[f]
    return out


def dense_hadamard_sum(Qh, ops, vgeo):
    """``dense_hadamard_sum`` of dg2D_euler_tri.jl:88-126, vectorised over elements."""
    rxJ, sxJ, ryJ, syJ = vgeo                      # each (K,)
    Qr, Qs = ops.Qskew
    nf = len(Qh)
    Nh, K = Qh[0].shape
    Qlog = (P._log(Qh[0]), P._log(Qh[-1]))
    out = [np.zeros((Nh, K)) for _ in range(nf)]
    for i in range(Nh):
        Qi = tuple(q[i] for q in Qh)
        Li = tuple(q[i] for q in Qlog)
        acc = [np.zeros(K) for _ in range(nf)]
        for j in range(Nh):
            Qj = tuple(q[j] for q in Qh)
            Lj = tuple(q[j] for q in Qlog)
            Fx, Fy = P.euler_fluxes_2d(Qi, Qj, Li, Lj)
            Qx = rxJ * Qr[i, j] + sxJ * Qs[i, j]   # QxTr[j,i]
            Qy = ryJ * Qr[i, j] + syJ * Qs[i, j]
            for f in range(nf):
                acc[f] = acc[f] + (Qx * Fx[f] + Qy * Fy[f])
        for f in range(nf):
            out[f][i] = acc[f]
    return out


# -------------------------------------------------------------------- rhs
def _gather(A, mapP):
    return A.flatten(order="F")[mapP]


def rhs_collocated(Q, md, ops, rd, lf_scale, compute_rhstest=False):
    """``rhs`` for quad (dg2D_euler_quad.jl:141-194) and hex (dg3D_euler_hex.jl:167-222).

    ``lf_scale`` is the constant in ``LFc = lf_scale*max(lam,lamP)*sJ``:
    .5 in the quad script (:165); the hex script has ``0*.25`` (:193).
    """
    dim = 3 if rd.elem == "hex" else 2
    nf = len(Q)
    Nq = rd.wq.size
    VU = P.v_of_u(Q)
    Uf = P.u_of_v(tuple(ops.Ef @ v for v in VU))
    Uh = tuple(np.vstack([q, uf]) for q, uf in zip(Q, Uf))
    beta = P.beta_of_u(Uh)
    Qh = (Uh[0],) + tuple(Uh[k] / Uh[0] for k in range(1, nf - 1)) + (beta,)
    QM = tuple(q[Nq:, :] for q in Qh)
    QP = tuple(_gather(q, md.mapP) for q in QM)
    normals = (md.nxJ, md.nyJ) + ((md.nzJ,) if dim == 3 else ())
    rhoUn = sum(Uf[1 + d] * normals[d] for d in range(dim)) / md.sJ
    lam = np.abs(P.wavespeed(Uf[0], rhoUn, Uf[-1]))
    LFc = lf_scale * np.maximum(lam, _gather(lam, md.mapP)) * md.sJ
    fS = P.euler_fluxes(QM, QP)
    rhsQ = []
    for f in range(nf):
        flux = sum(fS[d][f] * normals[d] for d in range(dim)) - LFc * (_gather(Uf[f], md.mapP) - Uf[f])
        rhsQ.append(ops.Lf @ flux)
    if dim == 2:
        vgeo = tuple(a[0, :] for a in (md.rxJ, md.sxJ, md.ryJ, md.syJ))
    else:
        vgeo = (md.rxJ, md.sxJ, md.txJ, md.ryJ, md.syJ, md.tyJ, md.rzJ, md.szJ, md.tzJ)
    QF = sparse_hadamard_sum(Qh, ops, vgeo, dim)
    for f in range(nf):
        rhsQ[f] = rhsQ[f] + ops.volume_scale * (ops.Ph @ QF[f])
    rhsQ = [-x / md.J for x in rhsQ]
    rhstest = 0.0
    if compute_rhstest:
        for f in range(nf):
            rhstest += np.sum(md.wJq * VU[f] * rhsQ[f])
    return rhsQ, rhstest


def rhs_tri(Q, md, ops, rd, lf_scale=.25, compute_rhstest=False):
    """``rhs`` of dg2D_euler_tri.jl:130-186 (modal, Vq/Pq projection, dense ops)."""
    nf = len(Q)
    Nh, Nq = ops.VhP.shape
    VU = P.v_of_u(tuple(ops.Vq @ q for q in Q))
    VU = tuple(ops.VhP @ v for v in VU)
    Uh = P.u_of_v(VU)
    beta = P.beta_of_u(Uh)
    Qh = (Uh[0], Uh[1] / Uh[0], Uh[2] / Uh[0], beta)
    QM = tuple(q[Nq:, :] for q in Qh)
    QP = tuple(_gather(q, md.mapP) for q in QM)
    Uf = tuple(u[Nq:, :] for u in Uh)
    rhoUn = (Uf[1] * md.nxJ + Uf[2] * md.nyJ) / md.sJ
    lam = np.abs(P.wavespeed(Uf[0], rhoUn, Uf[3]))
    LFc = lf_scale * np.maximum(lam, _gather(lam, md.mapP)) * md.sJ
    fSx, fSy = P.euler_fluxes_2d(QM, QP)
    rhsQ = []
    for f in range(nf):
        flux = fSx[f] * md.nxJ + fSy[f] * md.nyJ - LFc * (_gather(Uf[f], md.mapP) - Uf[f])
        rhsQ.append(ops.Lf @ flux)
    vgeo = tuple(a[0, :] for a in (md.rxJ, md.sxJ, md.ryJ, md.syJ))
    QF = dense_hadamard_sum(Qh, ops, vgeo)
    for f in range(nf):
        rhsQ[f] = rhsQ[f] + 2 * (ops.Ph @ QF[f])
    rhsQ = [-x / md.J for x in rhsQ]
    rhstest = 0.0
    if compute_rhstest:
        for f in range(nf):
            rhstest += np.sum(md.wJq * VU[f][:Nq, :] * (ops.Vq @ rhsQ[f]))
    return rhsQ, rhstest


def lsrk45_step(Q, resQ, dt, rhs_fun):
    """One 5-stage step (dg2D_euler_quad.jl:198-207). Returns Q, resQ, last rhstest."""
    rk4a, rk4b, _ = rk45_coeffs()
    rhstest = 0.0
    for INTRK in range(5):
        rhsQ, rhstest = rhs_fun(Q, INTRK == 4)
        resQ = [rk4a[INTRK] * r + dt * rq for r, rq in zip(resQ, rhsQ)]
        Q = [q + rk4b[INTRK] * r for q, r in zip(Q, resQ)]
    return Q, resQ, rhstest


# ------------------------------------------------------------- case setup
def setup_quad_vortex(N=3, Kx=16, Ky=16):
    """Config 1 (dg2D_euler_quad.jl:21-99 with H3: uniform_quad_mesh(Kx,Ky) directly)."""
    VX, VY, EToV = M.uniform_quad_mesh(Kx, Ky)
    VX = 15 * (1 + VX) / 2
    VY = 5 * VY
    rd = S.init_reference_quad(N)
    md = S.init_mesh((VX, VY), EToV, rd)
    S.make_periodic(md, rd, (VX.max() - VX.min(), VY.max() - VY.min()))
    ops = build_sbp_ops(rd)
    rho, u, v, p = P.vortex(md.xq, md.yq, 0)
    Q = list(P.primitive_to_conservative(rho, (u, v), p))
    interp_geofacs(md, rd, ops)
    return rd, md, ops, Q


def setup_tri_vortex(N=4, Kx=8, Ky=8):
    """Config 2 shape (dg2D_euler_tri.jl:19-85)."""
    VX, VY, EToV = M.uniform_tri_mesh(Kx, Ky)
    VX = 15 * (1 + VX) / 2
    VY = 5 * VY
    rd = S.init_reference_tri(N)
    md = S.init_mesh((VX, VY), EToV, rd)
    S.make_periodic(md, rd, (VX.max() - VX.min(), VY.max() - VY.min()))
    ops = build_sbp_ops(rd)
    rho, u, v, p = P.vortex(md.x, md.y, 0)
    Q = list(P.primitive_to_conservative(rho, (u, v), p))
    interp_geofacs(md, rd, ops)
    return rd, md, ops, Q


def setup_hex(N=2, K1D=4, scale=(1.0, 1.0, 1.0), shift=(0.0, 0.0, 0.0)):
    """Hex mesh on scale*[-1,1]^3+shift, periodic (dg3D_euler_hex.jl:26-101)."""
    if np.isscalar(K1D):
        K1D = (K1D, K1D, K1D)
    VX, VY, VZ, EToV = M.uniform_hex_mesh(*K1D)
    VX = scale[0] * VX + shift[0]
    VY = scale[1] * VY + shift[1]
    VZ = scale[2] * VZ + shift[2]
    rd = S.init_reference_hex(N)
    md = S.init_mesh((VX, VY, VZ), EToV, rd)
    S.make_periodic(md, rd, (2 * scale[0], 2 * scale[1], 2 * scale[2]))
    ops = build_sbp_ops(rd)
    interp_geofacs(md, rd, ops)
    return rd, md, ops


def setup_hex_curved(N=2, K1D=3, amp=.08, scale=(1.0, 1.0, 1.0), seed=1):
    """``setup_hex`` with the interior mesh vertices displaced (trilinear, non-affine elements): the
    curved-mesh branch of dg3D_euler_hex.jl (geometric factors at every hybridized point, :69-101, averaged per
    pair in the Hadamard sum, :145-151).  Boundary vertices stay put, so the periodic matching is unchanged."""
    if np.isscalar(K1D):
        K1D = (K1D, K1D, K1D)
    VX, VY, VZ, EToV = M.uniform_hex_mesh(*K1D)
    rng = np.random.default_rng(seed)
    inner = (np.abs(VX) < 1 - 1e-12) & (np.abs(VY) < 1 - 1e-12) & (np.abs(VZ) < 1 - 1e-12)
    h = 2.0 / np.asarray(K1D)
    VX = VX + inner * amp * h[0] * rng.uniform(-1, 1, VX.shape)
    VY = VY + inner * amp * h[1] * rng.uniform(-1, 1, VY.shape)
    VZ = VZ + inner * amp * h[2] * rng.uniform(-1, 1, VZ.shape)
    VX, VY, VZ = scale[0] * VX, scale[1] * VY, scale[2] * VZ
    rd = S.init_reference_hex(N)
    md = S.init_mesh((VX, VY, VZ), EToV, rd)
    S.make_periodic(md, rd, (2 * scale[0], 2 * scale[1], 2 * scale[2]))
    ops = build_sbp_ops(rd)
    interp_geofacs(md, rd, ops)
    return rd, md, ops
