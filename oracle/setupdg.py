"""Reference-element and mesh setup (oracle; test infrastructure only).

Restates ``src/SetupDG.jl``: ``RefElemData`` (:38-75), ``MeshData`` (:77-115),
``init_reference_tri`` (:151-201), ``init_reference_quad`` (:205-268),
``init_reference_hex`` (:323-387), ``init_mesh`` 2D (:275-318) / 3D (:389-434).
Field names are the reference's.  Index maps are 0-based.
"""
from types import SimpleNamespace

import numpy as np

from . import basis as B
from . import mesh as M


class RefElemData(SimpleNamespace):
    pass


class MeshData(SimpleNamespace):
    pass


def _droptol(A, tol):
    A = np.array(A, dtype=np.float64, copy=True)
    A[np.abs(A) <= tol] = 0.0   # SparseArrays.droptol! drops |x| <= tol
    return A


def init_reference_quad(N, quad_nodes_1D=None):
    rd = RefElemData()
    rd.elem = "quad"
    rd.N = N
    rd.fv = M.quad_face_vertices()
    rd.Nfaces = len(rd.fv)
    r, s = B.quad_nodes_2D(N)
    VDM = B.quad_vandermonde_2D(N, r, s)
    Vr, Vs = B.quad_grad_vandermonde_2D(N, r, s)
    Dr = np.linalg.solve(VDM.T, Vr.T).T
    Ds = np.linalg.solve(VDM.T, Vs.T).T
    rd.r, rd.s, rd.VDM = r, s, VDM
    r1, s1 = B.quad_nodes_2D(1)
    rd.V1 = np.linalg.solve(B.quad_vandermonde_2D(1, r1, s1).T, B.quad_vandermonde_2D(1, r, s).T).T
    r1D, w1D = quad_nodes_1D if quad_nodes_1D is not None else B.gauss_quad(0, 0, N)
    e = np.ones_like(r1D)
    z = np.zeros_like(r1D)
    rd.rf = np.concatenate([r1D, e, -r1D, -e])
    rd.sf = np.concatenate([-e, r1D, e, -r1D])
    rd.wf = np.tile(w1D, rd.Nfaces)
    rd.nrJ = np.concatenate([z, e, z, -e])
    rd.nsJ = np.concatenate([-e, z, e, z])
    rq, sq = (B.vec(a) for a in B.meshgrid2(r1D))     # s fastest (SetupDG.jl:244)
    wr, ws = (B.vec(a) for a in B.meshgrid2(w1D))
    wq = wr * ws
    Vq = np.linalg.solve(VDM.T, B.quad_vandermonde_2D(N, rq, sq).T).T
    Mm = Vq.T @ np.diag(wq) @ Vq
    Pq = np.linalg.solve(Mm, Vq.T @ np.diag(wq))
    rd.rq, rd.sq, rd.wq, rd.Vq, rd.M, rd.Pq = rq, sq, wq, Vq, Mm, Pq
    Vf = np.linalg.solve(VDM.T, B.quad_vandermonde_2D(N, rd.rf, rd.sf).T).T
    LIFT = np.linalg.solve(Mm, Vf.T @ np.diag(rd.wf))
    rd.Dr = _droptol(Dr, 1e-10)
    rd.Ds = _droptol(Ds, 1e-10)
    rd.Vf = _droptol(Vf, 1e-10)
    rd.LIFT = _droptol(LIFT, 1e-10)
    return rd


def init_reference_hex(N, quad_nodes_1D=None):
    rd = RefElemData()
    rd.elem = "hex"
    rd.N = N
    rd.fv = M.hex_face_vertices()
    rd.Nfaces = len(rd.fv)
    r, s, t = B.hex_nodes_3D(N)
    VDM = B.hex_vandermonde_3D(N, r, s, t)
    Vr, Vs, Vt = B.hex_grad_vandermonde_3D(N, r, s, t)
    Dr, Ds, Dt = (np.linalg.solve(VDM.T, A.T).T for A in (Vr, Vs, Vt))
    rd.r, rd.s, rd.t, rd.VDM = r, s, t, VDM
    r1, s1, t1 = B.hex_nodes_3D(1)
    rd.V1 = np.linalg.solve(B.hex_vandermonde_3D(1, r1, s1, t1).T, B.hex_vandermonde_3D(1, r, s, t).T).T
    r1D, w1D = quad_nodes_1D if quad_nodes_1D is not None else B.gauss_quad(0, 0, N)
    rquad, squad = (B.vec(a) for a in B.meshgrid2(r1D, r1D))   # squad fastest (SetupDG.jl:346)
    wr, ws = (B.vec(a) for a in B.meshgrid2(w1D, w1D))
    wquad = wr * ws
    e = np.ones_like(rquad)
    zz = np.zeros_like(rquad)
    rd.rf = np.concatenate([-e, e, rquad, rquad, rquad, rquad])
    rd.sf = np.concatenate([rquad, rquad, -e, e, squad, squad])
    rd.tf = np.concatenate([squad, squad, squad, squad, -e, e])
    rd.wf = np.tile(wquad, rd.Nfaces)
    rd.nrJ = np.concatenate([-e, e, zz, zz, zz, zz])
    rd.nsJ = np.concatenate([zz, zz, -e, e, zz, zz])
    rd.ntJ = np.concatenate([zz, zz, zz, zz, -e, e])
    rq, sq, tq = B.tensor3_r_fastest(r1D)      # r fastest: deviation H2, see basis.tensor3_r_fastest
    wr, ws, wt = B.tensor3_r_fastest(w1D)
    wq = wr * ws * wt
    Vq = np.linalg.solve(VDM.T, B.hex_vandermonde_3D(N, rq, sq, tq).T).T
    Mm = Vq.T @ np.diag(wq) @ Vq
    Pq = np.linalg.solve(Mm, Vq.T @ np.diag(wq))
    rd.rq, rd.sq, rd.tq, rd.wq, rd.Vq, rd.M, rd.Pq = rq, sq, tq, wq, Vq, Mm, Pq
    Vf = np.linalg.solve(VDM.T, B.hex_vandermonde_3D(N, rd.rf, rd.sf, rd.tf).T).T
    LIFT = np.linalg.solve(Mm, Vf.T @ np.diag(rd.wf))
    rd.Dr, rd.Ds, rd.Dt = (_droptol(A, 1e-12) for A in (Dr, Ds, Dt))
    rd.Vf = _droptol(Vf, 1e-12)
    rd.LIFT = _droptol(LIFT, 1e-12)
    return rd


def init_reference_tri(N):
    rd = RefElemData()
    rd.elem = "tri"
    rd.N = N
    rd.fv = M.tri_face_vertices()
    rd.Nfaces = len(rd.fv)
    r, s = B.tri_nodes_2D(N)
    VDM = B.tri_vandermonde_2D(N, r, s)
    Vr, Vs = B.tri_grad_vandermonde_2D(N, r, s)
    rd.Dr = np.linalg.solve(VDM.T, Vr.T).T
    rd.Ds = np.linalg.solve(VDM.T, Vs.T).T
    rd.r, rd.s, rd.VDM = r, s, VDM
    r1, s1 = B.tri_nodes_2D(1)
    rd.V1 = np.linalg.solve(B.tri_vandermonde_2D(1, r1, s1).T, B.tri_vandermonde_2D(1, r, s).T).T
    r1D, w1D = B.gauss_quad(0, 0, N)
    e = np.ones_like(r1D)
    z = np.zeros_like(r1D)
    rd.rf = np.concatenate([r1D, -r1D, -e])
    rd.sf = np.concatenate([-e, r1D, -r1D])
    rd.wf = np.tile(w1D, 3)
    rd.nrJ = np.concatenate([z, e, -e])
    rd.nsJ = np.concatenate([-e, e, z])
    rq, sq, wq = B.tri_quad_nodes_2D(2 * N)
    Vq = np.linalg.solve(VDM.T, B.tri_vandermonde_2D(N, rq, sq).T).T
    Mm = Vq.T @ np.diag(wq) @ Vq
    Pq = np.linalg.solve(Mm, Vq.T @ np.diag(wq))
    rd.rq, rd.sq, rd.wq, rd.Vq, rd.M, rd.Pq = rq, sq, wq, Vq, Mm, Pq
    rd.Vf = np.linalg.solve(VDM.T, B.tri_vandermonde_2D(N, rd.rf, rd.sf).T).T
    rd.LIFT = np.linalg.solve(Mm, rd.Vf.T @ np.diag(rd.wf))
    return rd


def init_mesh(VXYZ, EToV, rd):
    md = MeshData()
    dim = len(VXYZ)
    md.FToF = M.connect_mesh(EToV, rd.fv)
    Nfaces, K = md.FToF.shape
    md.K = K
    md.EToV = EToV
    md.VX, md.VY = VXYZ[0], VXYZ[1]
    xyz = [rd.V1 @ V[EToV.T] for V in VXYZ]            # Np x K
    xyzf = [rd.Vf @ a for a in xyz]
    mapM, mapP, mapB = M.build_node_maps(xyzf, md.FToF)
    Nfq = rd.Vf.shape[0]
    md.mapM = mapM.reshape((Nfq, K), order="F")
    md.mapP = mapP.reshape((Nfq, K), order="F")
    md.mapB = mapB
    if dim == 2:
        md.x, md.y = xyz
        md.xf, md.yf = xyzf
        md.rxJ, md.sxJ, md.ryJ, md.syJ, md.J = M.geometric_factors_2d(md.x, md.y, rd.Dr, rd.Ds)
        md.xq, md.yq = rd.Vq @ md.x, rd.Vq @ md.y
        md.wJq = np.diag(rd.wq) @ (rd.Vq @ md.J)
        nr, ns = rd.nrJ[:, None], rd.nsJ[:, None]
        md.nxJ = (rd.Vf @ md.rxJ) * nr + (rd.Vf @ md.sxJ) * ns
        md.nyJ = (rd.Vf @ md.ryJ) * nr + (rd.Vf @ md.syJ) * ns
        md.sJ = np.sqrt(md.nxJ ** 2 + md.nyJ ** 2)
    else:
        md.VZ = VXYZ[2]
        md.x, md.y, md.z = xyz
        md.xf, md.yf, md.zf = xyzf
        (md.rxJ, md.sxJ, md.txJ, md.ryJ, md.syJ, md.tyJ,
         md.rzJ, md.szJ, md.tzJ, md.J) = M.geometric_factors_3d(md.x, md.y, md.z, rd.Dr, rd.Ds, rd.Dt)
        md.xq, md.yq, md.zq = rd.Vq @ md.x, rd.Vq @ md.y, rd.Vq @ md.z
        md.wJq = np.diag(rd.wq) @ (rd.Vq @ md.J)
        nr, ns, nt = rd.nrJ[:, None], rd.nsJ[:, None], rd.ntJ[:, None]
        md.nxJ = nr * (rd.Vf @ md.rxJ) + ns * (rd.Vf @ md.sxJ) + nt * (rd.Vf @ md.txJ)
        md.nyJ = nr * (rd.Vf @ md.ryJ) + ns * (rd.Vf @ md.syJ) + nt * (rd.Vf @ md.tyJ)
        md.nzJ = nr * (rd.Vf @ md.rzJ) + ns * (rd.Vf @ md.szJ) + nt * (rd.Vf @ md.tzJ)
        md.sJ = np.sqrt(md.nxJ ** 2 + md.nyJ ** 2 + md.nzJ ** 2)
    return md


def make_periodic(md, rd, L):
    """``mapP[mapB] = mapPB`` (dg2D_euler_quad.jl:39-44, dg3D_euler_hex.jl:60-66)."""
    NfacesTotal = rd.Nfaces * md.K
    mapP = md.mapP.flatten(order="F")
    if len(L) == 2:
        mapPB = M.build_periodic_boundary_maps_2d(md.xf, md.yf, L[0], L[1], NfacesTotal, md.mapM, md.mapP, md.mapB)
    else:
        mapPB = M.build_periodic_boundary_maps_3d(md.xf, md.yf, md.zf, L[0], L[1], L[2], NfacesTotal,
                                                  md.mapM, md.mapP, md.mapB)
    mapP[md.mapB] = mapPB
    md.mapP = mapP.reshape(md.mapP.shape, order="F")
    return md
