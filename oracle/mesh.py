"""Uniform meshes, connectivity, node maps, geometric factors (oracle).

Restates ``src/Uniform{Quad,Hex,Tri}Mesh.jl``, ``src/connect_mesh.jl``,
``src/node_map_functions.jl`` and ``src/geometric_factors.jl``.  All index
arrays returned here are **0-based** (the Julia ones are 1-based); layouts are
otherwise identical (column-major ``Nfq x K`` etc.).
"""
import numpy as np

from .basis import meshgrid2, vec


# ------------------------------------------------------ UniformQuadMesh.jl
def uniform_quad_mesh(Nx, Ny):  # :25-50
    Nxp, Nyp = Nx + 1, Ny + 1
    x1D = np.linspace(-1, 1, Nxp)
    y1D = np.linspace(-1, 1, Nyp)
    x, y = meshgrid2(x1D, y1D)            # (Nyp, Nxp)
    I, J = meshgrid2(np.arange(1, Nxp + 1), np.arange(1, Nyp + 1))
    inds = (I - 1) * Ny + (J + I - 1)     # 1-based, y fastest
    EToV = np.zeros((Nx * Ny, 4), dtype=np.int64)
    k = 0
    for i in range(Ny):
        for j in range(Nx):
            EToV[k, :] = [inds[i, j], inds[i, j + 1], inds[i + 1, j], inds[i + 1, j + 1]]
            k += 1
    return vec(x), vec(y), EToV - 1


def quad_face_vertices():  # :67-69 (0-based)
    return [0, 1], [1, 3], [2, 3], [0, 2]


# ------------------------------------------------------- UniformTriMesh.jl
def uniform_tri_mesh(Kx, Ky):  # :25-42
    VY, VX = meshgrid2(np.linspace(-1, 1, Ky + 1), np.linspace(-1, 1, Kx + 1))  # (Kx+1, Ky+1)
    EToV = np.zeros((2 * Kx * Ky, 3), dtype=np.int64)
    sk = 0
    for ey in range(1, Ky + 1):
        for ex in range(1, Kx + 1):
            idf = lambda a, b: a + (b - 1) * (Kx + 1)
            id1, id2, id3, id4 = idf(ex, ey), idf(ex + 1, ey), idf(ex + 1, ey + 1), idf(ex, ey + 1)
            EToV[2 * sk, :] = [id1, id2, id3]
            EToV[2 * sk + 1, :] = [id3, id4, id1]
            sk += 1
    return vec(VX), vec(VY), EToV - 1


def tri_face_vertices():  # :49-51
    return [0, 1], [1, 2], [2, 0]


# ------------------------------------------------------- UniformHexMesh.jl
def uniform_hex_mesh(Nx, Ny, Nz):  # :25-80
    Nxp, Nyp, Nzp = Nx + 1, Ny + 1, Nz + 1
    x1D = np.linspace(-1, 1, Nxp)
    y1D = np.linspace(-1, 1, Nyp)
    z1D = np.linspace(-1, 1, Nzp)
    kk, jj, ii = np.meshgrid(np.arange(Nzp), np.arange(Nyp), np.arange(Nxp), indexing="ij")
    x = x1D[ii.ravel()]
    y = y1D[jj.ravel()]
    z = z1D[kk.ravel()]                     # x fastest, then y, then z
    K = Nx * Ny * Nz
    em = np.arange(K)
    k = em // (Nx * Ny)
    j = (em - k * Nx * Ny) // Nx
    i = em % Nx
    EToV = np.zeros((K, 8), dtype=np.int64)
    for v, (di, dj, dk) in enumerate([(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0),
                                      (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]):
        EToV[:, v] = (i + di) + Nxp * (j + dj) + Nxp * Nyp * (k + dk)
    return x, y, z, EToV


def hex_face_vertices():
    """Vertices on r=-1, r=+1, s=-1, s=+1, t=-1, t=+1 (0-based).

    DEVIATION: ``UniformHexMesh.jl:87-97`` takes ``x->x[1]`` of
    ``CartesianIndex`` results (degenerate lists on Julia >= 0.7).  The intent
    -- the four vertices of each face -- is stated here for the r-fastest
    vertex numbering the fixed node ordering uses.
    """
    return [0, 2, 4, 6], [1, 3, 5, 7], [0, 1, 4, 5], [2, 3, 6, 7], [0, 1, 2, 3], [4, 5, 6, 7]


# ---------------------------------------------------------- connect_mesh.jl
def connect_mesh(EToV, fv):  # :17-36
    Nfaces = len(fv)
    K = EToV.shape[0]
    # face list in column-major (face fastest) order, like the comprehension
    fnodes = [tuple(sorted(EToV[e, ids])) for e in range(K) for ids in fv]
    p = sorted(range(len(fnodes)), key=lambda q: fnodes[q])
    FToF = np.arange(Nfaces * K, dtype=np.int64)
    for f in range(len(p) - 1):
        if fnodes[p[f]] == fnodes[p[f + 1]]:
            f1, f2 = FToF[p[f]], FToF[p[f + 1]]
            FToF[p[f]] = f2
            FToF[p[f + 1]] = f1
    return FToF.reshape((Nfaces, K), order="F")


# ----------------------------------------------------- node_map_functions.jl
def build_node_maps(Xf, FToF):  # :23-55
    NfacesK = FToF.size
    NODETOL = 1e-10
    n = Xf[0].size
    Nfp = n // NfacesK
    mapM = np.arange(n, dtype=np.int64).reshape((Nfp, NfacesK), order="F")
    mapP = mapM.copy()
    Xr = [np.asarray(x).reshape((Nfp, NfacesK), order="F") for x in Xf]
    for f1, f2 in enumerate(FToF.flatten(order="F")):
        D = np.zeros((Nfp, Nfp))
        for X in Xr:
            D += np.abs(X[:, f1][:, None] - X[:, f2][None, :])
        refd = D.max()
        idM, idP = np.nonzero(D < NODETOL * refd)
        mapP[idM, f1] = idP + f2 * Nfp
    mapB = np.nonzero(vec(mapM) == vec(mapP))[0]
    return mapM, mapP, mapB


def build_periodic_boundary_maps_2d(xf, yf, LX, LY, NfacesTotal, mapM, mapP, mapB):  # :66-136
    xb = vec(xf)[mapB]
    yb = vec(yf)[mapB]
    Nfp = xf.size // NfacesTotal
    Nb = xb.size // Nfp
    xb = xb.reshape((Nfp, Nb), order="F")
    yb = yb.reshape((Nfp, Nb), order="F")
    xc = xb.sum(0) / Nfp
    yc = yb.sum(0) / Nfp
    mapMB = vec(mapM)[mapB].reshape((Nfp, Nb), order="F")
    mapPB = vec(mapP)[mapB].reshape((Nfp, Nb), order="F").copy()
    TOL = 1e-12
    yfaces = np.nonzero((np.abs(yc - yc.max()) < TOL * LY) | (np.abs(yc - yc.min()) < TOL * LY))[0]
    xfaces = np.nonzero((np.abs(xc - xc.max()) < TOL * LX) | (np.abs(xc - xc.min()) < TOL * LX))[0]
    for i in yfaces:
        for j in yfaces:
            if i != j and abs(xc[i] - xc[j]) < TOL * LX and abs(abs(yc[i] - yc[j]) - LY) < TOL * LY:
                # Xa[p,q] = xb[q,i], Xb[p,q] = xb[p,j]; ids = row index p of matches, column-major order
                D = np.abs(xb[:, i][None, :] - xb[:, j][:, None])
                pp, qq = np.nonzero((D < TOL * LX).T)  # iterate column-major: q outer
                mapPB[:, i] = mapMB[qq, j]
    for i in xfaces:
        for j in xfaces:
            if i != j and abs(yc[i] - yc[j]) < TOL * LY and abs(abs(xc[i] - xc[j]) - LX) < TOL * LX:
                D = np.abs(yb[:, i][None, :] - yb[:, j][:, None])
                pp, qq = np.nonzero((D < TOL * LY).T)
                mapPB[:, i] = mapMB[qq, j]
    return vec(mapPB)


def build_periodic_boundary_maps_3d(xf, yf, zf, LX, LY, LZ, NfacesTotal, mapM, mapP, mapB):  # :139-213
    Nfp = xf.size // NfacesTotal
    xb, yb, zb = (vec(a)[mapB] for a in (xf, yf, zf))
    Nb = xb.size // Nfp
    xb, yb, zb = (a.reshape((Nfp, Nb), order="F") for a in (xb, yb, zb))
    xc, yc, zc = xb.sum(0) / Nfp, yb.sum(0) / Nfp, zb.sum(0) / Nfp
    mapMB = vec(mapM)[mapB].reshape((Nfp, Nb), order="F")
    mapPB = vec(mapP)[mapB].reshape((Nfp, Nb), order="F").copy()
    TOL = 1e-12
    L = (LX, LY, LZ)
    cb = (xb, yb, zb)
    cc = (xc, yc, zc)
    for d in range(3):
        o1, o2 = [q for q in range(3) if q != d]
        faces = np.nonzero((np.abs(cc[d] - cc[d].max()) < TOL * L[d]) | (np.abs(cc[d] - cc[d].min()) < TOL * L[d]))[0]
        # the reference double loop over (i,j) face pairs (:171-181), with the
        # centroid test evaluated as one broadcast
        c1, c2, cd = cc[o1][faces], cc[o2][faces], cc[d][faces]
        ok = ((np.abs(c1[:, None] - c1[None, :]) < TOL * L[o1])
              & (np.abs(c2[:, None] - c2[None, :]) < TOL * L[o2])
              & (np.abs(np.abs(cd[:, None] - cd[None, :]) - L[d]) < TOL * L[d]))
        np.fill_diagonal(ok, False)
        for a, b in zip(*np.nonzero(ok)):
            i, j = faces[a], faces[b]
            D = (np.abs(cb[o1][:, i][None, :] - cb[o1][:, j][:, None])
                 + np.abs(cb[o2][:, i][None, :] - cb[o2][:, j][:, None]))
            pp, qq = np.nonzero((D < TOL * L[o1]).T)
            mapPB[:, i] = mapMB[qq, j]
    return vec(mapPB)


# ------------------------------------------------------ geometric_factors.jl
def geometric_factors_2d(x, y, Dr, Ds):  # :16-27
    xr, xs = Dr @ x, Ds @ x
    yr, ys = Dr @ y, Ds @ y
    J = -xs * yr + xr * ys
    return ys, -yr, -xs, xr, J  # rxJ, sxJ, ryJ, syJ, J


def geometric_factors_3d(x, y, z, Dr, Ds, Dt):  # :34-67 (Kopriva curl form)
    xr, xs, xt = Dr @ x, Ds @ x, Dt @ x
    yr, ys, yt = Dr @ y, Ds @ y, Dt @ y
    zr, zs, zt = Dr @ z, Ds @ z, Dt @ z
    Fr, Fs, Ft = (Dr @ y) * z, (Ds @ y) * z, (Dt @ y) * z
    rxJ = Dt @ Fs - Ds @ Ft
    sxJ = Dr @ Ft - Dt @ Fr
    txJ = Ds @ Fr - Dr @ Fs
    Fr, Fs, Ft = (Dr @ x) * z, (Ds @ x) * z, (Dt @ x) * z
    ryJ = -(Dt @ Fs - Ds @ Ft)
    syJ = -(Dr @ Ft - Dt @ Fr)
    tyJ = -(Ds @ Fr - Dr @ Fs)
    Fr, Fs, Ft = (Dr @ y) * x, (Ds @ y) * x, (Dt @ y) * x
    rzJ = -(Dt @ Fs - Ds @ Ft)
    szJ = -(Dr @ Ft - Dt @ Fr)
    tzJ = -(Ds @ Fr - Dr @ Fs)
    J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)
    return rxJ, sxJ, txJ, ryJ, syJ, tyJ, rzJ, szJ, tzJ, J
