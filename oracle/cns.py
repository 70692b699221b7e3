"""Compressible Navier-Stokes rhs with entropy-stable wall BCs (oracle; test infrastructure only).

Restates ``examples/CompressibleNS/dg2D_CNS_cavity_optimized.jl``:
  init_BC_funs                :135-264   (wall / lid node sets, three BC closures)
  update_flux!                :308-324
  flux_differencing!          :326-347   (symmetric loop; here the equivalent dense Hadamard sum)
  rhs_inviscid!               :447-528
  dg_grad! / dg_div!          :548-569 / :590-611
  init_visc_fxn               :613-645   (note ``let lambda = -lambda``)
  rhs_viscous!                :749-849
  rhsRK!                      :955-972
The reference runs this on triangles; BASELINE config 3 applies the same rhs to the quad element
(``init_reference_quad``) -- both go through ``build_cns_ops``.
"""
from types import SimpleNamespace

import numpy as np

from . import physics as P
from . import setupdg as S
from . import mesh as M
from .solvers import _gather, dense_hadamard_sum

GAMMA = P.GAMMA


def build_cns_ops(rd):
    """Operator tuple of the CNS scripts (:61-89): modal-style hybridized SBP ops for any rd."""
    ops = SimpleNamespace()
    ops.elem, ops.N, ops.Nfaces, ops.wq = rd.elem, rd.N, rd.Nfaces, rd.wq
    Ef = rd.Vf @ rd.Pq
    Qh = []
    for D, n in ((rd.Dr, rd.nrJ), (rd.Ds, rd.nsJ)):
        Q = rd.Pq.T @ rd.M @ D @ rd.Pq
        B = np.diag(rd.wf * n)
        Qh.append(.5 * np.block([[Q - Q.T, Ef.T @ B], [-B @ Ef, B]]))
    ops.Qskew = [.5 * (A - A.T) for A in Qh]
    Vh = np.vstack([rd.Vq, rd.Vf])
    ops.Vh = Vh
    ops.Ph = np.linalg.solve(rd.M, Vh.T)
    ops.VhP = Vh @ rd.Pq
    ops.Lf = rd.LIFT
    ops.Vq = rd.Vq
    ops.Pq, ops.Vf, ops.Dr, ops.Ds = rd.Pq, rd.Vf, rd.Dr, rd.Ds
    ops.volume_scale = 2.0
    ops.lf_scale = .25
    ops.modal = True
    return ops


def cavity_bc_sets(md):
    """Boundary node sets of the lid-driven cavity on [-1,1]^2 (:139-155); 0-based trace indices."""
    xb = md.xf.flatten(order="F")[md.mapB]
    yb = md.yf.flatten(order="F")[md.mapB]
    lid = md.mapB[np.abs(yb - 1) < 1e-12]
    wall = md.mapB[np.abs(yb - 1) >= 1e-12]
    bc = SimpleNamespace(lid=lid, wall=wall, boundary=np.concatenate([lid, wall]), mapB=md.mapB)
    bc.vlid = np.ones(lid.size)
    nx = (md.nxJ / md.sJ).flatten(order="F")
    ny = (md.nyJ / md.sJ).flatten(order="F")
    bc.nx, bc.ny = nx[bc.boundary], ny[bc.boundary]
    return bc


def _flat(a):
    return a.flatten(order="F")


def impose_BCs_inviscid(QP, Qf, bc):  # :157-176
    b = bc.boundary
    QP = [_flat(q).copy() for q in QP]
    Qf = [_flat(q) for q in Qf]
    u1, u2 = Qf[1][b], Qf[2][b]
    Un = u1 * bc.nx + u2 * bc.ny
    QP[0][b] = Qf[0][b]
    QP[3][b] = Qf[3][b]
    QP[1][b] = Qf[1][b] - 2 * Un * bc.nx
    QP[2][b] = Qf[2][b] - 2 * Un * bc.ny
    if getattr(bc, "inflow", None) is not None:             # dg2D_CNS_modalESDG.jl:171-181
        rho, u, v, p = bc.inflow_state
        i, o = bc.inflow, bc.outflow
        QP[0][i], QP[1][i], QP[2][i], QP[3][i] = rho, u, v, rho / (2 * p)
        for k in range(4):
            QP[k][o] = Qf[k][o]
    return QP


def impose_BCs_entropyvars(VUP, VUf, bc, bctype):  # :178-216
    w, l = bc.wall, bc.lid
    if bctype == 1:
        VUP[1][w], VUP[2][w], VUP[3][w] = -VUf[1][w], -VUf[2][w], VUf[3][w]
        VUP[1][l] = -VUf[1][l] - 2 * bc.vlid * VUf[3][l]
        VUP[2][l] = -VUf[2][l]
        VUP[3][l] = VUf[3][l]
    elif bctype == 2:
        theta = 1.0 / 0.3 ** 2 / 1.4 / 0.4
        VUP[1][w], VUP[2][w], VUP[3][w] = -VUf[1][w], -VUf[2][w], -2.0 / theta - VUf[3][w]
        VUP[1][l] = 2.0 / theta - VUf[1][l]
        VUP[2][l] = -VUf[2][l]
        VUP[3][l] = -2.0 / theta - VUf[3][l]
    elif bctype == 3:
        b = bc.boundary
        v1, v2 = VUf[1][b], VUf[2][b]
        VUn = v1 * bc.nx + v2 * bc.ny
        VUP[3][b] = VUf[3][b]
        VUP[1][b] = VUf[1][b] - 2 * VUn * bc.nx
        VUP[2][b] = VUf[2][b] - 2 * VUn * bc.ny
    if getattr(bc, "inflow", None) is not None:             # dg2D_CNS_modalESDG.jl:190-203
        rho, u, v, p = bc.inflow_state
        E = p / (GAMMA - 1) + .5 * rho * (u * u + v * v)
        VL = P.v_of_u((np.array([rho]), np.array([rho * u]), np.array([rho * v]), np.array([E])))
        for k in range(4):
            VUP[k][bc.inflow] = VL[k][0]
            VUP[k][bc.outflow] = VUf[k][bc.outflow]
    return VUP


def impose_BCs_stress(sxP, syP, sxf, syf, bc, bctype):  # :218-262
    w, l, b = bc.wall, bc.lid, bc.boundary
    if bctype == 1:
        for k in (1, 2):
            for s_, f_ in ((sxP, sxf), (syP, syf)):
                s_[k][w] = f_[k][w]
                s_[k][l] = f_[k][l]
        sxP[3][w], syP[3][w] = -sxf[3][w], -syf[3][w]
        sxP[3][l] = -sxf[3][l] + 2 * bc.vlid * sxf[1][l]
        syP[3][l] = -syf[3][l] + 2 * bc.vlid * syf[1][l]
    elif bctype == 2:
        for k in (1, 2, 3):
            sxP[k][b], syP[k][b] = sxf[k][b], syf[k][b]
    elif bctype == 3:
        snx = sxf[1][b] * bc.nx + sxf[2][b] * bc.ny
        sny = syf[1][b] * bc.nx + syf[2][b] * bc.ny
        sxP[1][b] = -sxf[1][b] + 2 * bc.nx * snx
        syP[1][b] = -syf[1][b] + 2 * bc.nx * sny
        sxP[2][b] = -sxf[2][b] + 2 * bc.ny * snx
        syP[2][b] = -syf[2][b] + 2 * bc.ny * sny
        sxP[3][b], syP[3][b] = -sxf[3][b], -syf[3][b]
    if getattr(bc, "inflow", None) is not None:             # dg2D_CNS_modalESDG.jl:205-215: sigma+ = sigma on `vwall`
        vw = np.concatenate([bc.inflow, bc.outflow])
        for k in range(4):
            sxP[k][vw], syP[k][vw] = sxf[k][vw], syf[k][vw]
    return sxP, syP


def rhs_inviscid(Q, md, ops, bc, inviscid_dissp=True):
    """rhs_inviscid! (:447-528) with compute_rhstest = false (as rhsRK! calls it)."""
    Nh, Nq = ops.VhP.shape
    shape_f = md.nxJ.shape
    VU = P.v_of_u(tuple(ops.Vq @ q for q in Q))
    Uh = P.u_of_v(tuple(ops.VhP @ v for v in VU))
    beta = P.beta_of_u(Uh)
    Qh = (Uh[0], Uh[1] / Uh[0], Uh[2] / Uh[0], beta)
    QM = tuple(q[Nq:, :] for q in Qh)
    QP = tuple(_gather(q, md.mapP) for q in QM)
    QP = [q.reshape(shape_f, order="F") for q in impose_BCs_inviscid(QP, QM, bc)]
    Uf = tuple(u[Nq:, :] for u in Uh)
    rhoUn = (Uf[1] * md.nxJ + Uf[2] * md.nyJ) / md.sJ
    lam = np.abs(np.sqrt(np.abs(rhoUn / Uf[0])) + np.sqrt(1.4 * 0.4 * (Uf[3] - .5 * rhoUn ** 2 / Uf[0]) / Uf[0]))
    lamP = _gather(lam, md.mapP)
    if getattr(bc, "inflow", None) is not None:             # impose_BCs_lam! (dg2D_CNS_modalESDG.jl:183-188, :320-323)
        vw = np.concatenate([bc.inflow, bc.outflow])
        lam_f, lamP_f = _flat(lam).copy(), _flat(lamP).copy()
        lam_f[vw] = 0.0
        lamP_f[vw] = 0.0
        lam, lamP = lam_f.reshape(shape_f, order="F"), lamP_f.reshape(shape_f, order="F")
    LFc = .25 * np.maximum(lam, lamP) * md.sJ
    UP = tuple(_gather(u, md.mapP) for u in Uf)
    fx, fy = P.euler_fluxes_2d(tuple(QP), QM)                 # update_flux! evaluates (QP, QM)  (:311)
    rhsQ = []
    for f in range(4):
        flux = fx[f] * md.nxJ + fy[f] * md.nyJ
        if inviscid_dissp:
            flux = flux - LFc * (UP[f] - Uf[f])
        rhsQ.append(ops.Lf @ flux)
    vgeo = tuple(a[0, :] for a in (md.rxJ, md.sxJ, md.ryJ, md.syJ))
    QF = dense_hadamard_sum(Qh, ops, vgeo)                     # == flux_differencing!/2 (both (i,j),(j,i))
    rhsQ = [r + 2 * (ops.Ph @ qf) for r, qf in zip(rhsQ, QF)]
    return [-x / md.J for x in rhsQ]


def viscous_matrices(v, lam_in, mu, Pr):  # :613-645  (lambda sign flipped inside, H7)
    lam = -lam_in
    v1, v2, v3, v4 = v
    c = 1 / (v4 ** 3)
    l2m = lam + 2.0 * mu
    Z = np.zeros_like(v4)
    Kxx = [[Z] * 4 for _ in range(4)]
    Kxy = [[Z] * 4 for _ in range(4)]
    Kyy = [[Z] * 4 for _ in range(4)]
    Kxx[1][1] = c * -l2m * v4 ** 2
    Kxx[1][3] = c * l2m * v2 * v4
    Kxx[2][2] = c * -mu * v4 ** 2
    Kxx[2][3] = c * mu * v3 * v4
    Kxx[3][1] = c * l2m * v2 * v4
    Kxx[3][2] = c * mu * v3 * v4
    Kxx[3][3] = c * -(l2m * v2 ** 2 + mu * v3 ** 2 - GAMMA * mu * v4 / Pr)
    Kxy[1][2] = c * -lam * v4 ** 2
    Kxy[1][3] = c * lam * v3 * v4
    Kxy[2][1] = c * -mu * v4 ** 2
    Kxy[2][3] = c * mu * v2 * v4
    Kxy[3][1] = c * mu * v3 * v4
    Kxy[3][2] = c * lam * v2 * v4
    Kxy[3][3] = c * (lam + mu) * (-v2 * v3)
    Kyy[1][1] = c * -mu * v4 ** 2
    Kyy[1][3] = c * mu * v2 * v4
    Kyy[2][2] = c * -l2m * v4 ** 2
    Kyy[2][3] = c * l2m * v3 * v4
    Kyy[3][1] = c * mu * v2 * v4
    Kyy[3][2] = c * l2m * v3 * v4
    Kyy[3][3] = c * -(l2m * v3 ** 2 + mu * v2 ** 2 - GAMMA * mu * v4 / Pr)
    return Kxx, Kxy, Kyy


def dg_grad(V, Vf_, VP, md, ops):  # :548-569
    Np = ops.Dr.shape[0]
    rxj, sxj, ryj, syj = (a[:Np, :] for a in (md.rxJ, md.sxJ, md.ryJ, md.syJ))
    outx, outy = [], []
    for v, vf, vp in zip(V, Vf_, VP):
        vr, vs = ops.Dr @ v, ops.Ds @ v
        outx.append((rxj * vr + sxj * vs + ops.Lf @ (.5 * (vp - vf) * md.nxJ)) / md.J)
        outy.append((ryj * vr + syj * vs + ops.Lf @ (.5 * (vp - vf) * md.nyJ)) / md.J)
    return outx, outy


def dg_div(Sx, Sxf, SxP, Sy, Syf, SyP, md, ops):  # :590-611
    Np = ops.Dr.shape[0]
    rxj, sxj, ryj, syj = (a[:Np, :] for a in (md.rxJ, md.sxJ, md.ryJ, md.syJ))
    out = []
    for k in range(len(Sx)):
        vol = rxj * (ops.Dr @ Sx[k]) + sxj * (ops.Ds @ Sx[k]) + ryj * (ops.Dr @ Sy[k]) + syj * (ops.Ds @ Sy[k])
        surf = ops.Lf @ (.5 * ((SxP[k] - Sxf[k]) * md.nxJ + (SyP[k] - Syf[k]) * md.nyJ))
        out.append((vol + surf) / md.J)
    return out


def rhs_viscous(Q, md, ops, bc, par):
    """rhs_viscous! (:749-849).  par: Re, mu, lam (as passed to init_visc_fxn), Pr, bctype, viscous_dissp."""
    shape_f = md.nxJ.shape
    VUq = P.v_of_u(tuple(ops.Vq @ q for q in Q))
    VU = [ops.Pq @ v for v in VUq]
    VUf = [ops.Vf @ v for v in VU]
    VUP = [_gather(v, md.mapP) for v in VUf]
    VUPf = impose_BCs_entropyvars([_flat(v).copy() for v in VUP], [_flat(v) for v in VUf], bc, par.bctype)
    VUP = [v.reshape(shape_f, order="F") for v in VUPf]
    VUx, VUy = dg_grad(VU, VUf, VUP, md, ops)
    VUx = [ops.Vq @ v for v in VUx]
    VUy = [ops.Vq @ v for v in VUy]
    VUqq = [ops.Vq @ v for v in VU]
    Kxx, Kxy, Kyy = viscous_matrices(VUqq, par.lam, par.mu, par.Pr)
    sx = [np.zeros_like(VUqq[0]) for _ in range(4)]
    sy = [np.zeros_like(VUqq[0]) for _ in range(4)]
    for col in range(1, 4):
        for row in range(1, 4):
            sx[row] = sx[row] + (Kxx[row][col] * VUx[col] + Kxy[row][col] * VUy[col])
            sy[row] = sy[row] + (Kxy[col][row] * VUx[col] + Kyy[row][col] * VUy[col])
    rhstest = 0.0
    for f in range(4):
        rhstest += np.sum(md.wJq * VUx[f] * sx[f])
        rhstest += np.sum(md.wJq * VUy[f] * sy[f])
    sx = [ops.Pq @ s for s in sx]
    sy = [ops.Pq @ s for s in sy]
    sxf = [ops.Vf @ s for s in sx]
    syf = [ops.Vf @ s for s in sy]
    sxP = [_flat(_gather(s, md.mapP)).copy() for s in sxf]
    syP = [_flat(_gather(s, md.mapP)).copy() for s in syf]
    sxP, syP = impose_BCs_stress(sxP, syP, [_flat(s) for s in sxf], [_flat(s) for s in syf], bc, par.bctype)
    sxP = [s.reshape(shape_f, order="F") for s in sxP]
    syP = [s.reshape(shape_f, order="F") for s in syP]
    rhs = dg_div(sx, sxf, sxP, sy, syf, syP, md, ops)
    if par.viscous_dissp:
        tau = -1 / par.Re / VUf[3]
        dV = [vp - vf for vp, vf in zip(VUP, VUf)]
        avgV = [.5 * (vp + vf) for vp, vf in zip(VUP, VUf)]
        pen = [np.zeros(shape_f)] + [tau * dV[k] for k in (1, 2, 3)]
        pen = [_flat(p_).copy() for p_ in pen]
        mB = bc.mapB
        tB = _flat(tau)[mB]
        dVf = [_flat(d) for d in dV]
        aVf = [_flat(a) for a in avgV]
        v4B = _flat(VUf[3])[mB]
        pen[1][mB] = tB * dVf[1][mB]
        pen[2][mB] = tB * dVf[2][mB]
        if par.bctype == 1:
            pen[3][mB] = -tB * (aVf[1][mB] * dVf[1][mB] + aVf[2][mB] * dVf[2][mB]) / v4B
        else:
            pen[3][mB] = -tB * (aVf[1][mB] * dVf[1][mB] + aVf[2][mB] * dVf[2][mB] + dVf[3][mB] * dVf[3][mB] / 2) / v4B
        pen = [ops.Lf @ p_.reshape(shape_f, order="F") for p_ in pen]
        rhs = [r + p_ for r, p_ in zip(rhs, pen)]
    return rhs, rhstest


def rhsRK(Q, md, ops, bc, par):
    """rhsRK! (:955-972): returns (rhsQ, rhstest, rhstest_visc)."""
    rhsQ = rhs_inviscid(Q, md, ops, bc, par.inviscid_dissp)
    visc, visc_test = rhs_viscous(Q, md, ops, bc, par)
    rhsQ = [a + b for a, b in zip(rhsQ, visc)]
    VU = P.v_of_u(tuple(ops.Vq @ q for q in Q))
    VUq = [ops.Vq @ (ops.Pq @ v) for v in VU]
    rhstest = sum(np.sum(md.wJq * VUq[f] * (ops.Vq @ rhsQ[f])) for f in range(4))
    rhstest_visc = sum(np.sum(md.wJq * VUq[f] * (ops.Vq @ visc[f])) for f in range(4)) + visc_test
    return rhsQ, rhstest, rhstest_visc


def cavity_params(Re=1000.0, bctype=1, inviscid_dissp=True, viscous_dissp=True):
    """Script-top globals (:21-36); lambda is passed as -2/3 mu and flipped inside (H7)."""
    mu = 1 / Re
    return SimpleNamespace(Re=Re, mu=mu, lam=-2 / 3 * mu, Pr=.71, bctype=bctype,
                           inviscid_dissp=inviscid_dissp, viscous_dissp=viscous_dissp)


def setup_cavity(elem="tri", N=3, K1D=4):
    """Lid-driven cavity on [-1,1]^2 (:47-52, :859-865): rho=1, u=v=0, p=1/(Ma^2 gamma), Ma=.3."""
    if elem == "tri":
        VX, VY, EToV = M.uniform_tri_mesh(K1D, K1D)
        rd = S.init_reference_tri(N)
    else:
        VX, VY, EToV = M.uniform_quad_mesh(K1D, K1D)
        rd = S.init_reference_quad(N)
    md = S.init_mesh((VX, VY), EToV, rd)
    ops = build_cns_ops(rd)
    for name in ("rxJ", "sxJ", "ryJ", "syJ"):
        setattr(md, name, ops.Vh @ getattr(md, name))
    bc = cavity_bc_sets(md)
    Ma = .3
    rho = np.ones_like(md.x)
    Q = list(P.primitive_to_conservative(rho, (np.zeros_like(rho), np.zeros_like(rho)), (1 / (Ma ** 2 * GAMMA)) * np.ones_like(rho)))
    return rd, md, ops, bc, Q


def dopri45_coeffs():  # :919-934
    a = np.zeros((7, 7))
    a[1, 0] = .2
    a[2, :2] = [3.0 / 40.0, 9.0 / 40.0]
    a[3, :3] = [44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0]
    a[4, :4] = [19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0]
    a[5, :5] = [9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0]
    a[6, :6] = [35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0]
    c = np.array([0.0, .2, .3, .8, 8.0 / 9.0, 1.0, 1.0])
    E = np.array([71.0 / 57600.0, 0.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0, -1.0 / 40.0])
    return a, E, c


def dopri45(Q, rhs_fun, T, dt, err_tol=1e-5, max_attempts=10 ** 9):
    """The adaptive loop of dg2D_CNS_cavity_optimized.jl:974-1053 (FSAL, Hairer semi-norm, P/PI controller).
    ``rhs_fun(Q) -> (rhsQ, rhstest, rhstest_visc)``.  Quirks kept: the error estimate is NOT scaled by dt
    (:1014-1017), the loop does not clip the last step to T (:1001)."""
    rka, rkE, _ = dopri45_coeffs()
    Q = [q.copy() for q in Q]
    dt0 = dt
    K = [None] * 7
    K[0] = rhs_fun(Q)[0]
    t, i, prev = 0.0, 0, 0.0
    hist = {"t": [0.0], "dt": [dt], "err": [0.0], "rhstest": [0.0], "rhstest_visc": [0.0], "accepted": []}
    while t < T and i < max_attempts:
        for k in range(1, 7):
            kk = [np.zeros_like(q) for q in Q]
            for s_ in range(k):
                kk = [a + rka[k, s_] * b for a, b in zip(kk, K[s_])]
            Qtmp = [q + dt * a for q, a in zip(Q, kk)]
            K[k], rt, rtv = rhs_fun(Qtmp)
        ev = [np.zeros_like(q) for q in Q]
        for s_ in range(7):
            ev = [a + rkE[s_] * b for a, b in zip(ev, K[s_])]
        err = 0.0
        for f in range(len(Q)):
            err += np.sum((np.abs(ev[f]) / (err_tol * (1 + np.abs(Q[f])))) ** 2)
        err = np.sqrt(err / (Q[0].size * 4))
        ok = err < 1.0
        if ok:
            Q = Qtmp
            t += dt
            K[0] = K[6]
        order = 5
        dtnew = .8 * dt * (.9 / err) ** (.4 / (order + 1))
        if i > 0:
            dtnew *= (prev / max(1e-14, err)) ** (.3 / (order + 1))
        dt = max(min(10 * dt0, dtnew), 1e-9)
        prev = err
        i += 1
        hist["t"].append(t); hist["dt"].append(dt); hist["err"].append(err)
        hist["rhstest"].append(rt); hist["rhstest_visc"].append(rtv); hist["accepted"].append(bool(ok))
    return Q, hist


# ---------------------------------------------------------------- viscous shock tube (Becker)
# examples/CompressibleNS/dg2D_CNS_modalESDG.jl: parameters :31-58, mesh :70-86, BCs :161-218,
# exact solution :545-578.
def shocktube_params(viscous_dissp=False):
    g = GAMMA
    M0, mu, v_inf, m0, v0 = 3.0, 0.01, 0.2, 1.0, 1.0
    pr = SimpleNamespace(Re=10000.0, mu=mu, lam=2 / 3 * mu, Pr=3 / 4, bctype=1, inviscid_dissp=True,
                         viscous_dissp=viscous_dissp, M0=M0, v_inf=v_inf, m0=m0, v0=v0)
    pr.v1 = (g - 1 + 2 / M0 ** 2) / (g + 1)
    pr.v01 = np.sqrt(v0 * pr.v1)
    pr.kappa = mu * (g / (g - 1)) / pr.Pr
    pr.cv = 1 / (g - 1)
    uL, uR = v0 + v_inf, pr.v1 + v_inf
    rhoL, rhoR = m0 / v0, m0 / pr.v1
    eL = 1 / (2 * g) * ((g + 1) / (g - 1) * pr.v01 ** 2 - v0 ** 2)
    eR = 1 / (2 * g) * ((g + 1) / (g - 1) * pr.v01 ** 2 - pr.v1 ** 2)
    pr.left = (rhoL, uL, 0.0, (g - 1) * rhoL * eL)        # rho, u, v, p
    pr.right = (rhoR, uR, 0.0, (g - 1) * rhoR * eR)
    return pr


def becker_velocity(x, pr, max_iter=100, tol=1e-14):
    """``bisection_solve_velocity`` (:545-567), vectorised over x."""
    g = GAMMA
    x = np.asarray(x, dtype=float)
    Lk = pr.kappa / pr.m0 / pr.cv
    v0, v1, v01 = pr.v0, pr.v1, pr.v01

    def f(v):
        return -x + 2 * Lk / (g + 1) * (v0 / (v0 - v1) * np.log((v0 - v) / (v0 - v01)) - v1 / (v0 - v1) * np.log((v - v1) / (v01 - v1)))
    vL = np.full(x.shape, v1)
    vR = np.full(x.shape, v0)
    vn = .5 * (vL + vR)
    done = np.zeros(x.shape, dtype=bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        for _ in range(max_iter):
            vn = np.where(done, vn, .5 * (vL + vR))
            fn = f(vn)
            done |= np.abs(fn) < tol
            same = np.sign(f(vL)) == np.sign(fn)
            vL = np.where(~done & same, vn, vL)
            vR = np.where(~done & ~same, vn, vR)
    return vn


def becker_exact(x, t, pr):
    """``exact_sol_viscous_shocktube`` (:573-578): conservative state of the travelling viscous shock."""
    g = GAMMA
    u = becker_velocity(x - pr.v_inf * t, pr)
    rho = pr.m0 / u
    e = 1 / (2 * g) * ((g + 1) / (g - 1) * pr.v01 ** 2 - u ** 2)
    return [rho, rho * (pr.v_inf + u), np.zeros_like(rho), rho * (e + .5 * (pr.v_inf + u) ** 2)]


def shocktube_bc_sets(md, pr):
    """``leftwall`` (inflow) / ``rightwall`` (outflow) of init_BC_funs (:161-169) on [-1/2, 1] x [0, 1]; the mesh is
    made periodic first (:80-86), the top / bottom walls stay periodic."""
    xb = md.xf.flatten(order="F")[md.mapB]
    bc = SimpleNamespace()
    bc.inflow = md.mapB[np.abs(xb + .5) < 1e-12]
    bc.outflow = md.mapB[np.abs(xb - 1.0) < 1e-12]
    bc.inflow_state = pr.left
    e = np.zeros(0, dtype=np.int64)
    bc.wall, bc.lid, bc.boundary, bc.vlid, bc.nx, bc.ny = e, e, e, np.zeros(0), np.zeros(0), np.zeros(0)
    bc.mapB = e                                            # penalisation: interior form everywhere (the script's is commented out)
    return bc


def setup_shocktube(elem="tri", N=2, K1D=8):
    """Mesh, operators, BC sets and the exact initial state of the Becker shock tube (:20-30, :70-117, :580-581)."""
    pr = shocktube_params()
    if elem == "tri":
        VX, VY, EToV = M.uniform_tri_mesh(K1D // 2 * 3, K1D)
        rd = S.init_reference_tri(N)
    else:
        VX, VY, EToV = M.uniform_quad_mesh(K1D // 2 * 3, K1D)
        rd = S.init_reference_quad(N)
    VX = VX / 4 * 3 + 1 / 4
    VY = (VY + 1) / 2
    md = S.init_mesh((VX, VY), EToV, rd)
    bc = shocktube_bc_sets(md, pr)                         # node sets from the ORIGINAL mapB, before the periodic patch
    S.make_periodic(md, rd, (VX.max() - VX.min(), VY.max() - VY.min()))
    ops = build_cns_ops(rd)
    for name in ("rxJ", "sxJ", "ryJ", "syJ"):
        setattr(md, name, ops.Vh @ getattr(md, name))
    Q = becker_exact(md.x, 0.0, pr)
    return rd, md, ops, bc, Q, pr
