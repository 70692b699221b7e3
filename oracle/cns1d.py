"""1-D compressible Navier-Stokes, modal ESDG with SSPRK3 (oracle; TEST INFRASTRUCTURE ONLY).

Restates examples/CompressibleNS/dg1D_CNS_modalESDG.jl operation for operation: Becker's viscous
shock profile on [XL, XR], modal DG on K intervals with a (N+2)-point Gauss rule, entropy projection, flux
differencing with the dense hybridized operator (`flux_differencing!` :204-218), Lax-Friedrichs penalty (:250-260),
BR1 viscous terms with the 1-D coefficient matrix (:269-335), exact-solution states imposed on the two domain ends
(:242-247, :281-286, :321-326) and the SSPRK(3,3) loop (:377-391).  Every function cites the lines it follows.
"""
from types import SimpleNamespace

import numpy as np

from . import basis as B
from . import physics as P

GAMMA = P.GAMMA


def params(M_0=3.0, mu=0.1, v_inf=0.2, XL=-2.0, XR=2.0):  # :93-122
    lam = 2 / 3 * mu
    Pr = 3 / 4
    cp, cv = GAMMA / (GAMMA - 1), 1 / (GAMMA - 1)
    kappa = mu * cp / Pr
    v_0 = 1.0
    rho_0 = 1.0
    m_0 = rho_0 * v_0
    v_1 = (GAMMA - 1 + 2 / M_0 ** 2) / (GAMMA + 1)
    return SimpleNamespace(M_0=M_0, mu=mu, lam=lam, Pr=Pr, cp=cp, cv=cv, kappa=kappa, v_inf=v_inf, rho_0=rho_0, v_0=v_0,
                           m_0=m_0, v_1=v_1, v_01=np.sqrt(v_0 * v_1), XL=XL, XR=XR)


def bisection_solve_velocity(x, pr, max_iter=10000, tol=1e-16):  # :165-189
    v_L, v_R = pr.v_1, pr.v_0
    L_k = pr.kappa / pr.m_0 / pr.cv

    def f(v):
        v = np.float64(v)                       # IEEE semantics at the end points (Julia: 0.0^-a = Inf, log(Inf) = Inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            return -x + 2 * L_k / (GAMMA + 1) * np.log((pr.v_0 - v) ** (pr.v_0 / (pr.v_0 - pr.v_1)) *
                                                     (v - pr.v_1) ** (-pr.v_1 / (pr.v_0 - pr.v_1)))
    v_new = (v_L + v_R) / 2
    it = 0
    while it < max_iter:
        v_new = (v_L + v_R) / 2
        if v_new == v_L or v_new == v_R:        # the interval has collapsed to neighbouring doubles: 1e-16 is never met
            return v_new
        fn = f(v_new)
        if abs(fn) < tol:
            return v_new
        if np.sign(f(v_L)) == np.sign(fn):
            v_L = v_new
        else:
            v_R = v_new
        it += 1
    return v_new


def exact_sol_viscous_shocktube(x, t, pr):  # :194-199
    u = bisection_solve_velocity(x - pr.v_inf * t, pr)
    rho = pr.m_0 / u
    e = 1 / (2 * GAMMA) * ((GAMMA + 1) / (GAMMA - 1) * pr.v_01 ** 2 - u ** 2)
    return rho, rho * (pr.v_inf + u), rho * (e + .5 * (pr.v_inf + u) ** 2)


def setup(N=4, K=16, pr=None):
    """Reference interval (src/SetupDG.jl:117-149), mesh, maps and hybridized operators (:124-160)."""
    pr = pr or params()
    r, _ = B.gauss_lobatto_quad(0, 0, N)
    VDM = B.vandermonde_1D(N, r)
    Dr = np.linalg.solve(VDM.T, B.grad_vandermonde_1D(N, r).T).T
    rq, wq = B.gauss_quad(0, 0, N + 1)
    Vq = np.linalg.solve(VDM.T, B.vandermonde_1D(N, rq).T).T
    M = Vq.T @ np.diag(wq) @ Vq
    Pq = np.linalg.solve(M, Vq.T @ np.diag(wq))
    Vf = np.linalg.solve(VDM.T, B.vandermonde_1D(N, np.array([-1.0, 1.0])).T).T
    LIFT = np.linalg.solve(M, Vf.T)
    VX = np.linspace(pr.XL, pr.XR, K + 1)
    V1 = np.linalg.solve(B.vandermonde_1D(1, np.array([-1.0, 1.0])).T, B.vandermonde_1D(1, r).T).T
    x = V1 @ np.vstack([VX[:-1], VX[1:]])
    mapM = np.arange(2 * K).reshape((2, K), order="F")
    mapP = mapM.copy()
    mapP[0, 1:] = mapM[1, :-1]
    mapP[1, :-1] = mapM[0, 1:]
    nxJ = np.tile(np.array([[-1.0], [1.0]]), (1, K))
    Qr = Pq.T @ M @ Dr @ Pq
    Ef = Vf @ Pq
    Br = np.diag(np.array([-1.0, 1.0]))
    Qrh = .5 * np.block([[Qr - Qr.T, Ef.T @ Br], [-Br @ Ef, Br]])
    Vh = np.vstack([Vq, Vf])
    Ph = np.linalg.solve(M, Vh.T)
    ops = SimpleNamespace(N=N, K=K, r=r, Dr=Dr, Vq=Vq, Pq=Pq, Vf=Vf, LIFT=LIFT, wq=wq, M=M, VhP=Vh @ Pq, Ph=Ph,
                          Qrh_skew=.5 * (Qrh - Qrh.T), rxJ=1.0, J=(pr.XR - pr.XL) / K / 2, mapP=mapP, nxJ=nxJ, x=x,
                          Np=N + 1, Nq=rq.size, Nh=rq.size + 2)
    return ops, pr


def initial_state(ops, pr, t=0.0):  # :202-203
    U = [np.zeros_like(ops.x) for _ in range(3)]
    for idx in np.ndindex(ops.x.shape):
        s = exact_sol_viscous_shocktube(ops.x[idx], t, pr)
        for c in range(3):
            U[c][idx] = s[c]
    return U


def euler_fluxes_1D(UL, UR, logL, logR):  # :39-60
    rhoL, uL, betaL = UL
    rhoR, uR, betaR = UR
    rholog = P.logmean(rhoL, rhoR, logL[0], logR[0])
    betalog = P.logmean(betaL, betaR, logL[1], logR[1])
    rhoavg = .5 * (rhoL + rhoR)
    uavg = .5 * (uL + uR)
    unorm = uL * uR
    pa = rhoavg / (betaL + betaR)
    f4aux = rholog / (2 * (GAMMA - 1) * betalog) + pa + .5 * rholog * unorm
    FxS1 = rholog * uavg
    return FxS1, FxS1 * uavg + pa, f4aux * uavg


def euler_fluxes(UL, UR):  # :62-66
    return euler_fluxes_1D(UL, UR, (P._log(UL[0]), P._log(UL[2])), (P._log(UR[0]), P._log(UR[2])))


def wavespeed_1D(rho, rhou, E):  # :19-23
    p = (GAMMA - 1) * (E - .5 * rhou ** 2 / rho)
    return np.abs(rhou / rho) + np.sqrt(GAMMA * p / rho)


def boundary_data(t, pr):  # :339-349
    rhoL, rhouL, EL = exact_sol_viscous_shocktube(pr.XL, t, pr)
    rhoR, rhouR, ER = exact_sol_viscous_shocktube(pr.XR, t, pr)
    uL, uR = rhouL / rhoL, rhouR / rhoR
    pL = (GAMMA - 1) * (EL - .5 * rhoL * uL ** 2)
    pR = (GAMMA - 1) * (ER - .5 * rhoR * uR ** 2)
    vL = P.v_ufun(np.float64(rhoL), (np.float64(rhoL * uL),), np.float64(EL))
    vR = P.v_ufun(np.float64(rhoR), (np.float64(rhoR * uR),), np.float64(ER))
    return SimpleNamespace(rhoL=rhoL, uL=uL, betaL=rhoL / (2 * pL), rhoR=rhoR, uR=uR, betaR=rhoR / (2 * pR),
                           vL=tuple(float(v) for v in vL), vR=tuple(float(v) for v in vR))


def _gather(a, mapP):
    return a.flatten(order="F")[mapP]


def rhs_inviscid(U, ops, bd):  # :220-268
    Nq, Nh, K = ops.Nq, ops.Nh, ops.K
    VU = P.v_ufun(ops.Vq @ U[0], (ops.Vq @ U[1],), ops.Vq @ U[2])
    VU = tuple(ops.VhP @ v for v in VU)
    Uh = P.u_vfun(VU[0], (VU[1],), VU[2])
    rho, rhou, E = Uh
    beta = P.betafun(rho, (rhou,), E)
    Ubh = (rho, rhou / rho, beta)
    UM = tuple(a[Nq:Nh, :] for a in Ubh)
    UP = [_gather(a, ops.mapP) for a in UM]
    for c, (l, r_) in enumerate(((bd.rhoL, bd.rhoR), (bd.uL, bd.uR), (bd.betaL, bd.betaR))):
        UP[c] = UP[c].copy()
        UP[c][0, 0] = l          # UP[c][1]   (first face point of the first element)
        UP[c][-1, -1] = r_       # UP[c][end]
    Uf = tuple(a[Nq:Nh, :] for a in Uh)
    rhoM, rhouM, EM = Uf
    lam = np.abs(wavespeed_1D(rhoM, rhouM * ops.nxJ, EM))
    LFc = .25 * np.maximum(lam, _gather(lam, ops.mapP))
    fSx = euler_fluxes(UM, tuple(UP))
    flux = [fx * ops.nxJ - LFc * (_gather(u, ops.mapP) - u) for fx, u in zip(fSx, Uf)]
    rhsU = [ops.LIFT @ f for f in flux]
    UF = [np.zeros((Nh, K)) for _ in range(3)]
    logs = (P._log(Ubh[0]), P._log(Ubh[2]))
    for j in range(Nh):          # flux_differencing!, :204-218 (column j, rows i >= j, lower-right block skipped)
        for i in range(j, Nh):
            if i < Nq or j < Nq:
                Fx = euler_fluxes_1D(tuple(a[i] for a in Ubh), tuple(a[j] for a in Ubh),
                                     (logs[0][i], logs[1][i]), (logs[0][j], logs[1][j]))
                for c in range(3):
                    upd = 2 * ops.rxJ * ops.Qrh_skew[i, j] * Fx[c]
                    UF[c][i] += upd
                    UF[c][j] -= upd
    rhsU = [ops.Ph @ uf + r_ for uf, r_ in zip(UF, rhsU)]
    return [-x / ops.J for x in rhsU]


def rhs_viscous(U, ops, bd, pr):  # :270-336
    VU = P.v_ufun(ops.Vq @ U[0], (ops.Vq @ U[1],), ops.Vq @ U[2])
    VU = [ops.Pq @ v for v in VU]
    VUf = [ops.Vf @ v for v in VU]
    VUP = [_gather(a, ops.mapP).copy() for a in VUf]
    for c in range(3):
        VUP[c][0, 0] = bd.vL[c]
        VUP[c][-1, -1] = bd.vR[c]

    def surfx(uP, uf):
        return ops.LIFT @ (.5 * (uP - uf) * ops.nxJ)
    VUx = [(ops.Dr @ v + surfx(vp, vf)) / ops.J for v, vp, vf in zip(VU, VUP, VUf)]
    VUx = [ops.Vq @ v for v in VUx]
    VUq = [ops.Vq @ v for v in VU]
    v2, v4 = VUq[1], VUq[2]
    K22 = -(2 * pr.mu - pr.lam) / v4
    K23 = (2 * pr.mu - pr.lam) * v2 / v4 ** 2
    K33 = -(2 * pr.mu - pr.lam) * v2 ** 2 / v4 ** 3 + pr.kappa / pr.cv / v4 ** 2
    sigma = [np.zeros_like(v2), K22 * VUx[1] + K23 * VUx[2], K23 * VUx[1] + K33 * VUx[2]]
    sigma = [ops.Pq @ s for s in sigma]
    sxf = [ops.Vf @ s for s in sigma]
    sxP = [_gather(a, ops.mapP).copy() for a in sxf]
    for c in range(3):
        sxP[c][0, 0] = sxf[c][0, 0]
        sxP[c][-1, -1] = sxf[c][-1, -1]
    return [(ops.Dr @ s + surfx(sp, sf)) / ops.J for s, sp, sf in zip(sigma, sxP, sxf)]


def rhs_ESDG(U, ops, pr, t):  # :338-353
    bd = boundary_data(t, pr)
    rI = rhs_inviscid(U, ops, bd)
    rV = rhs_viscous(U, ops, bd, pr)
    return [a + b for a, b in zip(rI, rV)]


def ssprk3_dt(ops):  # :381-382
    CN = (ops.N + 1) * (ops.N + 2) / 2
    return 2 / CN / ops.K ** 2


def ssprk3_step(U, ops, pr, t, dt):  # :379-391 (all three stages use the boundary data of time t, as the script does)
    r = rhs_ESDG(U, ops, pr, t)
    W = [u + dt * x for u, x in zip(U, r)]
    r = rhs_ESDG(W, ops, pr, t)
    Z = [w + dt * x for w, x in zip(W, r)]
    W = [3 / 4 * u + 1 / 4 * z for u, z in zip(U, Z)]
    r = rhs_ESDG(W, ops, pr, t)
    Z = [w + dt * x for w, x in zip(W, r)]
    return [1 / 3 * u + 2 / 3 * z for u, z in zip(U, Z)]
