"""Pointwise Euler functions (oracle; test infrastructure only).

Restates, operation for operation, the *local* physics module the example
scripts import: ``examples/EntropyStableEuler/{logmean,euler_fluxes,
euler_variables,EntropyStableEuler}.jl``.  All functions broadcast over numpy
arrays exactly like the Julia ``@.`` versions.
"""
import numpy as np

GAMMA = 1.4  # examples/EntropyStableEuler/EntropyStableEuler.jl:9

# Every ``log`` of the path goes through this hook so that tests can measure the
# formula's own conditioning: ``noise_floor`` (tests/helpers.py) re-evaluates the
# oracle with each log result moved by a random -1/0/+1 ulp, which is the spread
# between two faithful libm implementations (Julia's, glibc's, CUDA's).
LOG_IMPL = np.log
EXP_IMPL = np.exp
POW_IMPL = np.power


def _log(x):
    return LOG_IMPL(x)


def _exp(x):
    return EXP_IMPL(x)


def _pow(x, y):
    return POW_IMPL(x, y)


# --------------------------------------------------------------------------
# logmean.jl:14-28
def logmean(aL, aR, logL=None, logR=None):
    aL = np.asarray(aL)
    aR = np.asarray(aR)
    if aL.dtype != np.longdouble:      # extended-precision arbiter (oracle/extended.py) keeps its dtype
        aL = aL.astype(np.float64)
        aR = aR.astype(np.float64)
    if logL is None:
        logL = _log(aL)  # logmean.jl:5-7
        logR = _log(aR)
    da = aR - aL
    aavg = .5 * (aR + aL)
    f = da / aavg
    v = f ** 2
    taylor = aavg * (1 + v * (-.2 - v * (.0512 - v * 0.026038857142857)))
    with np.errstate(divide="ignore", invalid="ignore"):
        exact = -da / (logL - logR)
    return np.where(np.abs(f) < 1e-4, taylor, exact)


# --------------------------------------------------------------------------
# euler_fluxes.jl:23-48 (2D) / :51-89 (3D).  U = (rho, u, v[, w], beta).
def euler_fluxes_2d(UL, UR, logL=None, logR=None):
    rhoL, uL, vL, betaL = UL
    rhoR, uR, vR, betaR = UR
    if logL is None:  # euler_fluxes.jl:9-15
        logL = (_log(rhoL), _log(betaL))
        logR = (_log(rhoR), _log(betaR))
    rholog = logmean(rhoL, rhoR, logL[0], logR[0])
    betalog = logmean(betaL, betaR, logL[1], logR[1])
    rhoavg = .5 * (rhoL + rhoR)
    uavg = .5 * (uL + uR)
    vavg = .5 * (vL + vR)
    unorm = uL * uR + vL * vR
    pa = rhoavg / (betaL + betaR)
    f4aux = rholog / (2 * (GAMMA - 1) * betalog) + pa + .5 * rholog * unorm
    FxS1 = rholog * uavg
    FxS2 = FxS1 * uavg + pa
    FxS3 = FxS1 * vavg
    FxS4 = f4aux * uavg
    FyS1 = rholog * vavg
    FyS2 = FxS3
    FyS3 = FyS1 * vavg + pa
    FyS4 = f4aux * vavg
    return (FxS1, FxS2, FxS3, FxS4), (FyS1, FyS2, FyS3, FyS4)


def euler_fluxes_3d(UL, UR, logL=None, logR=None):
    rhoL, uL, vL, wL, betaL = UL
    rhoR, uR, vR, wR, betaR = UR
    if logL is None:
        logL = (_log(rhoL), _log(betaL))
        logR = (_log(rhoR), _log(betaR))
    rholog = logmean(rhoL, rhoR, logL[0], logR[0])
    betalog = logmean(betaL, betaR, logL[1], logR[1])
    rhoavg = .5 * (rhoL + rhoR)
    uavg = .5 * (uL + uR)
    vavg = .5 * (vL + vR)
    wavg = .5 * (wL + wR)
    unorm = uL * uR + vL * vR + wL * wR
    pa = rhoavg / (betaL + betaR)
    Ep = rholog / (2 * (GAMMA - 1) * betalog) + pa + .5 * rholog * unorm
    FxS1 = rholog * uavg
    FxS2 = FxS1 * uavg + pa
    FxS3 = FxS1 * vavg
    FxS4 = FxS1 * wavg
    FxS5 = Ep * uavg
    FyS1 = rholog * vavg
    FyS2 = FxS3
    FyS3 = FyS1 * vavg + pa
    FyS4 = FyS1 * wavg
    FyS5 = Ep * vavg
    FzS1 = rholog * wavg
    FzS2 = FxS4
    FzS3 = FyS4
    FzS4 = FzS1 * wavg + pa
    FzS5 = Ep * wavg
    return ((FxS1, FxS2, FxS3, FxS4, FxS5),
            (FyS1, FyS2, FyS3, FyS4, FyS5),
            (FzS1, FzS2, FzS3, FzS4, FzS5))


def euler_fluxes(UL, UR, logL=None, logR=None):
    """Dispatch on the number of fields (euler_fluxes.jl:17-20)."""
    if len(UL) == 4:
        return euler_fluxes_2d(UL, UR, logL, logR)
    return euler_fluxes_3d(UL, UR, logL, logR)


# --------------------------------------------------------------------------
# euler_variables.jl
def _unorm(U):  # :12
    s = U[0] ** 2
    for x in U[1:]:
        s = s + x ** 2
    return s


def primitive_to_conservative(rho, U, p):  # :15-20
    rhoU = tuple(rho * x for x in U)
    unorm = _unorm(U)
    E = p / (GAMMA - 1) + .5 * rho * unorm
    return (rho,) + rhoU + (E,)


def pfun(rho, rhoU, E):  # :42-48
    rhounorm = _unorm(rhoU) / rho
    return (GAMMA - 1) * (E - .5 * rhounorm)


def betafun(rho, rhoU, E):  # :30-39
    p = pfun(rho, rhoU, E)
    return rho / (2 * p)


def wavespeed(rho, rhou, E):  # :7-10  (sic: sqrt of |u_n|; H4)
    cvel = np.sqrt(GAMMA * pfun(rho, (rhou,), E) / rho)
    return np.sqrt(np.abs(rhou / rho)) + cvel


def rhoefun(rho, rhoU, E):  # :59-62
    rhoUnorm = _unorm(rhoU)
    return E - .5 * rhoUnorm / rho


def sfun(rho, rhoU, E):  # :65-68
    rhoe = rhoefun(rho, rhoU, E)
    return _log((GAMMA - 1) * rhoe / _pow(rho, GAMMA))


def Sfun(rho, rhoU, E):  # :71-76
    return -rho * sfun(rho, rhoU, E)


def v_ufun(rho, rhoU, E):  # :79-86
    rhoe = rhoefun(rho, rhoU, E)
    sU = sfun(rho, rhoU, E)
    v1 = (-E + rhoe * (GAMMA + 1 - sU)) / rhoe
    vU = tuple(x / rhoe for x in rhoU)
    vE = (-rho) / rhoe
    return (v1,) + vU + (vE,)


def s_vfun(v1, vU, vE):  # :95-98
    vUnorm = _unorm(vU)
    return GAMMA - v1 + vUnorm / (2 * vE)


def rhoe_vfun(v1, vU, vE):  # :101-104
    s = s_vfun(v1, vU, vE)
    return _pow((GAMMA - 1) / _pow(-vE, GAMMA), 1 / (GAMMA - 1)) * _exp(-s / (GAMMA - 1))


def u_vfun(v1, vU, vE):  # :107-113
    rhoeV = rhoe_vfun(v1, vU, vE)
    vUnorm = _unorm(vU)
    rho = rhoeV * (-vE)
    rhoU = tuple(rhoeV * x for x in vU)
    E = rhoeV * (1 - vUnorm / (2 * vE))
    return (rho,) + rhoU + (E,)


def v_of_u(Q):
    """Tuple-in / tuple-out wrapper (the ``v_ufun(Q...)`` call shape)."""
    return v_ufun(Q[0], tuple(Q[1:-1]), Q[-1])


def u_of_v(V):
    return u_vfun(V[0], tuple(V[1:-1]), V[-1])


def beta_of_u(Q):
    return betafun(Q[0], tuple(Q[1:-1]), Q[-1])


# --------------------------------------------------------------------------
# EntropyStableEuler.jl:21-35
def vortex(x, y, t, gamma=GAMMA):
    x0 = 5
    y0 = 0
    beta = 5
    r2 = (x - x0 - t) ** 2 + (y - y0) ** 2
    u = 1 - beta * np.exp(1 - r2) * (y - y0) / (2 * np.pi)
    v = beta * np.exp(1 - r2) * (x - x0 - t) / (2 * np.pi)
    rho = 1 - (1 / (8 * gamma * np.pi ** 2)) * (gamma - 1) / 2 * (beta * np.exp(1 - r2)) ** 2
    rho = rho ** (1 / (gamma - 1))
    p = rho ** gamma
    return rho, u, v, p


def taylor_green(x, y, z, mach=0.1):
    """Project-defined IC for config 5 (not in the reference)."""
    rho = np.ones_like(x)
    u = np.sin(x) * np.cos(y) * np.cos(z)
    v = -np.cos(x) * np.sin(y) * np.cos(z)
    w = np.zeros_like(x)
    p = 1.0 / (GAMMA * mach ** 2) + (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2) / 16
    return rho, u, v, w, p
