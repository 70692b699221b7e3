"""Shared test helpers (may import the oracle: tests are oracle territory)."""
import contextlib

import numpy as np

from oracle import physics as P

NORTH_STAR_TOL = 1e-12   # BASELINE.json north_star: "within 1e-12 relative in FP64"


def relerr(out, ref):
    """max|out-ref| / max|ref| over all fields."""
    scale = max(np.abs(np.asarray(r)).max() for r in ref)
    return max(np.abs(np.asarray(a) - np.asarray(b)).max() for a, b in zip(out, ref)) / scale


@contextlib.contextmanager
def ulp_noisy_log(seed=1234):
    """Move every log/exp/pow result of the oracle by a random -1/0/+1 ulp."""
    rng = np.random.default_rng(seed)

    def jitter(y):
        r = rng.integers(-1, 2, size=np.shape(y))
        return np.where(r > 0, np.nextafter(y, np.inf), np.where(r < 0, np.nextafter(y, -np.inf), y))

    old = (P.LOG_IMPL, P.EXP_IMPL, P.POW_IMPL)
    P.LOG_IMPL = lambda x: jitter(np.log(x))
    P.EXP_IMPL = lambda x: jitter(np.exp(x))
    P.POW_IMPL = lambda x, y: jitter(np.power(x, y))
    try:
        yield
    finally:
        P.LOG_IMPL, P.EXP_IMPL, P.POW_IMPL = old


def noise_floor(fun, ref, trials=3):
    """Relative change of ``fun()`` (an oracle evaluation) under 1-ulp log/exp/pow noise.

    The reference's logmean (logmean.jl:26) divides by ``logL-logR``: two *faithful* log
    implementations (<= 1 ulp) already move the rhs by this much, so no implementation --
    Julia included -- can agree with another to better than this on the same input.
    """
    worst = 0.0
    for t in range(trials):
        with ulp_noisy_log(seed=100 + t):
            worst = max(worst, relerr(fun(), ref))
    return worst


def parity_tol(fun, ref):
    """1e-12 (north star) wherever the formula is conditioned to it, else 4x the oracle's own
    1-ulp-log noise floor."""
    return max(NORTH_STAR_TOL, 4.0 * noise_floor(fun, ref))
