"""Extended-precision arbiter for round-off questions (oracle; TEST INFRASTRUCTURE ONLY).

``rhs_collocated_ld`` evaluates the SAME restatement of the reference rhs (oracle/solvers.py:rhs_collocated,
dg2D_euler_quad.jl:141-194 / dg3D_euler_hex.jl:167-222) on the SAME float64 inputs, but carries every
intermediate in x87 80-bit ``numpy.longdouble`` (eps 1.1e-19).  Where two float64 evaluations of the
reference formulas (numpy, C, CUDA, Julia) differ only through rounding -- low-Mach states on fine meshes,
where the rhs is a small difference of O(p) fluxes -- this value tells which one is closer to what the
formulas define.  Branch decisions (``abs(f) < 1e-4`` in logmean.jl:17) are taken on the extended values;
a pair within an ulp of the switch may therefore sit in the other branch than in a float64 run.
"""
import copy

import numpy as np

from . import solvers as SV

LD = np.longdouble


def _ld(a):
    return np.asarray(a, dtype=LD)


def to_longdouble(md, ops):
    """Deep copies of (md, ops) with every floating array widened (values unchanged)."""
    md2, ops2 = copy.copy(md), copy.copy(ops)
    for obj in (md2, ops2):
        for k, v in list(vars(obj).items()):
            if isinstance(v, np.ndarray) and v.dtype == np.float64:
                setattr(obj, k, _ld(v))
            elif isinstance(v, (tuple, list)) and len(v) and all(isinstance(x, np.ndarray) and x.dtype == np.float64 for x in v):
                setattr(obj, k, type(v)(_ld(x) for x in v))
    return md2, ops2


def rhs_collocated_ld(Q, md, ops, rd, lf_scale, compute_rhstest=False):
    md2, ops2 = to_longdouble(md, ops)
    out, rt = SV.rhs_collocated([_ld(q) for q in Q], md2, ops2, rd, lf_scale, compute_rhstest)
    return out, rt
