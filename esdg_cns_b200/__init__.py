"""esdg_cns_b200 -- B200-native rhs + LSRK hot path of yiminllin/ESDG-CNS.

``api``       host mirror of the reference's ``rhs`` / LSRK call shape over the C ABI
``setup_dg``  tensor-product setup of uniform quad/hex meshes (not on the hot path)
``_lib``      ctypes binding of libesdg_b200.so (include/esdg_b200.h)
"""
from . import _lib, api, setup_dg  # noqa: F401
from .api import Solver, rhs, lsrk45_steps, euler_fluxes, rk45_coeffs  # noqa: F401
