"""Host mirror of the 1-D compressible Navier-Stokes path (include/esdg_b200.h: esdg1d_*), the call shape of
examples/CompressibleNS/dg1D_CNS_modalESDG.jl: ``rhs_ESDG(U, ..., t)`` and the SSPRK(3,3) loop body.

The caller supplies the script's operators (``ops``: Vq, Pq, Vf, LIFT, Dr, VhP, Ph, Qrh_skew, J, rxJ, N, K) and, per
call, the exact-solution data the script imposes on the two domain ends (``bd``: rhoL, uL, betaL, rhoR, uR, betaR,
vL, vR).  No CPU fallback: the CUDA library does all the arithmetic."""
import ctypes as C

import numpy as np

from . import _lib as L


class Solver1D:
    def __init__(self, ops, mu, lam, kappa_over_cv, gamma=1.4, device=0):
        self._lib = L.load()
        self.Np, self.K = int(ops.N) + 1, int(ops.K)
        mats = [L.f64(getattr(ops, n)) for n in ("Vq", "Pq", "Vf", "LIFT", "Dr", "VhP", "Ph", "Qrh_skew")]
        self._ctx = C.c_void_p()
        rc = self._lib.esdg1d_create(int(device), int(ops.N), self.K, int(ops.Vq.shape[0]), *[m.ctypes.data for m in mats],
                                     float(ops.J), float(ops.rxJ), float(gamma), float(mu), float(lam), float(kappa_over_cv),
                                     C.byref(self._ctx))
        self._ck(rc)

    def _ck(self, rc):
        if rc:
            msg = self._lib.esdg1d_last_error(self._ctx)
            raise L.EsdgError(rc, msg.decode() if msg else "")

    @staticmethod
    def _bc(bd):
        return np.ascontiguousarray([bd.rhoL, bd.uL, bd.betaL, bd.rhoR, bd.uR, bd.betaR, *bd.vL, *bd.vR], dtype=np.float64)

    def upload(self, U):
        q = [L.f64(u) for u in U]
        self._ck(self._lib.esdg1d_upload_state(self._ctx, L.ptr_array(q)))

    def download(self):
        out = [np.empty((self.Np, self.K), order="F") for _ in range(3)]
        self._ck(self._lib.esdg1d_download_state(self._ctx, L.ptr_array(out)))
        return out

    def rhs(self, bd):
        out = [np.empty((self.Np, self.K), order="F") for _ in range(3)]
        b = self._bc(bd)
        self._ck(self._lib.esdg1d_rhs(self._ctx, b.ctypes.data, L.ptr_array(out)))
        return out

    def ssprk3_step(self, dt, bd):
        b = self._bc(bd)
        self._ck(self._lib.esdg1d_ssprk3_step(self._ctx, float(dt), b.ctypes.data))

    def kernel_launches(self):
        return int(self._lib.esdg1d_kernel_launches(self._ctx))

    def close(self):
        if self._ctx:
            self._lib.esdg1d_destroy(self._ctx)
            self._ctx = C.c_void_p()
