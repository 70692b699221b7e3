"""ctypes wrapper of oracle/c/libesdg_oracle.so (scalar C restatement; checker + CPU baseline).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "libesdg_oracle.so")
_lib = None


class oracle_desc(C.Structure):
    _fields_ = [("dim", C.c_int), ("nf", C.c_int), ("Nq", C.c_int), ("Nfq", C.c_int), ("Nh", C.c_int),
                ("K", C.c_int64)] + \
               [(n, C.c_void_p) for n in ("ef_ptr", "ef_col", "ef_val", "ph_ptr", "ph_col", "ph_val",
                                          "lf_ptr", "lf_col", "lf_val", "q_ptr", "q_col", "q_val",
                                          "G", "J", "nxJ", "nyJ", "nzJ", "sJ", "wq", "mapP")] + \
               [("lf_scale", C.c_double), ("volume_scale", C.c_double)]


def load(build=True):
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) and build:
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")])
        lib = C.CDLL(_SO)
        lib.esdg_oracle_threads.restype = C.c_int
        lib.esdg_oracle_set_threads.argtypes = [C.c_int]
        lib.esdg_oracle_set_threads.restype = None
        lib.esdg_oracle_work_doubles.restype = C.c_int64
        lib.esdg_oracle_work_doubles.argtypes = [C.POINTER(oracle_desc)]
        lib.esdg_oracle_rhs.restype = C.c_double
        lib.esdg_oracle_rhs.argtypes = [C.POINTER(oracle_desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.esdg_oracle_lsrk45.restype = C.c_double
        lib.esdg_oracle_lsrk45.argtypes = [C.POINTER(oracle_desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_double, C.c_int]
        _lib = lib
    return _lib


def _csr(A, tol=1e-12):
    A = np.asarray(A)
    ptr, col, val = [0], [], []
    for i in range(A.shape[0]):
        js = np.nonzero(np.abs(A[i]) > tol)[0]
        col.extend(js.tolist())
        val.extend(A[i, js].tolist())
        ptr.append(len(col))
    return (np.array(ptr, dtype=np.int32), np.array(col, dtype=np.int32), np.array(val, dtype=np.float64))


class COracle:
    """Collocated quad/hex rhs + LSRK45 in scalar C, from oracle-format (md, ops)."""

    def __init__(self, md, ops, lf_scale=None):
        self.lib = load()
        dim = 3 if ops.elem == "hex" else 2
        nf = dim + 2
        Nq = ops.wq.size
        Nfq = ops.Lf.shape[1]
        K = int(md.K)
        self.nf, self.Nq, self.K = nf, Nq, K
        keep = []
        d = oracle_desc(dim=dim, nf=nf, Nq=Nq, Nfq=Nfq, Nh=Nq + Nfq, K=K)

        def put(names, arrs):
            for n, a in zip(names, arrs):
                a = np.ascontiguousarray(a)
                keep.append(a)
                setattr(d, n, a.ctypes.data)

        put(("ef_ptr", "ef_col", "ef_val"), _csr(ops.Ef))
        put(("ph_ptr", "ph_col", "ph_val"), _csr(ops.Ph))
        put(("lf_ptr", "lf_col", "lf_val"), _csr(ops.Lf))
        # union pattern in the reference's first-seen order (Qrsids / Qnzids)
        Qs = [np.where(np.abs(A) > 1e-12, A, 0.0) for A in ops.Qskew]
        ptr, col, val = [0], [], []
        for i in range(Nq + Nfq):
            seen = []
            for A in Qs:
                for j in np.nonzero(A[i])[0]:
                    if j not in seen:
                        seen.append(int(j))
            for j in seen:
                col.append(j)
                val.extend([Qs[0][i, j], Qs[1][i, j], Qs[2][i, j] if dim == 3 else 0.0])
            ptr.append(len(col))
        put(("q_ptr", "q_col", "q_val"), (np.array(ptr, np.int32), np.array(col, np.int32), np.array(val, np.float64)))
        g = np.zeros((K, 9))
        names = ["rxJ", "sxJ", "txJ", "ryJ", "syJ", "tyJ", "rzJ", "szJ", "tzJ"]
        for q, n in enumerate(names):
            if hasattr(md, n):
                g[:, q] = np.atleast_2d(getattr(md, n))[0, :]
        Nfp = Nfq // ops.Nfaces

        def facepts(a):
            a = np.atleast_2d(a)
            if a.shape[0] == ops.Nfaces:
                a = np.repeat(a, Nfp, axis=0)
            return a.flatten(order="F")

        put(("G", "J"), (g, np.atleast_2d(md.J)[0, :].copy()))
        put(("nxJ", "nyJ", "sJ"), (facepts(md.nxJ), facepts(md.nyJ), facepts(md.sJ)))
        put(("nzJ",), (facepts(md.nzJ) if dim == 3 else np.zeros(1),))
        put(("wq", "mapP"), (np.asarray(ops.wq, dtype=np.float64), np.asarray(md.mapP, dtype=np.int64).flatten(order="F")))
        d.lf_scale = float(ops.lf_scale if lf_scale is None else lf_scale)
        d.volume_scale = float(ops.volume_scale)
        self.d, self._keep = d, keep
        self.work = np.empty(self.lib.esdg_oracle_work_doubles(C.byref(d)))
        self.threads = self.lib.esdg_oracle_threads()

    def _pack(self, Q):
        return np.ascontiguousarray(np.stack([np.asarray(q).flatten(order="F") for q in Q]))

    def _unpack(self, a):
        return [a[f].reshape((self.Nq, self.K), order="F") for f in range(self.nf)]

    def rhs(self, Q, compute_rhstest=False):
        q = self._pack(Q)
        out = np.empty_like(q)
        rt = self.lib.esdg_oracle_rhs(C.byref(self.d), q.ctypes.data, out.ctypes.data, self.work.ctypes.data,
                                      int(compute_rhstest))
        return self._unpack(out), rt

    def lsrk45(self, Q, dt, nsteps=1):
        q = self._pack(Q)
        res = np.zeros_like(q)
        rhs = np.empty_like(q)
        rt = self.lib.esdg_oracle_lsrk45(C.byref(self.d), q.ctypes.data, res.ctypes.data, rhs.ctypes.data,
                                         self.work.ctypes.data, float(dt), int(nsteps))
        return self._unpack(q), rt
