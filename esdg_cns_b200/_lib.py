"""ctypes binding of ``libesdg_b200.so`` (C ABI in ``include/esdg_b200.h``).

This is the same marshalling a Julia ``ccall`` shim performs.
There is no CPU fallback: if the CUDA library is missing or no GPU is visible
the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESDG_B200_LIB", os.path.join(_HERE, "libesdg_b200.so"))   # override: kernel-variant experiments only

ESDG_OK, ESDG_EINVAL, ESDG_ECUDA, ESDG_ENCCL, ESDG_ENOMEM, ESDG_ESTATE = range(6)
ESDG_TRI, ESDG_QUAD, ESDG_HEX = 0, 1, 2


class EsdgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"esdg_b200 error {code}: {msg}")
        self.code = code


class DomainError(EsdgError):
    """Mirrors Julia's DomainError from log/sqrt of a negative density/pressure."""


class esdg_desc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("elem", C.c_int32), ("N", C.c_int32), ("Nfields", C.c_int32),
        ("Ns", C.c_int32), ("Nq", C.c_int32), ("Nfq", C.c_int32), ("Nfaces", C.c_int32),
        ("collocated", C.c_int32), ("affine", C.c_int32),
        ("K", C.c_int64), ("n_halo", C.c_int64),
        ("gamma", C.c_double), ("lf_scale", C.c_double), ("volume_scale", C.c_double),
    ]


_lib = None

_PP = C.POINTER(C.c_void_p)

_SIGNATURES = {
    "esdg_create": (C.c_int, [C.POINTER(esdg_desc), C.c_int, _PP]),
    "esdg_destroy": (C.c_int, [C.c_void_p]),
    "esdg_last_error": (C.c_char_p, [C.c_void_p]),
    "esdg_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg_set_operator": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64]),
    "esdg_set_geometry": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64]),
    "esdg_set_maps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "esdg_set_viscosity": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]),
    "esdg_set_wall_bc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]),
    "esdg_set_inflow_outflow_bc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]),
    "esdg_set_boundary_nodes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "esdg_set_structured_hex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "esdg_finalize": (C.c_int, [C.c_void_p]),
    "esdg_rhs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "esdg_upload_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg_download_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg_lsrk45_step": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_void_p]),
    "esdg_lsrk45_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int]),
    "esdg_dopri45_init": (C.c_int, [C.c_void_p]),
    "esdg_dopri45_step": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double,
                                    C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p]),
    "esdg_sync": (C.c_int, [C.c_void_p]),
    "esdg_set_send_list": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "esdg_compute_traces": (C.c_int, [C.c_void_p]),
    "esdg_pack_send": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg_halo_ptr": (C.c_int, [C.c_void_p, _PP, C.POINTER(C.c_int64)]),
    "esdg_unpack_halo": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg_stage_update": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]),
    "esdg_cns_sigma_pass": (C.c_int, [C.c_void_p, C.c_int]),
    "esdg_pack_send_sigma": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg_unpack_halo_sigma": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg_trace_planes": (C.c_int, [C.c_void_p]),
    "esdg_comm_unique_id": (C.c_int, [C.c_void_p]),
    "esdg_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esdg_comm_destroy": (C.c_int, [C.c_void_p]),
    "esdg1d_create": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int] + [C.c_void_p] * 8 + [C.c_double] * 6 + [_PP]),
    "esdg1d_destroy": (C.c_int, [C.c_void_p]),
    "esdg1d_last_error": (C.c_char_p, [C.c_void_p]),
    "esdg1d_upload_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg1d_download_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg1d_rhs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "esdg1d_ssprk3_step": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "esdg1d_kernel_launches": (C.c_int64, [C.c_void_p]),
    "esdg_kernel_path": (C.c_char_p, [C.c_void_p]),
    "esdg_read_rhstest": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg_read_rhstest_visc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esdg_kernel_launches": (C.c_int64, [C.c_void_p]),
    "esdg_kernel_time_ms": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "esdg_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "esdg_debug_log_exp": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "esdg_debug_pair_flux": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]),
    "esdg_measure_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "esdg_version": (C.c_char_p, []),
}


def load():
    """Load the shared library (raises if it was not built: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C esdg_cns_b200/csrc` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if "ESDG_B200_LIB" not in os.environ:      # (experiment builds carry other flags / hashes)
            from .roofline import kernel_source_hash
            built, now = lib.esdg_version().decode().rsplit("src=", 1)[-1], kernel_source_hash()
            if built != now:
                raise OSError(f"{LIB_PATH} was built from other sources (library {built}, tree {now}): "
                              "run `make -C esdg_cns_b200/csrc` -- a stale prebuilt library is never used")
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(ctx, code):
    if code == ESDG_OK:
        return
    msg = load().esdg_last_error(ctx)
    msg = msg.decode() if msg else ""
    if code == ESDG_ESTATE:
        raise DomainError(code, msg)
    raise EsdgError(code, msg)


def f64(a):
    """Column-major float64 view/copy (what Julia hands over)."""
    return np.asfortranarray(a, dtype=np.float64)


def ptr_array(arrays):
    """C array of pointers to the given numpy arrays (Julia: Vector{Ptr{Float64}})."""
    arr = (C.c_void_p * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data
    return arr
