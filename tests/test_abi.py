"""The C-ABI library loads and exports every symbol include/esdg_b200.h declares (no compute calls:
this runs without a GPU)."""
import os
import re
import subprocess

from esdg_cns_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "esdg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(esdg(?:1d)?_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert "esdg_rhs" in names and "esdg_lsrk45_step" in names
    out = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (esdg(?:1d)?_[a-z0-9_]+)", out))
    for n in names:
        assert n in exported, f"{n} declared in esdg_b200.h but not exported"
        assert hasattr(lib, n)
    assert sorted(_lib.exported_symbols()) == names, "ctypes signature table out of sync with the header"


def test_version_and_descriptor_layout():
    import ctypes
    lib = _lib.load()
    assert b"sm_100a" in lib.esdg_version()
    assert ctypes.sizeof(_lib.esdg_desc) == 10 * 4 + 2 * 8 + 3 * 8


def test_no_cpu_fallback_create_fails_without_gpu():
    """Without a CUDA device esdg_create must fail loudly (ESDG_ECUDA), never compute on the host."""
    import ctypes
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    d = _lib.esdg_desc(dim=3, elem=2, N=1, Nfields=5, Ns=8, Nq=8, Nfq=24, Nfaces=6, collocated=1, affine=1, K=1,
                       n_halo=0, gamma=1.4, lf_scale=.25, volume_scale=1.0)
    ctx = ctypes.c_void_p()
    rc = lib.esdg_create(ctypes.byref(d), 0, ctypes.byref(ctx))
    assert rc == _lib.ESDG_ECUDA
    assert b"no CUDA device" in lib.esdg_last_error(ctx)
    lib.esdg_destroy(ctx)
