"""The C-ABI library builds, loads and exports every symbol include/maple_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "maple_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(maple_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from maple_b200.build import build_extension
    lib = ctypes.CDLL(build_extension())
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.maple_version() >= 100


def test_ctypes_signatures_cover_header():
    from maple_b200 import capi
    assert set(declared_symbols()) == set(capi.SIGNATURES)


def test_no_gpu_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        return
    import numpy as np
    from maple_b200 import capi
    lib = capi.load()
    ctx = ctypes.c_void_p()
    pi = (ctypes.c_double * 4)(0.25, 0.25, 0.25, 0.25)
    rc = lib.maple_ctx_create(ctypes.byref(ctx), 0, 100, pi, 0)
    assert rc == -4  # MAPLE_E_NOGPU: no CPU fallback
    assert b"no CPU fallback" in lib.maple_last_error(None)
