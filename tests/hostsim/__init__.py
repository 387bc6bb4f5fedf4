"""TEST INFRASTRUCTURE: the lane-level CUDA source of maple_b200/csrc compiled for the host and driven through the oracle's
python wrapper, so the golden-vector tests can hold the kernel SOURCE (not a restatement of it) to the reference in a
container without a GPU.  See hostsim.cpp and shim/cuda_runtime.h.  The product never loads this library."""
import ctypes as C
import os
import subprocess

from oracle.oracle import Oracle, _p, declare

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "maple_b200", "csrc")
LIB = os.path.join(HERE, "libhostsim.so")
SOURCES = [os.path.join(HERE, "hostsim.cpp"), os.path.join(HERE, "shim", "cuda_runtime.h")] + [
    os.path.join(CSRC, f) for f in ("glist.cuh", "likelihood.cuh", "search.cuh", "place.cuh")]
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.isfile(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in SOURCES):
        # -ffp-contract=off: no fused multiply-add, like nvcc -fmad=false for the real build
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-I", os.path.join(HERE, "shim"), "-I", CSRC, SOURCES[0], "-o", LIB])
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = declare(C.CDLL(build()))
        for f in (L.hs_append_sitewise, L.hs_append_q4):
            f.restype = C.c_double
            f.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_double]
        _lib = L
    return _lib


class KernelSourceOnHost(Oracle):
    """Same interface as the oracle, computing with dev_append / dev_merge / dev_blen / ... / search_node / place_sample."""

    def __init__(self, model, with_root_tables: bool = False):
        super().__init__(model, with_root_tables)
        self.L = lib()

    def append_variant(self, which: str, P, C_, isTipC, bLen):
        """'sitewise' / 'q4': the per-lane forms the warp scans of the search kernel call (search_fsm.cuh)."""
        a, b = self._one(P), self._one(C_)
        f = {"sitewise": self.L.hs_append_sitewise, "q4": self.L.hs_append_q4}[which]
        return f(self.mp, _p(a.key), _p(a.pay), _p(b.key), _p(b.pay), int(bool(isTipC)), float(bLen))
