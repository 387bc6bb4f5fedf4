"""TEST INFRASTRUCTURE: the lane-level CUDA source of maple_b200/csrc compiled for the host and driven through the oracle's
python wrapper, so the golden-vector tests can hold the kernel SOURCE (not a restatement of it) to the reference in a
container without a GPU.  See hostsim.cpp and shim/cuda_runtime.h.  The product never loads this library."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle.oracle import PLACE_RESULT_DTYPE, SEARCH_RESULT_DTYPE, Oracle, OrPlaceParams, OrSearchParams, _p, declare

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "maple_b200", "csrc")
LIB = os.path.join(HERE, "libhostsim.so")
SOURCES = [os.path.join(HERE, "hostsim.cpp"), os.path.join(HERE, "shim", "cuda_runtime.h")] + [
    os.path.join(CSRC, f) for f in ("glist.cuh", "likelihood.cuh", "search.cuh", "search_fsm.cuh", "place.cuh", "place_scan.cuh")]
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
        L.hs_search_batch_fsm.restype = None
        L.hs_search_batch_fsm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.hs_place_batch_scan.restype = None
        L.hs_place_batch_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]
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

    def place_batch_scan(self, tree: dict, lists, params: dict, samples, scratch_keys: int = 4096, mat: int = 0):
        """Placement variant 1 (one sample per warp, place_scan.cuh: place_sample_warp), or with mat=True variant 2
        (place_sample_warp_mat: MAT trees covered) and with mat=2 variant 3 (parallel window replay), lanes emulated in turn."""
        t, keep = self._tree_struct(tree, lists)
        pp = OrPlaceParams()
        for k, v in params.items():
            setattr(pp, k, v)
        out = np.zeros(len(samples), dtype=PLACE_RESULT_DTYPE)
        self.L.hs_place_batch_scan(self.mp, C.addressof(t), C.addressof(pp), len(samples), _p(samples.key), _p(samples.pay),
                                   _p(samples.key_start), _p(samples.pay_start), _p(samples.nkeys), int(scratch_keys), _p(lists.npay), int(mat), _p(out))
        return out

    def search_batch_fsm(self, tree: dict, lists, params: dict, nodes, scratch_keys: int = 8192):
        """The default search kernel's per-lane state machine (fsm_step / fsm_finish) with the warp scans off, one lane."""
        t, keep = self._tree_struct(tree, lists)
        sp = OrSearchParams()
        for k, v in params.items():
            setattr(sp, k, v)
        nodes = np.ascontiguousarray(nodes, np.int32)
        out = np.zeros(len(nodes), dtype=SEARCH_RESULT_DTYPE)
        self.L.hs_search_batch_fsm(self.mp, C.addressof(t), C.addressof(sp), len(nodes), _p(nodes), int(scratch_keys), _p(out))
        return out


# ---- the warp-level code with its 32 lanes emulated (hostwarp.cpp, shim_warp/cuda_runtime.h)
WARP_LIB = os.path.join(HERE, "libhostwarp.so")
WARP_SOURCES = [os.path.join(HERE, "hostwarp.cpp"), os.path.join(HERE, "shim_warp", "cuda_runtime.h")] + [
    os.path.join(CSRC, f) for f in ("glist.cuh", "likelihood.cuh", "search.cuh", "search_fsm.cuh", "scan2.cuh")]
_warp_lib = None


def build_warp(force: bool = False) -> str:
    if force or not os.path.isfile(WARP_LIB) or any(os.path.getmtime(s) > os.path.getmtime(WARP_LIB) for s in WARP_SOURCES):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-I", os.path.join(HERE, "shim_warp"), "-I", CSRC, WARP_SOURCES[0], "-o", WARP_LIB])
    return WARP_LIB


def warp_lib():
    global _warp_lib
    if _warp_lib is None:
        L = C.CDLL(build_warp())
        L.hw_search_batch.restype = C.c_int
        L.hw_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.hw_scan_append.restype = C.c_double
        L.hw_scan_append.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]
        _warp_lib = L
    return _warp_lib


class WarpKernelOnHost(Oracle):
    """The body of k_spr_search_fsm (state machines + warp-cooperative subtree scans) run on the host, lanes emulated."""

    def __init__(self, model):
        super().__init__(model)
        self.W = warp_lib()

    def search_batch_warp(self, tree: dict, lists, params: dict, nodes, scan_form: int = 2, scan_min_size: int = 8, lanes_per_warp: int = 3,
                          pool_bytes: int = 10240, scan_flags: int = 0, scratch_keys: int = 8192, stats=None, big_slots: int = 0,
                          owner_warps: int = 1, server_warps: int = 0, dense_rows: int = 0, eval_slice: int = 1024, head_end: int = 0):
        """scan_form: 0 no scans, 1 first form (search_fsm.cuh: warp_scan_job), 2 second form (scan2.cuh: warp_scan_job2).
        server_warps > 0 (scan form 2): the scan service -- owner_warps warps own the searches and post their subtree scans, server_warps
        warps only serve them (scan2.cuh: ScanQueue, scan_server_loop), all emulated together.  dense_rows > 0 (scan form 2): the
        dense scoring pass first (scan2.cuh: DenseScores) for up to that many searches; their subtree scans read the scores.
        eval_slice: scratch entries per lane for the warp-wide evaluation of queued phase-2 entries (search_fsm.cuh: warp_eval_queue;
        the device gives 1024), 0 = the owning lane evaluates its queue itself."""
        t, keep = self._tree_struct(tree, lists)
        sp = OrSearchParams()
        for k, v in params.items():
            setattr(sp, k, v)
        nodes = np.ascontiguousarray(nodes, np.int32)
        out = np.zeros(len(nodes), dtype=SEARCH_RESULT_DTYPE)
        npay = np.ascontiguousarray(lists.npay, np.int32)
        rc = self.W.hw_search_batch(self.mp, C.addressof(t), C.addressof(sp), len(nodes), _p(nodes), int(scratch_keys), _p(npay), int(scan_form),
                                    int(scan_min_size), int(lanes_per_warp), int(pool_bytes), int(scan_flags), int(big_slots), int(owner_warps), int(server_warps), int(dense_rows), int(eval_slice), int(head_end), _p(out), _p(stats))
        if rc != 0:
            raise RuntimeError("hw_search_batch: scan form %d not available for this tree" % scan_form)
        return out

    def scan_append(self, P, C_, isTipC, bLen, convert_slow=False):
        """appendProbNode through the scan-format copies of both lists (scan2.cuh).  convert_slow: with the candidate side's O
        entries below the 0.02 shortcut given their factor against a plain reference run first, as a scan job does to its
        staged copies (bit 27 of the scan format)."""
        a, b = self._one(P), self._one(C_)
        return self.W.hw_scan_append(self.mp, _p(a.key), _p(a.pay), int(a.nkeys[0]), _p(b.key), _p(b.pay), int(bool(isTipC)), float(bLen),
                                     int(bool(convert_slow)))
