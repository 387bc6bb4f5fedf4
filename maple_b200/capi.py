"""ctypes binding of include/maple_b200.h.  There is no CPU fallback: if the CUDA library is not
built, or no GPU is present, the first call raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAPLE_B200_LIB") or os.path.join(HERE, "libmaple_b200.so")

MAPLE_F_USING_ERROR_RATE, MAPLE_F_ERROR_SITE_SPECIFIC, MAPLE_F_RATE_VARIATION = 1, 2, 4
DEFAULT_PLACE_VARIANT = 3  # maple_place_batch: one sample per warp, MAT trees covered, parallel window replay
MAPLE_MERGE_UPDOWN, MAPLE_MERGE_RETURN_LK = 1, 2

_P, _I64, _I32, _D = C.c_void_p, C.c_int64, C.c_int32, C.c_double

SIGNATURES = {
    "maple_version": (C.c_int, []),
    "maple_last_error": (C.c_char_p, [_P]),
    "maple_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int, _I32, C.POINTER(_D), _I32]),
    "maple_ctx_destroy": (C.c_int, [_P]),
    "maple_ctx_set_model": (C.c_int, [_P, C.POINTER(_D), _P, _D, _P, _P, _P, _D]),
    "maple_ctx_set_thresholds": (C.c_int, [_P, _D, _D, _D, _D]),
    "maple_lists_bind": (C.c_int, [_P, _P, _P, _P, _P, _I64]),
    "maple_append_prob_batch": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, _P]),
    "maple_append_prob_batch_host": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P]),
    "maple_merge_batch": (C.c_int, [_P, _I64] + [_P] * 17 + [_I32, _P]),
    "maple_blen_batch": (C.c_int, [_P, _I64] + [_P] * 7 + [_P]),
    "maple_vectors_differ_batch": (C.c_int, [_P, _I64, _P, _P, _P, _P]),
    "maple_root_vector_batch": (C.c_int, [_P, _I64] + [_P] * 9 + [_I32, _P]),
    "maple_ctx_set_root_tables": (C.c_int, [_P, _P, _P]),
    "maple_prob_root_batch": (C.c_int, [_P, _I64, _P, _P, _P]),
    "maple_pass_branch_batch": (C.c_int, [_P, _I64] + [_P] * 11 + [_P]),
    "maple_lists_copy": (C.c_int, [_P, _I64] + [_P] * 10 + [_P]),
    "maple_tree_bind": (C.c_int, [_P, _I32, _I32] + [_P] * 9),
    "maple_spr_search_batch": (C.c_int, [_P, _P, _I64, _P, _P, _I32, _I32, _P, _P]),
    "maple_place_batch": (C.c_int, [_P, _P, _I64, _P, _P, _I32, _P]),
    "maple_ctx_set_place_variant": (C.c_int, [_P, _I32]),
    "maple_ctx_set_search_variant": (C.c_int, [_P, _I32]),
    "maple_ctx_set_scan_min_size": (C.c_int, [_P, _I32]),
    "maple_ctx_set_scan_service": (C.c_int, [_P, _I32]),
    "maple_ctx_set_lanes_per_warp": (C.c_int, [_P, _I32]),
    "maple_ctx_set_critical_searches": (C.c_int, [_P, _I32]),
    "maple_ctx_set_head_searches": (C.c_int, [_P, _I32]),
    "maple_ctx_set_dense_scoring": (C.c_int, [_P, _I32, _I64]),
    "maple_update_partials": (C.c_int, [_P, _P, _I32, _P, _P, _P]),
    "maple_blen_sweep_sequential": (C.c_int, [_P, _P, _P, _P, _P]),
    "maple_search_stats": (C.c_int, [_P, _I32, _P]),
    "maple_launch_count": (_I64, [_P]),
}

class SearchParams(C.Structure):
    """maple_search_params (include/maple_b200.h)."""
    _fields_ = [("strictTopologyStopRules", _I32), ("allowedFailsTopology", _I32), ("deeperSearchForLongBranches", _I32),
                ("reserved", _I32), ("thresholdLogLKtopology", _D), ("thresholdTopologyPlacement", _D),
                ("thresholdLogLKoptimizationTopology", _D), ("thresholdLogLKconsecutivePlacement", _D),
                ("effectivelyNon0BLen", _D), ("BLenThresholdDeeperSearch", _D), ("defaultBLen", _D)]


class TreeRW(C.Structure):
    """maple_tree_rw (include/maple_b200.h): the bound tree and arena, writable."""
    _fields_ = [("key", _P), ("pay", _P), ("key_start", _P), ("pay_start", _P), ("nkeys", _P), ("npay", _P), ("tails", _P),
                ("cap_keys", _I64), ("cap_pay", _I64), ("dist", _P), ("dirty", _P)]


class PlaceParams(C.Structure):
    """maple_place_params (include/maple_b200.h)."""
    _fields_ = [("strictStopRules", _I32), ("allowedFails", _I32), ("deeperSearchForLongBranches", _I32), ("onlyFindIdentical", _I32),
                ("thresholdLogLK", _D), ("thresholdLogLKoptimization", _D), ("thresholdLogLKconsecutivePlacement", _D),
                ("effectivelyNon0BLen", _D), ("BLenThresholdDeeperSearch", _D), ("oneMutBLen", _D)]


PLACE_RESULT_FIELDS = [("bestNode", "i4"), ("status", "i4"), ("phase1", "i4"), ("missedMinors", "i4"), ("bestScore", "f8"),
                       ("bLenTop", "f8"), ("bLenBottom", "f8"), ("bLenAppend", "f8")]

# maple_search_result as a numpy record
SEARCH_RESULT_FIELDS = [("placement", "i4"), ("bestNode", "i4"), ("status", "i4"), ("phase1", "i4"), ("improvement", "f8"),
                        ("bestCurrentLK", "f8"), ("bestScore", "f8"), ("bLenTop", "f8"), ("bLenBottom", "f8"), ("bLenAppend", "f8")]

_lib = None


class MapleError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MapleError("CUDA library %s is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                             "maple_b200 has no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(ctx, rc, what):
    if rc != 0:
        msg = load().maple_last_error(ctx)
        raise MapleError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
