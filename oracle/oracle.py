"""ctypes front-end of the CPU oracle (oracle/maple_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never from maple_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmaple_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "maple_oracle.c")
    if force or not os.path.isfile(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"], env={**os.environ, "MAKEFLAGS": ""})
    return LIB_PATH


class OrModel(C.Structure):
    _fields_ = [
        ("lRef", C.c_int32), ("U", C.c_int32), ("errSS", C.c_int32), ("rateVar", C.c_int32),
        ("Q", C.c_double * 16), ("pi", C.c_double * 4), ("errorRate", C.c_double), ("totError", C.c_double),
        ("thresholdProb", C.c_double), ("thresholdDiffForUpdate", C.c_double),
        ("thresholdFoldChangeUpdate", C.c_double), ("minBLenSensitivity", C.c_double),
        ("siteRates", C.c_void_p), ("errorRates", C.c_void_p), ("cumRate", C.c_void_p), ("cumErr", C.c_void_p),
        ("cumBases", C.c_void_p), ("piLogErrCum", C.c_void_p),
    ]


class OrTree(C.Structure):
    _fields_ = [("nNodes", C.c_int32), ("root", C.c_int32), ("up", C.c_void_p), ("child0", C.c_void_p), ("child1", C.c_void_p),
                ("dist", C.c_void_p), ("isTip", C.c_void_p), ("mutStart", C.c_void_p), ("mut", C.c_void_p), ("key", C.c_void_p),
                ("pay", C.c_void_p), ("keyStart", C.c_void_p), ("payStart", C.c_void_p), ("nkeys", C.c_void_p)]


class OrSearchParams(C.Structure):
    _fields_ = [("strictTopologyStopRules", C.c_int32), ("allowedFailsTopology", C.c_int32),
                ("deeperSearchForLongBranches", C.c_int32), ("reserved", C.c_int32),
                ("thresholdLogLKtopology", C.c_double), ("thresholdTopologyPlacement", C.c_double),
                ("thresholdLogLKoptimizationTopology", C.c_double), ("thresholdLogLKconsecutivePlacement", C.c_double),
                ("effectivelyNon0BLen", C.c_double), ("BLenThresholdDeeperSearch", C.c_double), ("defaultBLen", C.c_double)]


class OrSearchResult(C.Structure):
    _fields_ = [("placement", C.c_int32), ("bestNode", C.c_int32), ("status", C.c_int32), ("phase1", C.c_int32),
                ("improvement", C.c_double), ("bestCurrentLK", C.c_double), ("bestScore", C.c_double), ("bLenTop", C.c_double),
                ("bLenBottom", C.c_double), ("bLenAppend", C.c_double)]


SEARCH_RESULT_DTYPE = np.dtype([("placement", np.int32), ("bestNode", np.int32), ("status", np.int32), ("phase1", np.int32),
                                ("improvement", np.float64), ("bestCurrentLK", np.float64), ("bestScore", np.float64),
                                ("bLenTop", np.float64), ("bLenBottom", np.float64), ("bLenAppend", np.float64)])


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_lib = None


def declare(L):
    """ctypes signatures of the oracle's entry points (also used for tests/hostsim, which exports the same names)."""
    L.or_append.restype = C.c_double
    L.or_append.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_double]
    L.or_merge.restype = C.c_int
    L.or_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.or_shorten.restype = None
    L.or_shorten.argtypes = [C.c_void_p] * 7
    L.or_blen.restype = C.c_int
    L.or_blen.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p]
    L.or_differ.restype = C.c_int
    L.or_differ.argtypes = [C.c_void_p] * 5
    L.or_pass_branch.restype = None
    L.or_pass_branch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
    L.or_root_vector.restype = None
    L.or_root_vector.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int] + [C.c_void_p] * 4
    L.or_prob_root.restype = C.c_double
    L.or_prob_root.argtypes = [C.c_void_p] * 3
    L.or_append_batch.restype = None
    L.or_append_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64] + [C.c_void_p] * 5
    L.or_merge_batch.restype = None
    L.or_merge_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64] + [C.c_void_p] * 17
    L.or_blen_batch.restype = None
    L.or_blen_batch.argtypes = [C.c_void_p] * 6 + [C.c_int64] + [C.c_void_p] * 5
    L.or_differ_batch.restype = None
    L.or_differ_batch.argtypes = [C.c_void_p] * 5 + [C.c_int64] + [C.c_void_p] * 3
    L.or_search_batch.restype = None
    L.or_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    L.or_num_threads.restype = C.c_int
    if hasattr(L, "or_set_num_threads"):  # the oracle's own library; stand-ins that reuse this table (tests/hostsim) lack it
        L.or_set_num_threads.restype = None
        L.or_set_num_threads.argtypes = [C.c_int]
    L.or_shorten_slots.restype = None
    L.or_shorten_slots.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
    L.or_is_minor.restype = C.c_int
    L.or_is_minor.argtypes = [C.c_void_p] * 5 + [C.c_int]
    L.or_place_batch.restype = None
    L.or_place_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
    return L


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = declare(C.CDLL(LIB_PATH))
    return _lib


class OrPlaceParams(C.Structure):
    _fields_ = [("strictStopRules", C.c_int32), ("allowedFails", C.c_int32), ("deeperSearchForLongBranches", C.c_int32),
                ("onlyFindIdentical", C.c_int32), ("thresholdLogLK", C.c_double), ("thresholdLogLKoptimization", C.c_double),
                ("thresholdLogLKconsecutivePlacement", C.c_double), ("effectivelyNon0BLen", C.c_double),
                ("BLenThresholdDeeperSearch", C.c_double), ("oneMutBLen", C.c_double)]


PLACE_RESULT_DTYPE = np.dtype([("bestNode", "i4"), ("status", "i4"), ("phase1", "i4"), ("missedMinors", "i4"), ("bestScore", "f8"),
                               ("bLenTop", "f8"), ("bLenBottom", "f8"), ("bLenAppend", "f8")])


class Oracle:
    """Binds a MapleModel (maple_b200.model) to the C oracle; works on packed streams."""

    def __init__(self, model, with_root_tables: bool = False):
        from maple_b200.genome_list import pack_lists, decode_stream  # data containers only
        self._pack, self._decode = pack_lists, decode_stream
        self.model = model
        self.L = lib()
        m = OrModel()
        m.lRef, m.U, m.errSS, m.rateVar = model.lRef, int(model.usingErrorRate), int(model.errorRateSiteSpecific), int(model.useRateVariation)
        for i in range(16):
            m.Q[i] = float(model.Q.reshape(-1)[i])
        for i in range(4):
            m.pi[i] = float(model.rootFreqs[i])
        m.errorRate, m.totError = float(model.errorRate), float(model.totError)
        m.thresholdProb = model.thresholdProb
        m.thresholdDiffForUpdate = model.thresholdDiffForUpdate
        m.thresholdFoldChangeUpdate = model.thresholdFoldChangeUpdate
        m.minBLenSensitivity = model.minBLenSensitivity
        self._keep = [model.siteRates, model.errorRates, model.cumulativeRate, model.cumulativeErrorRate]
        m.siteRates, m.errorRates = _p(model.siteRates), _p(model.errorRates)
        m.cumRate, m.cumErr = _p(model.cumulativeRate), _p(model.cumulativeErrorRate)
        if with_root_tables:
            cb = np.ascontiguousarray(model.cumulative_bases())
            pl = model.root_freqs_log_error_cumulative() if model.usingErrorRate else None
            self._keep += [cb, pl]
            m.cumBases, m.piLogErrCum = _p(cb), _p(pl)
        self.m = m
        self.mp = C.addressof(m)

    # ---- helpers on python lists (pack on the fly)
    def _one(self, gl):
        return self._pack([gl], self.model.lRef, self.model.usingErrorRate)

    def _out(self, nk_cap):
        return np.zeros(max(nk_cap, 1), dtype=np.uint32), np.zeros(max(6 * nk_cap, 1), dtype=np.float64)

    def _dec(self, ok, op):
        return self._decode(ok, op, 0, 0, self.model.lRef, int(self.model.usingErrorRate))

    def append(self, P, C_, isTipC, bLen):
        a, b = self._one(P), self._one(C_)
        return self.L.or_append(self.mp, _p(a.key), _p(a.pay), _p(b.key), _p(b.pay), int(bool(isTipC)), float(bLen))

    def merge(self, v1, b1, t1, v2, b2, t2, returnLK=False, isUpDown=False, numMinor1=0, numMinor2=0):
        a, b = self._one(v1), self._one(v2)
        ok, op = self._out(int(a.nkeys[0] + b.nkeys[0]))
        nk, npay, lk = C.c_int32(0), C.c_int32(0), C.c_double(0.0)
        st = self.L.or_merge(self.mp, _p(a.key), _p(a.pay), float(b1), int(bool(t1)), _p(b.key), _p(b.pay), float(b2),
                             int(bool(t2)), (1 if isUpDown else 0) | (2 if returnLK else 0), int(numMinor1), int(numMinor2),
                             _p(ok), _p(op), C.addressof(nk), C.addressof(npay), C.addressof(lk))
        if st != 0:
            return (None, None) if returnLK else None
        out = self._dec(ok, op)
        return (out, lk.value) if returnLK else out

    def shorten(self, v):
        a = self._one(v)
        ok, op = self._out(int(a.nkeys[0]))
        nk, npay = C.c_int32(0), C.c_int32(0)
        self.L.or_shorten(self.mp, _p(a.key), _p(a.pay), _p(ok), _p(op), C.addressof(nk), C.addressof(npay))
        return self._dec(ok, op)

    def blen(self, P, C_, fromTipC=False):
        a, b = self._one(P), self._one(C_)
        ais = np.zeros(int(a.nkeys[0] + b.nkeys[0]) + 1)
        out = C.c_double(0.0)
        st = self.L.or_blen(self.mp, _p(a.key), _p(a.pay), _p(b.key), _p(b.pay), int(bool(fromTipC)), _p(ais), C.addressof(out))
        return None if st == 1 else out.value

    def differ(self, v1, v2):
        if v2 is None:
            return True
        a, b = self._one(v1), self._one(v2)
        return bool(self.L.or_differ(self.mp, _p(a.key), _p(a.pay), _p(b.key), _p(b.pay)))

    def pass_branch(self, v, mutations, dirIsUp=False):
        a = self._one(v)
        mut = np.ascontiguousarray(np.array(mutations, dtype=np.int32).reshape(-1, 3))
        ok, op = self._out(int(a.nkeys[0]) + 2 * len(mut) + 2)
        nk, npay = C.c_int32(0), C.c_int32(0)
        self.L.or_pass_branch(self.mp, _p(a.key), _p(a.pay), _p(mut), len(mut), int(bool(dirIsUp)), _p(ok), _p(op),
                              C.addressof(nk), C.addressof(npay))
        return self._dec(ok, op)

    def root_vector(self, v, bLen, isFromTip):
        a = self._one(v)
        ok, op = self._out(int(a.nkeys[0]))
        nk, npay = C.c_int32(0), C.c_int32(0)
        self.L.or_root_vector(self.mp, _p(a.key), _p(a.pay), float(bLen), int(bool(isFromTip)), _p(ok), _p(op),
                              C.addressof(nk), C.addressof(npay))
        return self._dec(ok, op)

    def prob_root(self, v):
        a = self._one(v)
        return self.L.or_prob_root(self.mp, _p(a.key), _p(a.pay))

    # ---- batch entry points over a PackedLists arena (all host cores via OpenMP)
    def append_batch(self, pl, pIdx, cIdx, isTip, bLen):
        n = len(pIdx)
        out = np.empty(n, dtype=np.float64)
        pIdx, cIdx = np.ascontiguousarray(pIdx, np.int32), np.ascontiguousarray(cIdx, np.int32)
        isTip, bLen = np.ascontiguousarray(isTip, np.uint8), np.ascontiguousarray(bLen, np.float64)
        self.L.or_append_batch(self.mp, _p(pl.key), _p(pl.pay), _p(pl.key_start), _p(pl.pay_start), n, _p(pIdx), _p(cIdx),
                               _p(isTip), _p(bLen), _p(out))
        return out

    def merge_batch(self, pl, idx1, b1, t1, idx2, b2, t2, flags, numMinor1=None, numMinor2=None, shorten=False):
        n = len(idx1)
        idx1, idx2 = np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32)
        b1, b2 = np.ascontiguousarray(b1, np.float64), np.ascontiguousarray(b2, np.float64)
        t1, t2 = np.ascontiguousarray(t1, np.uint8), np.ascontiguousarray(t2, np.uint8)
        flags = np.ascontiguousarray(flags, np.uint8)
        nm1 = None if numMinor1 is None else np.ascontiguousarray(numMinor1, np.int32)
        nm2 = None if numMinor2 is None else np.ascontiguousarray(numMinor2, np.int32)
        cap = (pl.nkeys[idx1].astype(np.int64) + pl.nkeys[idx2].astype(np.int64))
        cap = (cap + 3) // 4 * 4
        ks = np.zeros(n, dtype=np.int64)
        ks[1:] = np.cumsum(cap)[:-1]
        ps = ks * 6
        ok = np.zeros(int(cap.sum()) + 4, dtype=np.uint32)
        op = np.zeros(int(cap.sum()) * 6 + 4, dtype=np.float64)
        nk, npay = np.zeros(n, np.int32), np.zeros(n, np.int32)
        lk, st = np.zeros(n, np.float64), np.zeros(n, np.int32)
        self.L.or_merge_batch(self.mp, _p(pl.key), _p(pl.pay), _p(pl.key_start), _p(pl.pay_start), n, _p(idx1), _p(b1), _p(t1),
                              _p(idx2), _p(b2), _p(t2), _p(flags), _p(nm1), _p(nm2), _p(ok), _p(op), _p(ks), _p(ps), _p(nk),
                              _p(npay), _p(lk), _p(st))
        if shorten:
            self.L.or_shorten_slots(self.mp, n, _p(ok), _p(op), _p(ks), _p(ps), _p(nk), _p(npay), _p(st))
        return {"key": ok, "pay": op, "key_start": ks, "pay_start": ps, "nkeys": nk, "npay": npay, "lk": lk, "status": st}

    def blen_batch(self, pl, pIdx, cIdx, fromTip):
        n = len(pIdx)
        pIdx, cIdx = np.ascontiguousarray(pIdx, np.int32), np.ascontiguousarray(cIdx, np.int32)
        fromTip = np.ascontiguousarray(fromTip, np.uint8)
        out, st = np.zeros(n, np.float64), np.zeros(n, np.int32)
        self.L.or_blen_batch(self.mp, _p(pl.key), _p(pl.pay), _p(pl.key_start), _p(pl.pay_start), _p(pl.nkeys), n, _p(pIdx),
                             _p(cIdx), _p(fromTip), _p(out), _p(st))
        return out, st

    def differ_batch(self, pl, idx1, idx2):
        n = len(idx1)
        idx1, idx2 = np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32)
        out = np.zeros(n, np.uint8)
        self.L.or_differ_batch(self.mp, _p(pl.key), _p(pl.pay), _p(pl.key_start), _p(pl.pay_start), n, _p(idx1), _p(idx2), _p(out))
        return out

    def search_batch(self, tree: dict, lists, params: dict, nodes, scratch_keys: int = 1 << 20, lazy_mode: int = 1):
        """tree: host arrays up/child0/child1 (int32, -1 = none), dist, isTip, root, optional mutStart/mut;
        lists: PackedLists with list id = family*nNodes + node.  Returns a structured array (SEARCH_RESULT_DTYPE).
        lazy_mode 0: the reference's order-dependent lazy probVectTotUp fill (:7198-7200), single thread, `nodes` in the
        reference's order; lazy_mode 1: pre-filled, all host cores (the semantics of the device path)."""
        n = len(tree["up"])
        keep = {k: np.ascontiguousarray(tree[k], dt) for k, dt in (("up", np.int32), ("child0", np.int32), ("child1", np.int32),
                                                                   ("dist", np.float64), ("isTip", np.uint8))}
        t = OrTree()
        t.nNodes, t.root = n, int(tree["root"])
        t.up, t.child0, t.child1, t.dist, t.isTip = (_p(keep[k]) for k in ("up", "child0", "child1", "dist", "isTip"))
        if tree.get("mutStart") is not None:
            keep["mutStart"] = np.ascontiguousarray(tree["mutStart"], np.int32)
            keep["mut"] = np.ascontiguousarray(tree["mut"], np.int32)
            t.mutStart, t.mut = _p(keep["mutStart"]), _p(keep["mut"])
        t.key, t.pay, t.keyStart, t.payStart, t.nkeys = _p(lists.key), _p(lists.pay), _p(lists.key_start), _p(lists.pay_start), _p(lists.nkeys)
        sp = OrSearchParams()
        for k, v in params.items():
            setattr(sp, k, v)
        nodes = np.ascontiguousarray(nodes, np.int32)
        out = np.zeros(len(nodes), dtype=SEARCH_RESULT_DTYPE)
        self.L.or_search_batch(self.mp, C.addressof(t), C.addressof(sp), len(nodes), _p(nodes), int(scratch_keys), int(lazy_mode), _p(out))
        return out

    def _tree_struct(self, tree: dict, lists):
        n = len(tree["up"])
        keep = {k: np.ascontiguousarray(tree[k], dt) for k, dt in (("up", np.int32), ("child0", np.int32), ("child1", np.int32),
                                                                   ("dist", np.float64), ("isTip", np.uint8))}
        t = OrTree()
        t.nNodes, t.root = n, int(tree["root"])
        t.up, t.child0, t.child1, t.dist, t.isTip = (_p(keep[k]) for k in ("up", "child0", "child1", "dist", "isTip"))
        if tree.get("mutStart") is not None:
            keep["mutStart"] = np.ascontiguousarray(tree["mutStart"], np.int32)
            keep["mut"] = np.ascontiguousarray(tree["mut"], np.int32)
            t.mutStart, t.mut = _p(keep["mutStart"]), _p(keep["mut"])
        t.key, t.pay, t.keyStart, t.payStart, t.nkeys = _p(lists.key), _p(lists.pay), _p(lists.key_start), _p(lists.pay_start), _p(lists.nkeys)
        return t, keep

    def is_minor(self, v1, v2, onlyFindIdentical=False):
        """isMinorSequence(probVect1, probVect2, onlyFindIdentical) (:5919)."""
        a, b = self._one(v1), self._one(v2)
        return int(self.L.or_is_minor(self.mp, _p(a.key), _p(a.pay), _p(b.key), _p(b.pay), 1 if onlyFindIdentical else 0))

    def place_batch(self, tree: dict, lists, params: dict, samples, scratch_keys: int = 1 << 20):
        """findBestParentForNewSample (:7912) for every list of `samples` (PackedLists of tip genome lists) on the frozen tree.
        Returns a structured array (PLACE_RESULT_DTYPE)."""
        t, keep = self._tree_struct(tree, lists)
        pp = OrPlaceParams()
        for k, v in params.items():
            setattr(pp, k, v)
        out = np.zeros(len(samples), dtype=PLACE_RESULT_DTYPE)
        self.L.or_place_batch(self.mp, C.addressof(t), C.addressof(pp), len(samples), _p(samples.key), _p(samples.pay),
                              _p(samples.key_start), _p(samples.pay_start), _p(samples.nkeys), int(scratch_keys), _p(out))
        return out

    def num_threads(self):
        return int(self.L.or_num_threads())

    def use_all_cores(self) -> int:
        """One OpenMP thread per core this process may run on (its affinity mask), regardless of OMP_NUM_THREADS --
        torchrun sets that to 1 for its workers.  Returns the thread count now in effect."""
        import os
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        self.L.or_set_num_threads(int(n))
        return self.num_threads()
