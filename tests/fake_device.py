"""TEST INFRASTRUCTURE: a stand-in for libmaple_b200.so whose entry points run the CPU oracle on HOST memory, so that the
host-side orchestration above the C ABI (ListArena, DeviceTree: level-synchronous list building, tree likelihood, branch-length
sweeps, input-tree set-up) can be exercised in a container without a GPU.  It binds the same call signatures the ctypes table
in maple_b200/capi.py uses (pointers arrive as c_void_p, here of CPU torch tensors).  Nothing under maple_b200/ knows about
it; the product path always loads the CUDA library (capi.load) and fails without a device."""
import ctypes as C

import numpy as np
import torch

from maple_b200.engine import MapleEngine
from oracle.oracle import Oracle, OrTree


def _arr(ptr, n, dtype):
    if ptr is None or n == 0:
        return np.zeros(0, dtype)
    addr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr)
    return np.ctypeslib.as_array((C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(addr)).view(dtype)


def _off(ptr, nbytes):
    return C.c_void_p(ptr.value + int(nbytes))


class FakeLib:
    def __init__(self, orc: Oracle, engine):
        self.o, self.L, self.mp, self.eng = orc, orc.L, orc.mp, engine
        self.launches = 0

    # ---- lists
    def maple_lists_bind(self, ctx, key, pay, ks, ps, n):
        self.key, self.pay, self.ks, self.ps, self.n = key, pay, ks, ps, n
        return 0

    def maple_lists_copy(self, ctx, n, skey, spay, sks, sps, nk, npay, dkey, dpay, dks, dps, stream):
        sks_, sps_, nk_, np_ = _arr(sks, n, np.int64), _arr(sps, n, np.int64), _arr(nk, n, np.int32), _arr(npay, n, np.int32)
        dks_, dps_ = _arr(dks, n, np.int64), _arr(dps, n, np.int64)
        for i in range(n):
            if dks_[i] < 0 or nk_[i] == 0:
                continue
            C.memmove(dkey.value + 4 * int(dks_[i]), skey.value + 4 * int(sks_[i]), 4 * int(nk_[i]))
            if np_[i]:
                C.memmove(dpay.value + 8 * int(dps_[i]), spay.value + 8 * int(sps_[i]), 8 * int(np_[i]))
        return 0

    # ---- batches
    def maple_merge_batch(self, ctx, n, i1, b1, t1, i2, b2, t2, flags, nm1, nm2, ok, op, ks, ps, nk, npay, lk, st, shorten, stream):
        self.launches += 1
        self.L.or_merge_batch(self.mp, self.key, self.pay, self.ks, self.ps, n, i1, b1, t1, i2, b2, t2, flags, nm1, nm2, ok, op, ks, ps,
                              nk, npay, lk, st)
        if shorten:
            self.L.or_shorten_slots(self.mp, n, ok, op, ks, ps, nk, npay, st)
        return 0

    def maple_blen_batch(self, ctx, n, pIdx, cIdx, tip, scratch, ss, out, st, stream):
        self.launches += 1
        nkeys = C.c_void_p(self.eng.lists.nkeys.data_ptr())
        self.L.or_blen_batch(self.mp, self.key, self.pay, self.ks, self.ps, nkeys, n, pIdx, cIdx, tip, out, st)
        return 0

    def maple_prob_root_batch(self, ctx, n, idx, out, stream):
        self.launches += 1
        idx_, out_ = _arr(idx, n, np.int32), _arr(out, n, np.float64)
        ks, ps = _arr(self.ks, self.n, np.int64), _arr(self.ps, self.n, np.int64)
        for i in range(n):
            out_[i] = self.L.or_prob_root(self.mp, _off(self.key, 4 * ks[idx_[i]]), _off(self.pay, 8 * ps[idx_[i]]))
        return 0

    def maple_ctx_set_root_tables(self, ctx, cb, pl):
        return 0  # the oracle built with_root_tables=True already holds them

    def maple_root_vector_batch(self, ctx, n, idx, bLen, tip, ok, op, oks, ops, nk, npay, shorten, stream):
        self.launches += 1
        idx_, bl_, tip_ = _arr(idx, n, np.int32), _arr(bLen, n, np.float64), _arr(tip, n, np.uint8)
        oks_, ops_, nk_, np_ = _arr(oks, n, np.int64), _arr(ops, n, np.int64), _arr(nk, n, np.int32), _arr(npay, n, np.int32)
        ks, ps = _arr(self.ks, self.n, np.int64), _arr(self.ps, self.n, np.int64)
        for i in range(n):
            a, b = C.c_int32(0), C.c_int32(0)
            k, p = _off(ok, 4 * oks_[i]), _off(op, 8 * ops_[i])
            self.L.or_root_vector(self.mp, _off(self.key, 4 * ks[idx_[i]]), _off(self.pay, 8 * ps[idx_[i]]), float(bl_[i]), int(tip_[i]),
                                  k, p, C.addressof(a), C.addressof(b))
            if shorten:
                self.L.or_shorten(self.mp, k, p, k, p, C.addressof(a), C.addressof(b))
            nk_[i], np_[i] = a.value, b.value
        return 0

    def maple_pass_branch_batch(self, ctx, n, idx, mutNode, dirUp, mutStart, mut, ok, op, oks, ops, nk, npay, stream):
        self.launches += 1
        idx_, mn_, du_ = _arr(idx, n, np.int32), _arr(mutNode, n, np.int32), _arr(dirUp, n, np.uint8)
        oks_, ops_, nk_, np_ = _arr(oks, n, np.int64), _arr(ops, n, np.int64), _arr(nk, n, np.int32), _arr(npay, n, np.int32)
        ks, ps = _arr(self.ks, self.n, np.int64), _arr(self.ps, self.n, np.int64)
        ms = _arr(mutStart, int(mn_.max()) + 2 if n else 0, np.int32)
        for i in range(n):
            a, b = C.c_int32(0), C.c_int32(0)
            m0, m1 = int(ms[mn_[i]]), int(ms[mn_[i] + 1])
            self.L.or_pass_branch(self.mp, _off(self.key, 4 * ks[idx_[i]]), _off(self.pay, 8 * ps[idx_[i]]), _off(mut, 12 * m0), m1 - m0,
                                  int(du_[i]), _off(ok, 4 * oks_[i]), _off(op, 8 * ops_[i]), C.addressof(a), C.addressof(b))
            nk_[i], np_[i] = a.value, b.value
        return 0

    # ---- the search seam and placement batches: the kernel source on the host (tests/hostsim), which sizes its scratch like the
    # device does (status 3 when it is exhausted) -- so the retry logic of the wrappers can be exercised too
    def maple_tree_bind(self, ctx, n, root, up, c0, c1, dist, isTip, mutStart, mut, nkeys, npay):
        self.tree = (n, root, up, c0, c1, dist, isTip, mutStart, mut, nkeys, npay)
        return 0

    def _or_tree(self):
        n, root, up, c0, c1, dist, isTip, mutStart, mut, nkeys, npay = self.tree
        t = OrTree()
        t.nNodes, t.root = n, root
        t.up, t.child0, t.child1, t.dist, t.isTip, t.mutStart, t.mut = up, c0, c1, dist, isTip, mutStart, mut
        t.key, t.pay, t.keyStart, t.payStart, t.nkeys = self.key, self.pay, self.ks, self.ps, nkeys
        return t

    def _hs(self):
        from hostsim import lib
        return lib()

    # ---- updatePartials / the sequential sweep: update.cuh on the host (tests/hostsim: hs_update)
    def _update(self, rw_ref, mode, n_entries, entries, out_updates, out_status):
        self.launches += 1
        rw = rw_ref._obj
        t = self._or_tree()
        f = self._hs().hs_update
        f.restype = C.c_int
        f.argtypes = [C.c_void_p] * 9 + [C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        upd = C.c_int32(0)
        st = f(self.mp, C.addressof(t), rw.key, rw.pay, rw.key_start, rw.pay_start, rw.nkeys, rw.npay, rw.tails, rw.cap_keys, rw.cap_pay, rw.dist,
               rw.dirty, mode, n_entries, entries, C.addressof(upd))
        out_status._obj.value = st
        if out_updates is not None:
            out_updates._obj.value = upd.value
        return 0

    def maple_update_partials(self, ctx, rw, n_entries, entries, out_status, stream):
        return self._update(rw, 0, n_entries, entries, None, out_status)

    def maple_blen_sweep_sequential(self, ctx, rw, out_updates, out_status, stream):
        return self._update(rw, 1, 0, None, out_updates, out_status)

    def maple_ctx_set_place_variant(self, ctx, variant):
        self.place_variant = variant
        return 0

    def maple_place_batch(self, ctx, params, n, sampleLists, out, scratch_keys, stream):
        self.launches += 1
        ids = _arr(sampleLists, n, np.int32).astype(np.int64)
        ks, ps = _arr(self.ks, self.n, np.int64), _arr(self.ps, self.n, np.int64)
        nkeys_all = _arr(self.tree[9], self.n, np.int32)
        sks, sps, snk = np.ascontiguousarray(ks[ids]), np.ascontiguousarray(ps[ids]), np.ascontiguousarray(nkeys_all[ids])
        t = self._or_tree()
        keys = scratch_keys if scratch_keys > 0 else 4096
        variant = getattr(self, "place_variant", 3)  # the device default
        has_mut = self.tree[7] is not None and int(_arr(self.tree[7], self.tree[0] + 1, np.int32)[-1]) > 0
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        if variant == 0 or (variant == 1 and has_mut):
            self._hs().or_place_batch(self.mp, C.addressof(t), params, n, self.key, self.pay, p(sks), p(sps), p(snk), keys, out)
        else:
            self._hs().hs_place_batch_scan(self.mp, C.addressof(t), params, n, self.key, self.pay, p(sks), p(sps), p(snk), keys, self.tree[10],
                                           variant - 1, out)
        return 0

    def maple_ctx_set_search_variant(self, ctx, variant):
        self.search_variant = variant
        return 0

    def maple_ctx_set_scan_service(self, ctx, n):
        return 0

    def maple_ctx_set_lanes_per_warp(self, ctx, n):
        return 0

    def maple_ctx_set_critical_searches(self, ctx, n):
        return 0

    def maple_ctx_set_head_searches(self, ctx, n):
        return 0

    def maple_ctx_set_dense_scoring(self, ctx, mode, max_bytes):
        return 0

    def maple_ctx_set_scan_min_size(self, ctx, n):
        return 0

    def maple_spr_search_batch(self, ctx, params, n, nodes, out, scratch_keys, max_concurrent, cycles, stream):
        """Variant 1: the straight-line search, no retry.  Other variants: the per-lane state machine, and -- like the device -- a
        second pass with 8x the scratch for the searches that exhausted theirs."""
        self.launches += 1
        t = self._or_tree()
        keys = scratch_keys if scratch_keys > 0 else 8192
        if getattr(self, "search_variant", 0) == 1:
            self._hs().or_search_batch(self.mp, C.addressof(t), params, n, nodes, keys, 1, out)
            return 0
        self._hs().hs_search_batch_fsm(self.mp, C.addressof(t), params, n, nodes, keys, out)
        rec = _arr(out, n * 16, np.int32).reshape(n, 16)  # 64-byte records; status is the third int32
        again = np.nonzero(rec[:, 2] == 3)[0]
        if again.size:
            sub = np.ascontiguousarray(_arr(nodes, n, np.int32)[again])
            tmp = np.zeros((again.size, 16), np.int32)
            self._hs().hs_search_batch_fsm(self.mp, C.addressof(t), params, again.size, sub.ctypes.data_as(C.c_void_p), keys * 8,
                                           tmp.ctypes.data_as(C.c_void_p))
            rec[again] = tmp
        return 0

    def maple_last_error(self, ctx):
        return b""

    def maple_launch_count(self, ctx):
        return self.launches


class FakeEngine(MapleEngine):
    """MapleEngine over FakeLib: same python code paths, CPU tensors."""

    def __init__(self, model):  # noqa: super().__init__ needs a CUDA device on purpose
        self.model = model
        self.device = torch.device("cpu")
        self.ctx = None
        self.lists = None
        self._root_tables = True
        self.lib = FakeLib(Oracle(model, with_root_tables=True), self)

    def _stream(self):
        return None

    def set_place_variant(self, variant: int):
        self.lib.maple_ctx_set_place_variant(None, int(variant))
        self.place_variant = int(variant)

    def update_model(self):
        self.lib = FakeLib(Oracle(self.model, with_root_tables=True), self)
