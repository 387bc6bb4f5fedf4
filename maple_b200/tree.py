"""Device-resident mirror of the reference's Tree (MAPLEv0.7.5.4.py:331-376) and of
reCalculateAllGenomeLists (:6013-6347) as level-synchronous batches of the merge kernel.

Host side keeps the reference's struct-of-arrays (node = int index): up, children, dist, plus
isTip = "no children and no minor sequences" (the flag every likelihood call takes) and numMinor.
The four genome-list families live in ONE arena in HBM; list id = family * nNodes + node with
family 0 = probVect (lower), 1 = probVectUpRight, 2 = probVectUpLeft, 3 = probVectTotUp.

This module does not handle MAT local references (mutations[node] lists, :3749): trees built here
express every list relative to the reference genome (the reference's --noLocalRef behaviour).
"""
from __future__ import annotations

import ctypes as C
import itertools
from typing import Optional

import numpy as np
import torch

from . import capi
from .engine import MapleEngine, _dp
from .genome_list import PackedLists, pack_lists, decode_stream

FAM_LOWER, FAM_UPRIGHT, FAM_UPLEFT, FAM_TOTUP = 0, 1, 2, 3
_EPOCH = itertools.count(1)  # one sequence for all arenas: a rebuilt arena never repeats an epoch a tree was bound at


class ListArena:
    """Growable arena of packed lists in HBM, addressed by list id."""

    def __init__(self, engine: MapleEngine, nLists: int, key_capacity: int, pay_capacity: int):
        self.eng = engine
        dev = engine.device
        self.n = nLists
        self.key = torch.zeros(max(key_capacity, 16), dtype=torch.int32, device=dev)
        self.pay = torch.zeros(max(pay_capacity, 16), dtype=torch.float64, device=dev)
        self.key_start = torch.full((nLists,), -1, dtype=torch.int64, device=dev)
        self.pay_start = torch.full((nLists,), -1, dtype=torch.int64, device=dev)
        self.nkeys = torch.zeros(nLists, dtype=torch.int32, device=dev)
        self.npay = torch.zeros(nLists, dtype=torch.int32, device=dev)
        self.key_tail = 0
        self.pay_tail = 0
        self.lRef, self.U = engine.model.lRef, int(engine.model.usingErrorRate)
        self.epoch = 0  # renewed whenever the tables or streams move in memory: holders of raw pointers (maple_tree_bind) rebind
        self._bind()

    def _bind(self):
        self.epoch = next(_EPOCH)
        self.eng.lists = self
        rc = self.eng.lib.maple_lists_bind(self.eng.ctx, _dp(self.key), _dp(self.pay), _dp(self.key_start), _dp(self.pay_start), self.n)
        capi.check(self.eng.ctx, rc, "maple_lists_bind")

    def _reserve(self, nk: int, npay: int):
        grew = False
        if self.key_tail + nk + 8 > self.key.numel():
            new = torch.zeros(int((self.key_tail + nk) * 1.5) + 1024, dtype=torch.int32, device=self.key.device)
            new[: self.key_tail] = self.key[: self.key_tail]
            self.key, grew = new, True
        if self.pay_tail + npay + 8 > self.pay.numel():
            new = torch.zeros(int((self.pay_tail + npay) * 1.5) + 1024, dtype=torch.float64, device=self.pay.device)
            new[: self.pay_tail] = self.pay[: self.pay_tail]
            self.pay, grew = new, True
        if grew:
            self._bind()

    def store(self, list_ids: torch.Tensor, src_key, src_pay, src_key_start, src_pay_start, nkeys, npay, status=None):
        """Append the given source lists to the arena and point list_ids at them (status != 0 -> None)."""
        ok = torch.ones_like(nkeys, dtype=torch.bool) if status is None else (status == 0)
        nk = torch.where(ok, nkeys, torch.zeros_like(nkeys)).long()
        npy = torch.where(ok, npay, torch.zeros_like(npay)).long()
        k_al = (nk + 3) // 4 * 4
        p_al = (npy + 1) // 2 * 2
        tot_k, tot_p = int(k_al.sum().item()), int(p_al.sum().item())
        self._reserve(tot_k, tot_p)
        dks = self.key_tail + torch.cumsum(k_al, 0) - k_al
        dps = self.pay_tail + torch.cumsum(p_al, 0) - p_al
        dks = torch.where(ok, dks, torch.full_like(dks, -1))
        n = list_ids.numel()
        # NOTE: every tensor whose pointer is passed must stay referenced until the call returns
        nk32, np32 = nk.int(), npy.int()
        rc = self.eng.lib.maple_lists_copy(self.eng.ctx, n, _dp(src_key), _dp(src_pay), _dp(src_key_start), _dp(src_pay_start),
                                           _dp(nk32), _dp(np32), _dp(self.key), _dp(self.pay), _dp(dks), _dp(dps),
                                           self.eng._stream())
        capi.check(self.eng.ctx, rc, "maple_lists_copy")
        self.key_start[list_ids] = dks
        self.pay_start[list_ids] = torch.where(ok, dps, torch.full_like(dps, -1))
        self.nkeys[list_ids] = nk32
        self.npay[list_ids] = np32
        self.key_tail += tot_k
        self.pay_tail += tot_p

    def store_packed(self, list_ids, packed: PackedLists):
        dev = self.key.device
        self.store(torch.as_tensor(list_ids, dtype=torch.int64, device=dev),
                   torch.from_numpy(packed.key.view(np.int32)).to(dev), torch.from_numpy(packed.pay).to(dev),
                   torch.from_numpy(packed.key_start).to(dev), torch.from_numpy(packed.pay_start).to(dev),
                   torch.from_numpy(packed.nkeys).to(dev), torch.from_numpy(packed.npay).to(dev),
                   torch.from_numpy((packed.key_start < 0).astype(np.int32)).to(dev))

    def add_ids(self, k: int) -> int:
        """Make room for k more list ids (all None); returns the first new id."""
        first, dev = self.n, self.key.device
        self.key_start = torch.cat([self.key_start, torch.full((k,), -1, dtype=torch.int64, device=dev)])
        self.pay_start = torch.cat([self.pay_start, torch.full((k,), -1, dtype=torch.int64, device=dev)])
        self.nkeys = torch.cat([self.nkeys, torch.zeros(k, dtype=torch.int32, device=dev)])
        self.npay = torch.cat([self.npay, torch.zeros(k, dtype=torch.int32, device=dev)])
        self.n += k
        self._bind()
        return first

    def get(self, list_id: int):
        ks = int(self.key_start[list_id].item())
        if ks < 0:
            return None
        nk, npay = int(self.nkeys[list_id].item()), int(self.npay[list_id].item())
        ps = int(self.pay_start[list_id].item())
        key = self.key[ks: ks + nk].cpu().numpy().view(np.uint32)
        pay = self.pay[ps: ps + npay + 1].cpu().numpy()
        return decode_stream(key, pay, 0, 0, self.lRef, self.U, nk)

    def to_host(self) -> PackedLists:
        return PackedLists(self.key[: max(self.key_tail, 4)].cpu().numpy().view(np.uint32), self.pay[: max(self.pay_tail, 2)].cpu().numpy(),
                           self.key_start.cpu().numpy(), self.pay_start.cpu().numpy(), self.nkeys.cpu().numpy(),
                           self.npay.cpu().numpy(), self.lRef, self.U)

    def mark(self):
        """State to come back to with release(): temporary lists (re-referenced copies, trial root vectors) are appended
        after the mark and dropped again."""
        return (self.n, self.key_tail, self.pay_tail)

    def release(self, mark):
        n, self.key_tail, self.pay_tail = mark
        if n != self.n:
            self.key_start, self.pay_start = self.key_start[:n].contiguous(), self.pay_start[:n].contiguous()
            self.nkeys, self.npay = self.nkeys[:n].contiguous(), self.npay[:n].contiguous()
            self.n = n
            self._bind()

    def add_lists(self, r) -> torch.Tensor:
        """Append the lists of a MergeResult under fresh ids; returns the ids (int64, device)."""
        k = r.nkeys.numel()
        first = self.add_ids(k)
        ids = torch.arange(first, first + k, dtype=torch.int64, device=self.key.device)
        self.store(ids, r.key, r.pay, r.key_start, r.pay_start, r.nkeys, r.npay, r.status)
        return ids

    def used_bytes(self) -> int:
        return self.key_tail * 4 + self.pay_tail * 8 + self.n * (8 + 8 + 4 + 4)


class DeviceTree:
    def __init__(self, engine: MapleEngine, up, child0, child1, dist, root: int, isTip=None, numMinor=None):
        self.eng = engine
        self.n = len(up)
        self.up = np.ascontiguousarray(up, np.int32)
        self.child0 = np.ascontiguousarray(child0, np.int32)
        self.child1 = np.ascontiguousarray(child1, np.int32)
        self.dist = np.ascontiguousarray(dist, np.float64)
        self.root = int(root)
        self.numMinor = np.zeros(self.n, np.int32) if numMinor is None else np.ascontiguousarray(numMinor, np.int32)
        self.isTip = ((self.child0 < 0) & (self.numMinor == 0)).astype(np.uint8) if isTip is None else np.ascontiguousarray(isTip, np.uint8)
        self._levels()
        dev = engine.device
        self.d_up = torch.from_numpy(self.up).to(dev)
        self.d_child0 = torch.from_numpy(self.child0).to(dev)
        self.d_child1 = torch.from_numpy(self.child1).to(dev)
        self.d_dist = torch.from_numpy(self.dist).to(dev)
        self.d_isTip = torch.from_numpy(self.isTip).to(dev)
        self.arena: Optional[ListArena] = None

    def lid(self, fam: int, node):
        return fam * self.n + node

    def _levels(self):
        n = self.n
        depth = np.full(n, -1, np.int32)
        order = [self.root]
        depth[self.root] = 0
        for nd in order:
            for c in (self.child0[nd], self.child1[nd]):
                if c >= 0:
                    depth[c] = depth[nd] + 1
                    order.append(int(c))
        self.preorder = np.array(order, np.int32)
        height = np.zeros(n, np.int32)
        for nd in reversed(order):
            if self.child0[nd] >= 0:
                height[nd] = 1 + max(height[self.child0[nd]], height[self.child1[nd]])
        self.depth, self.height = depth, height
        live = depth >= 0
        self.by_height = [np.nonzero(live & (height == h))[0].astype(np.int64) for h in range(int(height[self.root]) + 1)]
        self.by_depth = [np.nonzero(depth == d)[0].astype(np.int64) for d in range(int(depth.max()) + 1)]

    # ------------------------------------------------------------------ reCalculateAllGenomeLists (:6013-6347)
    def recalculate_all_lists(self, tip_nodes, tip_lists_packed: PackedLists, key_capacity: Optional[int] = None, max_restarts: int = 8):
        """Build all four list families on the device from the tip lists.

        First pass (post-order, :6031-6216): lower lists by batches of equal node height.
        Root (:6225-6245): rootVector of each child's lower list.  Second pass (pre-order,
        :6247-6345): probVectTotUp / UpRight / UpLeft by batches of equal depth.  Every stored list
        is shortened, as the reference does at :6201, :6267, :6308, :6330 and inside rootVector.

        Inconsistent zero-length branches (mergeVectors returns None, :4753-4758): the reference gives the two
        branches involved the length oneMutBLen/2 on first set-up (:6181-6183) or re-estimates them (updateBLen);
        here they get oneMutBLen/2 and the affected pass is redone (lower pass: on the spot; upper pass: restart).
        """
        for _ in range(max_restarts):
            if self._recalculate_once(tip_nodes, tip_lists_packed, key_capacity):
                return self.arena
        raise capi.MapleError("genome lists still inconsistent after %d restarts" % max_restarts)

    def _bump(self, nodes: torch.Tensor):
        half = 0.5 / self.eng.model.lRef  # oneMutBLen/2
        self.d_dist[nodes] = half
        self.dist[nodes.cpu().numpy()] = half

    def _recalculate_once(self, tip_nodes, tip_lists_packed: PackedLists, key_capacity: Optional[int]) -> bool:
        eng, n, dev = self.eng, self.n, self.eng.device
        nk_tips = int(tip_lists_packed.nkeys.sum())
        cap_k = key_capacity or max(1 << 16, int(nk_tips * 12))
        self.arena = ListArena(eng, 4 * n, cap_k, cap_k)
        A = self.arena
        A.store_packed(np.asarray(tip_nodes, np.int64) + FAM_LOWER * n, tip_lists_packed)
        t64 = lambda a: torch.as_tensor(a, dtype=torch.int64, device=dev)  # noqa: E731
        isTip = self.d_isTip
        c0, c1 = self.d_child0.long(), self.d_child1.long()

        def merge_store(list_ids, i1, b1, t1, i2, b2, t2, updown):
            """returns the mask of pairs whose merge came back None (those list ids stay None)"""
            r = eng.merge_batch(i1.int(), b1, t1, i2.int(), b2, t2, torch.full((i1.numel(),), 1 if updown else 0, dtype=torch.uint8, device=dev),
                                shorten=True)
            A.store(list_ids, r.key, r.pay, r.key_start, r.pay_start, r.nkeys, r.npay, r.status)
            return r.status != 0

        for h in range(1, len(self.by_height)):
            nodes = t64(self.by_height[h])
            if nodes.numel() == 0:
                continue
            a, b = c0[nodes], c1[nodes]
            dist = self.d_dist
            bad = merge_store(nodes + FAM_LOWER * n, a + FAM_LOWER * n, dist[a], isTip[a], b + FAM_LOWER * n, dist[b], isTip[b], False)
            if bool(bad.any().item()):
                nb, ab, bb = nodes[bad], a[bad], b[bad]
                if bool(((dist[ab] != 0) | (dist[bb] != 0)).any().item()):
                    raise capi.MapleError("mergeVectors returned None for branches of positive length (the reference raises too, :6196-6198)")
                self._bump(torch.cat([ab, bb]))
                dist = self.d_dist
                bad2 = merge_store(nb + FAM_LOWER * n, ab + FAM_LOWER * n, dist[ab], isTip[ab], bb + FAM_LOWER * n, dist[bb], isTip[bb], False)
                if bool(bad2.any().item()):
                    raise capi.MapleError("None vector when merging two vectors despite updating branch lengths (:6193-6195)")
        dist = self.d_dist
        root = self.root
        if self.child0[root] >= 0:
            ch = t64([self.child1[root], self.child0[root]])  # UpRight from child 1, UpLeft from child 0
            cap = (A.nkeys[ch + FAM_LOWER * n].long() + 3) // 4 * 4 + 4
            ks = torch.cumsum(cap, 0) - cap
            ps = ks * 6
            ok_ = torch.empty(int(cap.sum().item()) + 4, dtype=torch.int32, device=dev)
            op_ = torch.empty(int(cap.sum().item()) * 6 + 4, dtype=torch.float64, device=dev)
            nk = torch.empty(2, dtype=torch.int32, device=dev)
            npay = torch.empty(2, dtype=torch.int32, device=dev)
            ids32, bl_, tip_ = (ch + FAM_LOWER * n).int(), dist[ch].contiguous(), isTip[ch].contiguous()
            rc = eng.lib.maple_root_vector_batch(eng.ctx, 2, _dp(ids32), _dp(bl_), _dp(tip_), _dp(ok_), _dp(op_), _dp(ks), _dp(ps),
                                                 _dp(nk), _dp(npay), 1, eng._stream())
            capi.check(eng.ctx, rc, "maple_root_vector_batch")
            A.store(t64([root + FAM_UPRIGHT * n, root + FAM_UPLEFT * n]), ok_, op_, ks, ps, nk, npay)
        clean = True
        for d in range(1, len(self.by_depth)):
            nodes = t64(self.by_depth[d])
            par = self.d_up.long()[nodes]
            is0 = c0[par] == nodes
            vectUp = torch.where(is0, par + FAM_UPRIGHT * n, par + FAM_UPLEFT * n)
            have = A.key_start[vectUp] >= 0  # a parent whose upper list is None (being repaired) is skipped this time
            nodes, vectUp = nodes[have], vectUp[have]
            dn = dist[nodes]
            pos = dn > 0  # :6262 `if dist[node]`
            if bool(pos.any().item()):
                m = nodes[pos]
                bad = merge_store(m + FAM_TOTUP * n, vectUp[pos], dn[pos] / 2, torch.zeros_like(isTip[m]), m + FAM_LOWER * n, dn[pos] / 2,
                                  isTip[m], True)
                if bool(bad.any().item()):
                    raise capi.MapleError("probVectTotUp merge returned None on a branch of positive length")
            internal = c0[nodes] >= 0
            if bool(internal.any().item()):
                m, vu = nodes[internal], vectUp[internal]
                a, b = c0[m], c1[m]
                z = torch.zeros_like(isTip[m])
                ids = torch.cat([m + FAM_UPRIGHT * n, m + FAM_UPLEFT * n])
                other = torch.cat([b, a])
                bad = merge_store(ids, torch.cat([vu, vu]), torch.cat([dist[m], dist[m]]), torch.cat([z, z]),
                                  other + FAM_LOWER * n, torch.cat([dist[b], dist[a]]), torch.cat([isTip[b], isTip[a]]), True)
                if bool(bad.any().item()):
                    mm, oo = torch.cat([m, m])[bad], other[bad]
                    if bool(((dist[mm] != 0) | (dist[oo] != 0)).any().item()):
                        raise capi.MapleError("upper-list merge returned None for branches of positive length (:6300-6302)")
                    self._bump(torch.cat([mm, oo]))  # the reference re-estimates these two lengths (updateBLen, :6289-6299)
                    clean = False
        return clean

    # ------------------------------------------------------------------ calculateTreeLikelihood (:9721-9779)
    def tree_likelihood(self) -> float:
        """Sum over internal nodes of the Lkcontribution of mergeVectors(child0, dist0, isTip0, child1, dist1, isTip1,
        returnLK=True, numMinor1, numMinor2) (:9756) plus findProbRoot(probVect[root]) (:9775).  The merges only read the
        STORED lower lists, so all of them run as one batch.  Children that carry MAT mutations are re-referenced first
        (passGenomeListThroughBranch, :9750-9755), also on the device."""
        eng, n, dev, A = self.eng, self.n, self.eng.device, self.arena
        t64 = lambda a: torch.as_tensor(a, dtype=torch.int64, device=dev)  # noqa: E731
        reach = np.zeros(n, bool)
        reach[self.preorder] = True
        internal = t64(np.nonzero(reach & (self.child0 >= 0))[0])
        ids0 = self.d_child0.long()[internal] + FAM_LOWER * n
        ids1 = self.d_child1.long()[internal] + FAM_LOWER * n
        root_id = self.root + FAM_LOWER * n
        mutStart = getattr(self, "mutStart", None)
        mark = A.mark()
        try:
            if mutStart is not None:  # lists that must cross a local-reference branch get a temporary id
                nmut = np.diff(mutStart)
                hit = np.nonzero(reach & (nmut > 0))[0]
                if hit.size:
                    d_ms = torch.from_numpy(mutStart).to(dev)
                    d_mu = torch.from_numpy(np.ascontiguousarray(self.mut.reshape(-1))).to(dev)
                    tmp = A.add_lists(eng.pass_branch_batch(t64(hit) + FAM_LOWER * n, hit, np.ones(hit.size, np.uint8), d_ms, d_mu))
                    remap = torch.arange(4 * n, device=dev)
                    remap[t64(hit) + FAM_LOWER * n] = tmp
                    ids0, ids1 = remap[ids0], remap[ids1]
                    root_id = int(remap[root_id].item())
            total = 0.0
            if internal.numel():
                c0, c1 = self.d_child0.long()[internal], self.d_child1.long()[internal]
                nm = torch.from_numpy(self.numMinor).to(dev)
                r = eng.merge_batch(ids0.int(), self.d_dist[c0], self.d_isTip[c0], ids1.int(), self.d_dist[c1], self.d_isTip[c1],
                                    torch.full((internal.numel(),), capi.MAPLE_MERGE_RETURN_LK, dtype=torch.uint8, device=dev), nm[c0], nm[c1])
                if bool((r.status != 0).any().item()):
                    raise capi.MapleError("inconsistent lower genome list creation in tree_likelihood (the reference raises, :9762-9764)")
                total += float(r.lk.sum().item())
            total += float(eng.prob_root_batch([root_id]).cpu()[0])
        finally:
            A.release(mark)  # the re-referenced copies were temporaries
        return total

    # ------------------------------------------------------------------ traverseTreeToOptimizeBranchLengths(fastPass=True) (:8727)
    def optimize_branch_lengths(self, effectivelyNon0BLen: float, dirty=None):
        """One sweep over all dirty branches with the genome lists frozen (the reference's fastPass mode, see blen_sweep.py):
        the root's two branches by a scan of their split (a batch of mergeVectors(returnLK) + findProbRoot), every other
        branch by ONE maple_blen_batch launch over (upper list of the parent side, lower list of the node); the acceptance
        rule (:8866-8884) is applied to the returned lengths.  Updates dist on host and device; the lists are NOT refreshed --
        call recalculate_all_lists (or the per-family updates) afterwards.  Returns (number of updated branches, dirty flags
        after the sweep)."""
        from . import blen_sweep
        eng, n, dev, A = self.eng, self.n, self.eng.device, self.arena
        t64 = lambda a: torch.as_tensor(a, dtype=torch.int64, device=dev)  # noqa: E731
        dirty = np.ones(n, bool) if dirty is None else np.array(dirty, bool)
        root = self.root
        mutStart = getattr(self, "mutStart", None)
        nmut = np.zeros(n, np.int64) if mutStart is None else np.diff(mutStart).astype(np.int64)
        if mutStart is not None:
            d_ms = torch.from_numpy(mutStart).to(dev)
            d_mu = torch.from_numpy(np.ascontiguousarray(self.mut.reshape(-1))).to(dev)
        mark = A.mark()

        def passed(ids: torch.Tensor, through: np.ndarray, dir_up: bool) -> torch.Tensor:
            """ids with the lists that cross a local-reference branch (mutations[node], :8754-8759, :8824) replaced by
            temporary re-referenced copies."""
            hit = np.nonzero(nmut[through] > 0)[0]
            if hit.size == 0:
                return ids
            r = eng.pass_branch_batch(ids[t64(hit)], through[hit], np.full(hit.size, 1 if dir_up else 0, np.uint8), d_ms, d_mu)
            ids = ids.clone()
            ids[t64(hit)] = A.add_lists(r)
            return ids

        try:
            if self.child0[root] >= 0:
                c = np.array([self.child0[root], self.child1[root]], np.int64)
                cand = blen_sweep.root_split_candidates(float(self.dist[c[0]]), float(self.dist[c[1]]), eng.model.lRef, effectivelyNon0BLen)
                if cand is not None:
                    low = passed(t64(c) + FAM_LOWER * n, c, True)
                    k = len(cand[0])
                    tips = self.d_isTip[t64(c)]
                    r = eng.merge_batch(low[0].expand(k).int().contiguous(), torch.from_numpy(cand[0]).to(dev), tips[0].expand(k).contiguous(),
                                        low[1].expand(k).int().contiguous(), torch.from_numpy(cand[1]).to(dev), tips[1].expand(k).contiguous(),
                                        torch.full((k,), capi.MAPLE_MERGE_RETURN_LK, dtype=torch.uint8, device=dev))
                    if bool((r.status != 0).any().item()):
                        raise capi.MapleError("inconsistent root vector while scanning the root's branch lengths (the reference fails too, :8768)")
                    trial = A.add_lists(r)
                    if nmut[root] > 0:
                        trial = passed(trial, np.full(k, root, np.int64), True)
                    cost = (r.lk + eng.prob_root_batch(trial.int())).cpu().numpy()
                    b1, b2 = blen_sweep.choose_root_split(cost, cand[0], float(self.dist[c[0]]), float(self.dist[c[1]]))
                    self.dist[c[0]], self.dist[c[1]] = b1, b2
            nodes = blen_sweep.sweep_nodes(self.up, self.child0, self.child1, root, dirty)
            updates = 0
            if nodes.size:
                par = self.up[nodes].astype(np.int64)
                upv = t64(np.where(self.child0[par] == nodes, par + FAM_UPRIGHT * n, par + FAM_UPLEFT * n))
                if bool((A.key_start[upv] < 0).any().item()):
                    raise capi.MapleError("a swept node has no upper list on its parent's side: build the lists first")
                upv = passed(upv, nodes, False)
                best, st = eng.blen_batch(upv.int(), (t64(nodes) + FAM_LOWER * n).int(), self.d_isTip[t64(nodes)])
                new, changed, still = blen_sweep.accept(self.dist[nodes], best.cpu().numpy(), st.cpu().numpy() != 0)
                self.dist[nodes] = new
                dirty[nodes] = still
                updates = int(changed.sum())
            self.d_dist.copy_(torch.from_numpy(self.dist))
        finally:
            A.release(mark)
        return updates, dirty

    # ------------------------------------------------------------------ updatePartials (:5479) and the sequential sweep (:8727)
    def _run_rw(self, call, slack_keys: int = 1 << 18):
        """Common part of the calls that edit the lists on the device (maple_tree_rw): room behind the arena's tails, a copy of
        the tables and lengths to come back to if the room runs out (then the arena grows and the call is repeated), and the
        bookkeeping afterwards (tails, epoch, host copy of dist)."""
        eng, dev, A = self.eng, self.eng.device, self.arena
        if getattr(self, "_bound_epoch", None) is None:
            self.prepare_search()  # these calls read the tree arrays of the binding; the lists come with the call
        if not hasattr(self, "d_dirty") or self.d_dirty.numel() != self.n:
            self.d_dirty = torch.ones(self.n, dtype=torch.uint8, device=dev)
        while True:
            A._reserve(slack_keys, 6 * slack_keys)
            keep = (A.key_start.clone(), A.pay_start.clone(), A.nkeys.clone(), A.npay.clone(), self.d_dist.clone(), self.d_dirty.clone())
            tails = torch.tensor([A.key_tail, A.pay_tail], dtype=torch.int64, device=dev)
            rw = capi.TreeRW(_dp(A.key), _dp(A.pay), _dp(A.key_start), _dp(A.pay_start), _dp(A.nkeys), _dp(A.npay), _dp(tails),
                             int(A.key.numel()), int(A.pay.numel()), _dp(self.d_dist), _dp(self.d_dirty))
            status, result = call(rw)
            if status == 3:  # out of room: back to the state before the call, twice the room
                A.key_start.copy_(keep[0]); A.pay_start.copy_(keep[1]); A.nkeys.copy_(keep[2]); A.npay.copy_(keep[3])
                self.d_dist.copy_(keep[4]); self.d_dirty.copy_(keep[5])
                slack_keys *= 4
                if slack_keys > (1 << 28):
                    raise capi.MapleError("updatePartials keeps running out of arena room")
                continue
            if status != 0:
                raise capi.MapleError("the reference would raise inside updatePartials (inconsistent lists), status %d" % status)
            t = tails.cpu().numpy()
            A.key_tail, A.pay_tail = int(t[0]), int(t[1])
            A.epoch = next(_EPOCH)  # the lists changed: whoever holds sizes derived from them (maple_tree_bind) must bind again
            self.dist = self.d_dist.cpu().numpy().copy()
            return result

    def update_partials(self, node_list):
        """updatePartials(tree, nodeList) (:5479): node_list = [(node, direction), ...] as the reference builds it (the last entry
        is processed first); direction 2 = the change comes from the parent, 0 / 1 = from that child.  The lists are re-derived on
        the device, in the reference's order; self.d_dirty (uint8 per node) receives the reference's dirty marks."""
        eng = self.eng
        ent = np.ascontiguousarray(np.array([(int(a), int(b)) for a, b in node_list], np.int32).reshape(-1))

        def call(rw):
            st = C.c_int32(0)
            rc = eng.lib.maple_update_partials(eng.ctx, C.byref(rw), len(ent) // 2, ent.ctypes.data_as(C.c_void_p), C.byref(st), eng._stream())
            capi.check(eng.ctx, rc, "maple_update_partials")
            return int(st.value), None

        return self._run_rw(call)

    def optimize_branch_lengths_sequential(self, effectivelyNon0BLen: float, dirty=None):
        """traverseTreeToOptimizeBranchLengths(tree, root, fastPass=False) (:8727), the reference's default mode: the root's two
        branches by the scan of their split followed by updatePartials for each (:8745-8814), then every dirty branch in the
        reference's visiting order, each accepted change followed by updatePartials before the next estimate -- the loop runs on
        the device (maple_blen_sweep_sequential).  Lengths, dirty flags and the update count equal the reference's.
        Returns (number of updated branches, dirty flags after the sweep)."""
        from . import blen_sweep
        eng, n, dev, A = self.eng, self.n, self.eng.device, self.arena
        t64 = lambda a: torch.as_tensor(a, dtype=torch.int64, device=dev)  # noqa: E731
        root = self.root
        self.d_dirty = torch.ones(n, dtype=torch.uint8, device=dev) if dirty is None else torch.as_tensor(np.array(dirty, np.uint8), device=dev)
        mutStart = getattr(self, "mutStart", None)
        nmut = np.zeros(n, np.int64) if mutStart is None else np.diff(mutStart).astype(np.int64)
        if self.child0[root] >= 0:
            c = np.array([self.child0[root], self.child1[root]], np.int64)
            cand = blen_sweep.root_split_candidates(float(self.dist[c[0]]), float(self.dist[c[1]]), eng.model.lRef, effectivelyNon0BLen)
            if cand is not None:
                mark = A.mark()
                try:
                    low = t64(c) + FAM_LOWER * n
                    hit = np.nonzero(nmut[c] > 0)[0]
                    if mutStart is not None:
                        d_ms = torch.from_numpy(mutStart).to(dev)
                        d_mu = torch.from_numpy(np.ascontiguousarray(self.mut.reshape(-1))).to(dev)
                    if hit.size:
                        low = low.clone()
                        low[t64(hit)] = A.add_lists(eng.pass_branch_batch(low[t64(hit)], c[hit], np.ones(hit.size, np.uint8), d_ms, d_mu))
                    k = len(cand[0])
                    tips = self.d_isTip[t64(c)]
                    r = eng.merge_batch(low[0].expand(k).int().contiguous(), torch.from_numpy(cand[0]).to(dev), tips[0].expand(k).contiguous(),
                                        low[1].expand(k).int().contiguous(), torch.from_numpy(cand[1]).to(dev), tips[1].expand(k).contiguous(),
                                        torch.full((k,), capi.MAPLE_MERGE_RETURN_LK, dtype=torch.uint8, device=dev))
                    if bool((r.status != 0).any().item()):
                        raise capi.MapleError("inconsistent root vector while scanning the root's branch lengths (the reference fails too, :8768)")
                    trial = A.add_lists(r)
                    if nmut[root] > 0:
                        trial = A.add_lists(eng.pass_branch_batch(trial, np.full(k, root, np.int64), np.ones(k, np.uint8), d_ms, d_mu))
                    cost = (r.lk + eng.prob_root_batch(trial.int())).cpu().numpy()
                    b1, b2 = blen_sweep.choose_root_split(cost, cand[0], float(self.dist[c[0]]), float(self.dist[c[1]]))
                finally:
                    A.release(mark)
                for child, new in ((int(c[0]), b1), (int(c[1]), b2)):  # :8788-8797 (both work lists name (root, 0), as the reference's do)
                    self.dist[child] = new
                    self.d_dist.copy_(torch.from_numpy(self.dist))
                    self.update_partials([(child, 2), (root, 0)])

        def call(rw):
            st, upd = C.c_int32(0), C.c_int32(0)
            rc = eng.lib.maple_blen_sweep_sequential(eng.ctx, C.byref(rw), C.byref(upd), C.byref(st), eng._stream())
            capi.check(eng.ctx, rc, "maple_blen_sweep_sequential")
            return int(st.value), int(upd.value)

        updates = self._run_rw(call)
        return updates, self.d_dirty.cpu().numpy().astype(bool)

    # ------------------------------------------------------------------ construction from existing lists
    @classmethod
    def from_lists(cls, engine: MapleEngine, up, child0, child1, dist, root, isTip, lists: PackedLists, mutStart=None, mut=None,
                   numMinor=None):
        """A tree whose four list families already exist (list id = family*nNodes + node), e.g. a snapshot of the
        reference's Tree.  mutStart/mut: MAT mutation lists (mutations[node], :336) in CSR form."""
        t = cls(engine, up, child0, child1, dist, root, isTip=isTip, numMinor=numMinor)
        assert len(lists) == 4 * t.n
        t.arena = ListArena(engine, 4 * t.n, int(lists.key.size) + 64, int(lists.pay.size) + 64)
        t.arena.store_packed(np.arange(4 * t.n, dtype=np.int64), lists)
        if mutStart is not None and int(np.asarray(mutStart)[-1]) > 0:
            t.mutStart = np.ascontiguousarray(mutStart, np.int32)
            t.mut = np.ascontiguousarray(mut, np.int32).reshape(-1, 3)
        return t

    @classmethod
    def from_host_tree(cls, engine: MapleEngine, host_tree, root: int, tip_nodes, tip_lists):
        """An input tree as set up by maple_b200.newick.load_input_tree: all four list families are built on the device
        (reCalculateAllGenomeLists, :6431-6441)."""
        a = host_tree.arrays()
        t = cls(engine, a["up"], a["child0"], a["child1"], a["dist"], root, isTip=a["isTip"], numMinor=a["numMinor"])
        m = engine.model
        t.recalculate_all_lists(np.asarray(tip_nodes, np.int64), pack_lists(tip_lists, m.lRef, m.usingErrorRate))
        return t

    # ------------------------------------------------------------------ SPR search round (startTopologyUpdatesParallel, :9580)
    def prepare_search(self):
        """Fill probVectTotUp of zero-length children of the root (the reference does it lazily and order-dependently
        inside the round, :7198-7200), then hand the tree arrays to the search kernel."""
        eng, n, dev, A = self.eng, self.n, self.eng.device, self.arena
        root = self.root
        if self.child0[root] >= 0:
            for c, fam in ((int(self.child0[root]), FAM_UPRIGHT), (int(self.child1[root]), FAM_UPLEFT)):
                if self.dist[c] == 0.0 and int(A.key_start[FAM_TOTUP * n + c].item()) < 0 and int(A.key_start[fam * n + root].item()) >= 0:
                    r = eng.merge_batch([fam * n + root], [0.0], [0], [FAM_LOWER * n + c], [0.0], [0], [capi.MAPLE_MERGE_UPDOWN])
                    A.store(torch.as_tensor([FAM_TOTUP * n + c], dtype=torch.int64, device=dev), r.key, r.pay, r.key_start, r.pay_start,
                            r.nkeys, r.npay, r.status)
        mutStart = getattr(self, "mutStart", None)
        self._d_mutStart = None if mutStart is None else torch.from_numpy(mutStart).to(dev)
        self._d_mut = None if mutStart is None else torch.from_numpy(np.ascontiguousarray(self.mut.reshape(-1))).to(dev)
        rc = eng.lib.maple_tree_bind(eng.ctx, n, root, _dp(self.d_up), _dp(self.d_child0), _dp(self.d_child1), _dp(self.d_dist),
                                     _dp(self.d_isTip), _dp(self._d_mutStart), _dp(self._d_mut), _dp(A.nkeys), _dp(A.npay))
        capi.check(eng.ctx, rc, "maple_tree_bind")
        self._bound_epoch = A.epoch

    # what a search that had an SM to itself would have taken next to the others: the cycles kept in search_cost are the busy kind
    ALONE_TO_BUSY = 1.6

    def spr_search(self, nodes, params: "capi.SearchParams", scratch_keys: int = 0, max_concurrent: int = 0, cycles=None,
                   schedule: bool = True, critical=None):
        """Run the searches of the listed nodes; returns a device tensor of raw records [n, 64 bytes] viewed as uint8 (in the
        order of `nodes`) -- search_records() reads it as a numpy record array.

        schedule: searches differ by orders of magnitude in length and the kernel's lanes pull them from the list in order, so
        the list is handed over longest-first, using the SM cycles each node's search took the last time it ran on this tree
        (kept on the device in self.search_cost; nodes never searched before go first).  The tree hardly changes between the
        rounds of a run, so this is the LPT rule with last round's lengths.  Results do not depend on the order.

        critical: how many of the longest searches get an SM each (maple_ctx_set_critical_searches).  None = found out by
        measurement: a run repeats rounds of about the same size on a tree that hardly changes, so the first sorted round of a
        size runs without, the next with the longest searches (_critical_count) on SMs of their own, and the faster of the two
        shapes -- device time of the launch -- is kept for that size.  (On a whole 100 000-sequence round the plain launch
        wins; on an eighth of it, one GPU's share of an 8-GPU round, the other one does by a third.)"""
        eng, dev = self.eng, self.eng.device
        if getattr(self, "_bound_epoch", None) != self.arena.epoch:
            self.prepare_search()  # never bound, or the arena's tables moved since (temporary lists added / released)
        nodes = torch.as_tensor(nodes, dtype=torch.int32, device=dev).contiguous()
        n = nodes.numel()
        cost = getattr(self, "search_cost", None)
        if cost is None or cost.numel() != self.n:
            cost = self.search_cost = torch.full((self.n,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
        perm = None
        k = 0
        trial = None
        if schedule and n > 1:
            mine = cost[nodes.long()]
            perm = torch.argsort(mine, descending=True, stable=True)
            run_nodes = nodes[perm].contiguous()
            if critical is None:
                k, trial = self._critical_auto(mine[perm], n)
            else:
                k = int(critical)
        else:
            run_nodes = nodes
        k = max(0, min(k, n // 4))
        if k != getattr(self, "_critical_set", 0):
            eng.set_critical_searches(k)
            self._critical_set = k
        out = torch.zeros((n, 64), dtype=torch.uint8, device=dev)
        cyc = torch.zeros(n, dtype=torch.int64, device=dev)
        if trial is not None:
            trial["ev"] = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            trial["ev"][0].record()
        rc = eng.lib.maple_spr_search_batch(eng.ctx, C.byref(params), n, _dp(run_nodes), _dp(out), int(scratch_keys),
                                            int(max_concurrent), _dp(cyc), eng._stream())
        capi.check(eng.ctx, rc, "maple_spr_search_batch")
        if trial is not None:
            trial["ev"][1].record()
        cost[run_nodes.long()] = cyc
        if k:
            cost[run_nodes[:k].long()] = (cyc[:k].double() * self.ALONE_TO_BUSY).long()
        if perm is not None:
            back = torch.empty_like(out)
            back[perm] = out
            out = back
            if cycles is not None:
                cycles[perm] = cyc
        elif cycles is not None:
            cycles.copy_(cyc)
        return out

    def _critical_auto(self, sorted_cost: torch.Tensor, n: int):
        """The number of searches to give an SM each in this round, and (while the two launch shapes are still being compared
        for this batch size) the record in which spr_search leaves the events that time the launch.  See spr_search."""
        if not torch.cuda.is_available():
            return 0, None
        st = getattr(self, "_critical_state", None)
        if st is None or abs(n - st["n"]) > 0.2 * st["n"]:
            st = self._critical_state = {"n": n, "ms": {}, "pending": None, "k": 0, "choice": None}
        if st["pending"] is not None:  # the launch timed last time
            which, rec = st["pending"]
            st["pending"] = None
            try:
                rec["ev"][1].synchronize()
                st["ms"][which] = float(rec["ev"][0].elapsed_time(rec["ev"][1]))
            except Exception:
                st["ms"][which] = float("inf")
        if st["choice"] is not None:
            return (st["k"] if st["choice"] == "on" else 0), None
        if "off" not in st["ms"]:
            if float(sorted_cost[0].item()) >= float(torch.iinfo(torch.int64).max) / 2:
                return 0, None  # searches with no recorded length yet: this round only measures them
            rec = {}
            st["pending"] = ("off", rec)
            return 0, rec
        if "on" not in st["ms"]:
            st["k"] = self._critical_count(sorted_cost, n)
            if st["k"] == 0:
                st["choice"] = "off"
                return 0, None
            rec = {}
            st["pending"] = ("on", rec)
            return st["k"], rec
        st["choice"] = "on" if st["ms"]["on"] < 0.97 * st["ms"]["off"] else "off"
        return (st["k"] if st["choice"] == "on" else 0), None

    def _critical_count(self, sorted_cost: torch.Tensor, n: int) -> int:
        """How many of the longest searches would get an SM each: those that took at least half as long as the longest one, at
        most a ninth of the SMs.  sorted_cost: last round's cycles of this batch's searches, longest first (device)."""
        sms = int(getattr(self.eng, "num_sms", 148))
        cap = max(1, sms // 9)
        if n < 64:
            return 0
        c = sorted_cost[:cap].double().cpu().numpy()
        if c[0] <= 0 or c[0] >= float(torch.iinfo(torch.int64).max) / 2:
            return 0
        return int(max(4, min(cap, (c >= 0.5 * c[0]).sum())))

    # ------------------------------------------------------------------ findBestParentForNewSample for a batch (:7912, :11190-11287)
    def stage_samples(self, samples: PackedLists):
        """Append the tip genome lists of new samples to the arena as temporaries and bind the tree again (the arena's tables
        moved).  Returns (device int32 ids, mark); give the mark to release_samples() when the batch is done."""
        A, dev = self.arena, self.eng.device
        n = len(samples)
        mark = A.mark()
        first = A.add_ids(n)
        A.store_packed(np.arange(first, first + n, dtype=np.int64), samples)
        self.prepare_search()
        return torch.arange(first, first + n, dtype=torch.int32, device=dev), mark

    def release_samples(self, mark):
        self.arena.release(mark)
        if getattr(self, "_bound_epoch", None) != self.arena.epoch:
            self.prepare_search()

    def place_staged(self, ids: torch.Tensor, params: "capi.PlaceParams", scratch_keys: int = 0) -> torch.Tensor:
        """maple_place_batch on staged samples: raw records [n, 48 bytes] on the device (asynchronous)."""
        eng, dev = self.eng, self.eng.device
        out = torch.zeros((ids.numel(), 48), dtype=torch.uint8, device=dev)
        rc = eng.lib.maple_place_batch(eng.ctx, C.byref(params), ids.numel(), _dp(ids), _dp(out), int(scratch_keys), eng._stream())
        capi.check(eng.ctx, rc, "maple_place_batch")
        return out

    @staticmethod
    def place_records(out: torch.Tensor) -> np.ndarray:
        return out.cpu().numpy().view(np.dtype(capi.PLACE_RESULT_FIELDS)).reshape(-1)

    def place_samples(self, samples: PackedLists, params: "capi.PlaceParams", scratch_keys: int = 0) -> np.ndarray:
        """Place every tip genome list of `samples` (probVectTerminalNode output) on the frozen tree; returns records
        (capi.PLACE_RESULT_FIELDS).  The tree is not modified and the arena is left as it was found."""
        eng, dev = self.eng, self.eng.device
        ids, mark = self.stage_samples(samples)
        try:
            rec = self.place_records(self.place_staged(ids, params, scratch_keys))
            retry, keys = np.nonzero(rec["status"] == 3)[0], max(int(scratch_keys), 4096)
            variant = getattr(eng, "place_variant", capi.DEFAULT_PLACE_VARIANT)
            try:
                while retry.size and keys < (1 << 22):  # per-sample scratch (lists or the bestNodes table) exhausted: again with 8x
                    keys *= 8
                    if variant != 0:  # the retries go through the one-sample-per-thread kernel, whose whole scratch scales with `keys`
                        eng.set_place_variant(0)
                    again = self.place_records(self.place_staged(ids[torch.as_tensor(retry, device=dev)].contiguous(), params, keys))
                    rec[retry] = again
                    retry = retry[again["status"] == 3]
            finally:
                if variant != 0:
                    eng.set_place_variant(variant)
            if retry.size:
                raise capi.MapleError("%d samples still exhaust their scratch with %d entries (first: sample %d)" % (retry.size, keys, int(retry[0])))
        finally:
            self.release_samples(mark)
        return rec

    @staticmethod
    def search_records(out: torch.Tensor) -> np.ndarray:
        return out.cpu().numpy().view(np.dtype(capi.SEARCH_RESULT_FIELDS)).reshape(-1)

    def lists_of(self, node: int):
        A = self.arena
        return [A.get(self.lid(f, node)) for f in range(4)]
