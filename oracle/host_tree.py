"""CPU construction of the four genome-list families of a tree (reCalculateAllGenomeLists, MAPLEv0.7.5.4.py:6013-6347) as
level-synchronous batches of the oracle's mergeVectors -- TEST INFRASTRUCTURE, like everything under oracle/.

It exists so that `bench.py --impl reference` (the CPU arm) can build the same frozen tree as the GPU arm without touching
the CUDA library, and as a second, independent producer of the lists for tests.  The orchestration (levels by node height,
then by depth; zero-length branch repair with oneMutBLen/2) is the one maple_b200/tree.py runs on the device.
"""
from __future__ import annotations

import numpy as np

from maple_b200.genome_list import PackedLists, pack_lists

FAM_LOWER, FAM_UPRIGHT, FAM_UPLEFT, FAM_TOTUP = 0, 1, 2, 3


class HostArena:
    def __init__(self, nLists, lRef, U, cap=1 << 16):
        self.key = np.zeros(cap, np.uint32)
        self.pay = np.zeros(cap, np.float64)
        self.key_start = np.full(nLists, -1, np.int64)
        self.pay_start = np.full(nLists, -1, np.int64)
        self.nkeys = np.zeros(nLists, np.int32)
        self.npay = np.zeros(nLists, np.int32)
        self.kt = self.pt = 0
        self.lRef, self.U = lRef, U

    def view(self) -> PackedLists:
        return PackedLists(self.key, self.pay, self.key_start, self.pay_start, self.nkeys, self.npay, self.lRef, self.U)

    def _reserve(self, nk, npay):
        if self.kt + nk + 8 > self.key.size:
            new = np.zeros(int((self.kt + nk) * 1.5) + 1024, np.uint32)
            new[: self.kt] = self.key[: self.kt]
            self.key = new
        if self.pt + npay + 8 > self.pay.size:
            new = np.zeros(int((self.pt + npay) * 1.5) + 1024, np.float64)
            new[: self.pt] = self.pay[: self.pt]
            self.pay = new

    def store(self, ids, key, pay, key_start, pay_start, nkeys, npay, status=None):
        ok = np.ones(len(ids), bool) if status is None else (np.asarray(status) == 0)
        nk = np.where(ok, nkeys, 0).astype(np.int64)
        npy = np.where(ok, npay, 0).astype(np.int64)
        k_al, p_al = (nk + 3) // 4 * 4, (npy + 1) // 2 * 2
        self._reserve(int(k_al.sum()), int(p_al.sum()))
        dks = self.kt + np.cumsum(k_al) - k_al
        dps = self.pt + np.cumsum(p_al) - p_al
        for i in np.nonzero(ok)[0]:  # contiguous slices: cheap memcpy per list
            a, b = int(key_start[i]), int(pay_start[i])
            self.key[dks[i]: dks[i] + nk[i]] = key[a: a + nk[i]]
            self.pay[dps[i]: dps[i] + npy[i]] = pay[b: b + npy[i]]
        self.key_start[ids] = np.where(ok, dks, -1)
        self.pay_start[ids] = np.where(ok, dps, -1)
        self.nkeys[ids] = nk
        self.npay[ids] = npy
        self.kt += int(k_al.sum())
        self.pt += int(p_al.sum())


def build_tree_lists(orc, up, child0, child1, dist, root, tip_nodes, tip_lists, lRef, U, max_restarts=8, isTip=None):
    """Returns (PackedLists with list id = family*nNodes + node, dist after zero-length repairs, isTip).
    isTip: "no children and no minor sequences" per node when the tree carries minor sequences (default: no children)."""
    up, child0, child1 = (np.asarray(a, np.int32) for a in (up, child0, child1))
    dist = np.array(dist, np.float64)
    n = len(up)
    isTip = (child0 < 0).astype(np.uint8) if isTip is None else np.asarray(isTip, np.uint8)
    depth = np.full(n, -1, np.int32)
    order = [int(root)]
    depth[root] = 0
    for nd in order:
        for c in (child0[nd], child1[nd]):
            if c >= 0:
                depth[c] = depth[nd] + 1
                order.append(int(c))
    height = np.zeros(n, np.int32)
    for nd in reversed(order):
        if child0[nd] >= 0:
            height[nd] = 1 + max(height[child0[nd]], height[child1[nd]])
    live = depth >= 0
    by_height = [np.nonzero(live & (height == h))[0] for h in range(int(height[root]) + 1)]
    by_depth = [np.nonzero(depth == d)[0] for d in range(int(depth.max()) + 1)]
    tips = pack_lists(tip_lists, lRef, U)
    half = 0.5 / lRef

    def once():
        A = HostArena(4 * n, lRef, U, cap=max(1 << 16, int(tips.nkeys.sum()) * 12))
        A.store(np.asarray(tip_nodes, np.int64) + FAM_LOWER * n, tips.key, tips.pay, tips.key_start, tips.pay_start, tips.nkeys, tips.npay)

        def merge_store(ids, i1, b1, t1, i2, b2, t2, updown):
            r = orc.merge_batch(A.view(), i1, b1, t1, i2, b2, t2, np.full(len(i1), 1 if updown else 0, np.uint8), shorten=True)
            A.store(ids, r["key"], r["pay"], r["key_start"], r["pay_start"], r["nkeys"], r["npay"], r["status"])
            return r["status"] != 0

        for h in range(1, len(by_height)):
            nodes = by_height[h]
            if nodes.size == 0:
                continue
            a, b = child0[nodes].astype(np.int64), child1[nodes].astype(np.int64)
            bad = merge_store(nodes + FAM_LOWER * n, a + FAM_LOWER * n, dist[a], isTip[a], b + FAM_LOWER * n, dist[b], isTip[b], False)
            if bad.any():
                nb, ab, bb = nodes[bad], a[bad], b[bad]
                assert not ((dist[ab] != 0) | (dist[bb] != 0)).any(), "mergeVectors returned None for branches of positive length"
                dist[np.concatenate([ab, bb])] = half
                bad2 = merge_store(nb + FAM_LOWER * n, ab + FAM_LOWER * n, dist[ab], isTip[ab], bb + FAM_LOWER * n, dist[bb], isTip[bb], False)
                assert not bad2.any()
        if child0[root] >= 0:
            for c, fam in ((int(child1[root]), FAM_UPRIGHT), (int(child0[root]), FAM_UPLEFT)):
                v = A.view().get(c + FAM_LOWER * n)
                rv = orc.shorten(orc.root_vector(v, float(dist[c]), bool(isTip[c])))
                pk = pack_lists([rv], lRef, U)
                A.store(np.array([root + fam * n]), pk.key, pk.pay, pk.key_start, pk.pay_start, pk.nkeys, pk.npay)
        clean = True
        for d in range(1, len(by_depth)):
            nodes = by_depth[d].astype(np.int64)
            par = up[nodes].astype(np.int64)
            vectUp = np.where(child0[par] == nodes, par + FAM_UPRIGHT * n, par + FAM_UPLEFT * n)
            have = A.key_start[vectUp] >= 0
            nodes, vectUp = nodes[have], vectUp[have]
            dn = dist[nodes]
            pos = dn > 0
            if pos.any():
                m = nodes[pos]
                bad = merge_store(m + FAM_TOTUP * n, vectUp[pos], dn[pos] / 2, np.zeros(len(m), np.uint8), m + FAM_LOWER * n, dn[pos] / 2, isTip[m], True)
                assert not bad.any(), "probVectTotUp merge returned None on a branch of positive length"
            internal = child0[nodes] >= 0
            if internal.any():
                m, vu = nodes[internal], vectUp[internal]
                a, b = child0[m].astype(np.int64), child1[m].astype(np.int64)
                z = np.zeros(len(m), np.uint8)
                ids = np.concatenate([m + FAM_UPRIGHT * n, m + FAM_UPLEFT * n])
                other = np.concatenate([b, a])
                bad = merge_store(ids, np.concatenate([vu, vu]), np.concatenate([dist[m], dist[m]]), np.concatenate([z, z]),
                                  other + FAM_LOWER * n, np.concatenate([dist[b], dist[a]]), np.concatenate([isTip[b], isTip[a]]), True)
                if bad.any():
                    mm, oo = np.concatenate([m, m])[bad], other[bad]
                    assert not ((dist[mm] != 0) | (dist[oo] != 0)).any()
                    dist[np.concatenate([mm, oo])] = half
                    clean = False
        return A, clean

    for _ in range(max_restarts):
        A, clean = once()
        if clean:
            break
    else:
        raise RuntimeError("genome lists still inconsistent after %d restarts" % max_restarts)
    # probVectTotUp of zero-length children of the root: pre-filled like DeviceTree.prepare_search does
    if child0[root] >= 0:
        for c, fam in ((int(child0[root]), FAM_UPRIGHT), (int(child1[root]), FAM_UPLEFT)):
            if dist[c] == 0.0 and A.key_start[FAM_TOTUP * n + c] < 0 and A.key_start[fam * n + root] >= 0:
                r = orc.merge_batch(A.view(), [fam * n + root], [0.0], [0], [FAM_LOWER * n + c], [0.0], [0], [1])
                A.store(np.array([FAM_TOTUP * n + c]), r["key"], r["pay"], r["key_start"], r["pay_start"], r["nkeys"], r["npay"], r["status"])
    pl = PackedLists(A.key[: max(A.kt, 4) + 8].copy(), A.pay[: max(A.pt, 2) + 8].copy(), A.key_start, A.pay_start, A.nkeys, A.npay, lRef, U)
    return pl, dist, isTip


def tree_likelihood(orc, pl, child0, child1, dist, root, isTip, numMinor=None):
    """calculateTreeLikelihood (:9721-9779) over the stored lower lists: sum of the mergeVectors(returnLK=True) contributions of
    every internal node reachable from the root, plus findProbRoot of the root's lower list.  `orc` needs its root tables."""
    n = len(child0)
    numMinor = np.zeros(n, np.int32) if numMinor is None else np.asarray(numMinor, np.int32)
    internal, stack = [], [int(root)]
    while stack:
        nd = stack.pop()
        if child0[nd] >= 0:
            internal.append(nd)
            stack.extend((int(child0[nd]), int(child1[nd])))
    total = 0.0
    if internal:
        a, b = child0[internal].astype(np.int64), child1[internal].astype(np.int64)
        r = orc.merge_batch(pl, a + FAM_LOWER * n, dist[a], isTip[a], b + FAM_LOWER * n, dist[b], isTip[b],
                            np.full(len(a), 2, np.uint8), numMinor[a], numMinor[b])
        assert not (r["status"] != 0).any(), "inconsistent lower genome lists"
        total += float(np.sum(r["lk"]))
    return total + orc.prob_root(pl.get(int(root) + FAM_LOWER * n))
