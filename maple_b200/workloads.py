"""Candidate batches for the placement-cost kernel.

`neighbourhood_pairs` enumerates, for every non-root node s of a tree (the subtree that an SPR move
would prune), the branches within `radius` hops of its parent that do not lie inside s's own
subtree -- the region findBestParentTopology's phase-1 walk visits around the pruning point
(MAPLEv0.7.5.4.py:6964-7429) -- and pairs each of them with s:
    P = probVectTotUp[t] (the stored mid-branch list),  C = probVect[s],  bLen = dist[s],  isTipC = isTip[s]
which is exactly the call at :7011 / :7223 once the passed partials have converged to the stored ones
(needsUpdating == False), and the call of findBestParentForNewSample at :8050.  Pairs are grouped by s.
"""
from __future__ import annotations

import torch

from .tree import DeviceTree, FAM_LOWER, FAM_TOTUP


def neighbourhood_pairs(tree: DeviceTree, radius: int, nodes: torch.Tensor | None = None):
    dev = tree.eng.device
    n = tree.n
    up, c0, c1 = tree.d_up.long(), tree.d_child0.long(), tree.d_child1.long()
    if nodes is None:
        nodes = torch.arange(n, device=dev)
        nodes = nodes[(up >= 0)]
    src = nodes.clone()
    cur = up[nodes]
    prev = nodes.clone()
    out_src, out_t = [], []
    has_tot = tree.arena.key_start[FAM_TOTUP * n: FAM_TOTUP * n + n] >= 0
    for _ in range(radius):
        ok = has_tot[cur]
        out_src.append(src[ok])
        out_t.append(cur[ok])
        nxt_src, nxt_cur, nxt_prev = [], [], []
        for nb in (up[cur], c0[cur], c1[cur]):
            m = (nb >= 0) & (nb != prev)
            nxt_src.append(src[m])
            nxt_cur.append(nb[m])
            nxt_prev.append(cur[m])
        src, cur, prev = torch.cat(nxt_src), torch.cat(nxt_cur), torch.cat(nxt_prev)
        if src.numel() == 0:
            break
    s_all, t_all = torch.cat(out_src), torch.cat(out_t)
    order = torch.argsort(s_all, stable=True)
    s_all, t_all = s_all[order], t_all[order]
    pIdx = (t_all + FAM_TOTUP * n).int()
    cIdx = (s_all + FAM_LOWER * n).int()
    isTip = tree.d_isTip[s_all].contiguous()
    bLen = tree.d_dist[s_all].contiguous()
    return s_all.int(), pIdx, cIdx, isTip, bLen


def algorithmic_bytes(tree: DeviceTree, s_all, pIdx, cIdx) -> int:
    """Bytes one scoring pass must move at the very least, in THIS format: each candidate's parent list
    (keys 4 B + payload 8 B/double) once per pair, each child list once per search, 17 B of arguments in
    and 8 B of score out per pair, two 16-B list descriptors per pair."""
    A = tree.arena
    p = pIdx.long()
    parent = (A.nkeys[p].long() * 4 + A.npay[p].long() * 8).sum()
    uniq = torch.unique(cIdx.long())
    child = (A.nkeys[uniq].long() * 4 + A.npay[uniq].long() * 8).sum()
    n = pIdx.numel()
    return int(parent.item() + child.item()) + n * (17 + 8 + 16) + int(uniq.numel()) * 16
