"""Branch-length sweep over a tree with frozen genome lists: traverseTreeToOptimizeBranchLengths(tree, root, fastPass=True)
(MAPLEv0.7.5.4.py:8727-8890) as one batch.

With fastPass the reference does not touch any genome list during the traversal (every updatePartials call is skipped, :8789,
:8875), so the estimate of a branch depends only on the stored lists: the traversal is a set of independent
estimateBranchLengthWithDerivative calls -- one `maple_blen_batch` launch -- plus the scan of the root's two branches
(:8745-8786: the sum of the two lengths is kept and the split is chosen among half-mutation steps by the likelihood of the root,
a batch of mergeVectors(returnLK=True) + findProbRoot).  This module holds the backend-neutral parts: which nodes enter the
batch, the candidate splits of the root, and the acceptance rule; DeviceTree.optimize_branch_lengths runs them on the device,
tests/ runs the same plan over the CPU oracle against sweeps recorded from the reference.

The reference's default mode (fastPass=False) re-derives the lists around a branch as soon as its length changes
(updatePartials, :5479), so later estimates of the same traversal see earlier changes -- a sequential Gauss-Seidel sweep that
is not a batch and is not reproduced here; a round of `optimize_branch_lengths` + `recalculate_all_lists` is its Jacobi
counterpart.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def sweep_nodes(up, child0, child1, root: int, dirty) -> np.ndarray:
    """Nodes whose branch the traversal re-estimates (:8815-8829): everything below the root's children that is dirty, in
    the reference's visiting order (stack, grandchildren of the root pushed child0-side first)."""
    if child0[root] < 0:
        return np.zeros(0, np.int64)
    stack = []
    for c in (child0[root], child1[root]):
        if child0[c] >= 0:
            stack.extend((int(child0[c]), int(child1[c])))
    out = []
    while stack:
        nd = stack.pop()
        if dirty[nd]:
            out.append(nd)
        if child0[nd] >= 0:
            stack.extend((int(child0[nd]), int(child1[nd])))
    return np.array(out, np.int64)


def root_split_candidates(d1: float, d2: float, lRef: int, effectivelyNon0BLen: float):
    """(bLen1[], bLen2[]) tried for the root's two branches (:8747-8762), or None when both are effectively zero."""
    if not (d1 > effectivelyNon0BLen or d2 > effectivelyNon0BLen):
        return None
    tot = (d1 + d2) * lRef
    k = max(1, round(tot)) * 2 + 1  # python round(): half to even, as in the reference
    b1 = np.array([min(tot, i / 2) for i in range(k)], np.float64)
    b2 = np.array([max(tot - x, 0.0) for x in b1], np.float64)
    return b1 / lRef, b2 / lRef


def choose_root_split(cost: np.ndarray, b1: np.ndarray, d1: float, d2: float) -> Tuple[float, float]:
    """First strict maximum (:8781-8785); the second length keeps the sum."""
    best = int(np.argmax(cost))  # numpy returns the first maximum, like the `>` scan
    return float(b1[best]), max(d1 + d2 - float(b1[best]), 0.0)


def accept(dist: np.ndarray, best: np.ndarray, is_false: np.ndarray):
    """The update rule (:8831, :8866-8884) for the swept nodes: returns (new_dist, updated, still_dirty).
    A length is replaced when the estimate or the old length is zero/False (and not both), or when they differ by more than
    1 %; otherwise the node is marked clean."""
    best = np.where(is_false, 0.0, best)
    either = (best != 0) | (dist != 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = dist / best
    change = either & ((best == 0) | (dist == 0) | (ratio > 1.01) | (ratio < 0.99))
    return np.where(change, best, dist), change, change
