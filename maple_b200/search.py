"""Host-side mirror of the reference's parallel SPR seam (MAPLEv0.7.5.4.py:9580-9716, :12283-12312).

`start_topology_updates_parallel(tree, ...)` plays the role of Pool.map(startTopologyUpdatesParallel, inputs)
followed by the concatenation of the workers' lists: on a frozen tree it searches every dirty node whose
replacements count allows it and returns proposedMoves = [(node, placementNode, improvement), ...] sorted ascending
by improvement, exactly what the reference hands to applySPRMovesParallel (:12312-12316).  The searches run on the
GPU (one per thread); nothing is evaluated on the CPU.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from .sharding import moves_from_records
from .tree import DeviceTree


def search_params(lRef: int, strictTopologyStopRules: bool, allowedFailsTopology: int, thresholdLogLKtopology: float,
                  thresholdTopologyPlacement: float = -0.1, thresholdLogLKoptimizationTopology: Optional[float] = None,
                  thresholdLogLKconsecutivePlacement: float = 1.0, deeperSearchForLongBranches: bool = False,
                  defaultBLen: float = 0.000033) -> capi.SearchParams:
    """Arguments named as in the reference.  thresholdLogLKtopology / ...optimizationTopology are the values AFTER the
    reference multiplied them by log(lRef) (:3609-3614); the latter defaults to 1*log(lRef) but the main process
    raises it adaptively after the initial placement (:11770-11773) -- pass the current value."""
    p = capi.SearchParams()
    p.strictTopologyStopRules = int(bool(strictTopologyStopRules))
    p.allowedFailsTopology = int(allowedFailsTopology)
    p.deeperSearchForLongBranches = int(bool(deeperSearchForLongBranches))
    p.thresholdLogLKtopology = float(thresholdLogLKtopology)
    p.thresholdTopologyPlacement = float(thresholdTopologyPlacement)
    p.thresholdLogLKoptimizationTopology = float(math.log(lRef) if thresholdLogLKoptimizationTopology is None
                                                 else thresholdLogLKoptimizationTopology)
    p.thresholdLogLKconsecutivePlacement = float(thresholdLogLKconsecutivePlacement)
    p.effectivelyNon0BLen = 1.0 / (10 * lRef)
    p.BLenThresholdDeeperSearch = (math.log(lRef) + 5) / float(lRef)
    p.defaultBLen = float(defaultBLen)
    return p


def dirty_nodes(tree: DeviceTree, dirty: Optional[Sequence[bool]] = None, replacements: Optional[Sequence[int]] = None,
                maxReplacements: int = 10) -> np.ndarray:
    """The nodes startTopologyUpdatesParallel searches (:9615-9626): reachable from the root, dirty,
    replacements <= maxReplacements, not the root.  Order: the reference's pre-order stack walk."""
    out, stack = [], [tree.root]
    while stack:
        n = stack.pop()
        if tree.child0[n] >= 0:
            stack.append(int(tree.child0[n]))
            stack.append(int(tree.child1[n]))
        if n != tree.root and (dirty is None or dirty[n]) and (replacements is None or replacements[n] <= maxReplacements):
            out.append(n)
    return np.array(out, np.int32)


def start_topology_updates_parallel(tree: DeviceTree, params: capi.SearchParams, nodes: Optional[np.ndarray] = None,
                                    scratch_keys: int = 0) -> Tuple[List[Tuple[int, int, float]], np.ndarray]:
    if nodes is None:
        nodes = dirty_nodes(tree)
    tree.prepare_search()
    rec = tree.search_records(tree.spr_search(nodes, params, scratch_keys))
    retry = np.nonzero(rec["status"] == 3)[0]
    grow = max(scratch_keys, 8192)
    while retry.size:  # per-search scratch exhausted: re-run those searches alone with 4x the scratch
        grow *= 4
        if grow > (1 << 24):
            raise capi.MapleError("SPR search scratch exhausted for %d nodes even at %d entries" % (retry.size, grow))
        again = tree.search_records(tree.spr_search(nodes[retry], params, grow, max_concurrent=max(64, (1 << 28) // grow)))
        rec[retry] = again
        retry = retry[again["status"] == 3]
    return moves_from_records(nodes, rec), rec  # improvementsFound.sort(reverse=False,key=itemgetter(2)) (:12312)
