"""Multi-GPU plumbing of a search round (or a placement batch): which rank takes which unit, and the one exchange step.

The reference deals dirty nodes to its worker processes round-robin in pre-order (coreNum[node]==corNum, :9619,
assignCoreNumbers :12164-12195) and concatenates the workers' proposedMoves lists (:12294-12311).  Here every rank holds
the whole tree, searches nodes[rank::world] and contributes its 64-byte result records to ONE all-gather
(NCCL over NVLink on the GPU box, gloo in the CPU tests); there is no collective inside a search.  Placement batches
(findBestParentForNewSample per new sample, the reference's joblib chunks, :11280-11287) shard the same way over samples with
48-byte records.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import capi

RECORD_BYTES = 64


def shard_nodes(nodes: np.ndarray, rank: int, world: int, cost: np.ndarray = None) -> np.ndarray:
    """The nodes rank `rank` searches.  Without costs: round-robin in the given (pre-)order, the reference's
    coreNum[node]==corNum deal (:9619).  With cost[i] = how long the search of nodes[i] took last round (any unit, the same
    on every rank): nodes sorted by decreasing cost, the heavy head dealt greedily to the least loaded rank (LPT), the long tail
    of short searches in snake order; each rank's part comes out longest-first.  Every rank must pass the same
    `cost`.  shard_positions() gives the matching positions in `nodes`."""
    return nodes[shard_positions(len(nodes), rank, world, cost)]


def _owners(n: int, world: int, cost) -> np.ndarray:
    """order (positions by decreasing cost) and the rank that owns each element of it."""
    cost = np.asarray(cost, dtype=np.float64)
    order = np.argsort(-cost, kind="stable")
    owner = np.empty(n, np.int64)
    # the heavy head: greedy longest-processing-time-first (each search to the least loaded rank); search lengths have a heavy tail
    head = min(n, 256 * world)
    load = [0.0] * world
    c = cost[order[:head]]
    for i in range(head):
        r = min(range(world), key=load.__getitem__)
        owner[i] = r
        load[r] += c[i]
    # the long tail of short searches: snake deal, starting with the least loaded rank
    if head < n:
        by_load = np.argsort(np.asarray(load), kind="stable")
        k = np.arange(n - head)
        rnd, col = k // world, k % world
        owner[head:] = by_load[np.where(rnd % 2 == 0, col, world - 1 - col)]
    return order, owner


def shard_positions(n: int, rank: int, world: int, cost: np.ndarray = None) -> np.ndarray:
    if world <= 1 and cost is None:
        return np.arange(n)
    if cost is None:
        return np.arange(rank, n, world)
    order, owner = _owners(n, world, cost)
    return order[owner == rank]


def shard_sizes(n: int, world: int, cost: np.ndarray = None) -> np.ndarray:
    if cost is None:
        return np.array([len(range(r, n, world)) for r in range(world)])
    return np.bincount(_owners(n, world, cost)[1], minlength=world)


def all_gather_raw(raw: torch.Tensor, n_total: int, world: int, per_rank: int = 0) -> torch.Tensor:
    """The exchange step: every rank contributes its records (padded to per_rank rows, default ceil(n_total/world)) to one
    all-gather and gets [world * per_rank, 64] uint8 back, on the device `raw` lives on."""
    per_rank = per_rank or (n_total + world - 1) // world
    width = raw.shape[1]
    pad = torch.zeros((per_rank, width), dtype=torch.uint8, device=raw.device)
    pad[: raw.shape[0]] = raw
    out = torch.empty((world * per_rank, width), dtype=torch.uint8, device=raw.device)
    dist.all_gather_into_tensor(out, pad)
    return out


def gather_rows(raw: torch.Tensor, n_total: int, rank: int, world: int, cost: np.ndarray = None) -> np.ndarray:
    """raw: this rank's rows [n_mine, width] uint8, in the order of shard_nodes(nodes, rank, world, cost).  Returns all n_total
    rows [n_total, width] in the order of the un-sharded list (one all-gather when world > 1)."""
    if world == 1 and cost is None:
        return raw.cpu().numpy()
    width = raw.shape[1]
    full = np.empty((n_total, width), dtype=np.uint8)
    if world == 1:
        full[shard_positions(n_total, 0, 1, cost)] = raw.cpu().numpy()
        return full
    per_rank = int(shard_sizes(n_total, world, cost).max())
    rec = all_gather_raw(raw, n_total, world, per_rank).cpu().numpy().reshape(world, per_rank, width)
    for r in range(world):
        pos = shard_positions(n_total, r, world, cost)
        full[pos] = rec[r, : len(pos)]
    return full


def gather_records(raw: torch.Tensor, n_total: int, rank: int, world: int, fields=None, cost: np.ndarray = None) -> np.ndarray:
    """raw: this rank's records [n_mine, record bytes] uint8 (on the device the search ran on, or on the CPU).  Returns all
    n_total records as a numpy record array (`fields`: capi.SEARCH_RESULT_FIELDS by default, capi.PLACE_RESULT_FIELDS for
    placement batches), in the order of the un-sharded node / sample list.  `cost`: what shard_nodes was called with."""
    dt = np.dtype(capi.SEARCH_RESULT_FIELDS if fields is None else fields)
    assert raw.shape[1] == dt.itemsize
    return gather_rows(raw, n_total, rank, world, cost).view(dt).reshape(-1)


def moves_from_records(nodes: np.ndarray, rec: np.ndarray) -> List[Tuple[int, int, float]]:
    """proposedMoves = [(node, placementNode, improvement)], ascending by improvement (:12312)."""
    moves = [(int(n), int(r["placement"]), float(r["improvement"])) for n, r in zip(nodes, rec) if r["placement"] >= 0]
    moves.sort(key=lambda m: m[2])
    return moves
