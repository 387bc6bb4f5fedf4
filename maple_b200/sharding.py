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


def shard_nodes(nodes: np.ndarray, rank: int, world: int) -> np.ndarray:
    return nodes[rank::world] if world > 1 else nodes


def all_gather_raw(raw: torch.Tensor, n_total: int, world: int) -> torch.Tensor:
    """The exchange step: every rank contributes its records (padded to ceil(n_total/world) rows) to one all-gather and
    gets [world * per_rank, 64] uint8 back, on the device `raw` lives on."""
    per_rank = (n_total + world - 1) // world
    width = raw.shape[1]
    pad = torch.zeros((per_rank, width), dtype=torch.uint8, device=raw.device)
    pad[: raw.shape[0]] = raw
    out = torch.empty((world * per_rank, width), dtype=torch.uint8, device=raw.device)
    dist.all_gather_into_tensor(out, pad)
    return out


def gather_records(raw: torch.Tensor, n_total: int, rank: int, world: int, fields=None) -> np.ndarray:
    """raw: this rank's records [n_mine, record bytes] uint8 (on the device the search ran on, or on the CPU).  Returns all
    n_total records as a numpy record array (`fields`: capi.SEARCH_RESULT_FIELDS by default, capi.PLACE_RESULT_FIELDS for
    placement batches), in the order of the un-sharded node / sample list."""
    dt = np.dtype(capi.SEARCH_RESULT_FIELDS if fields is None else fields)
    assert raw.shape[1] == dt.itemsize
    if world == 1:
        return raw.cpu().numpy().view(dt).reshape(-1)
    per_rank = (n_total + world - 1) // world
    rec = all_gather_raw(raw, n_total, world).cpu().numpy().view(dt).reshape(world, per_rank)
    full = np.empty(n_total, dtype=rec.dtype)
    for r in range(world):  # rank r holds nodes r, r+world, ...
        k = len(range(r, n_total, world))
        full[r::world] = rec[r, :k]
    return full


def moves_from_records(nodes: np.ndarray, rec: np.ndarray) -> List[Tuple[int, int, float]]:
    """proposedMoves = [(node, placementNode, improvement)], ascending by improvement (:12312)."""
    moves = [(int(n), int(r["placement"]), float(r["improvement"])) for n, r in zip(nodes, rec) if r["placement"] >= 0]
    moves.sort(key=lambda m: m[2])
    return moves
