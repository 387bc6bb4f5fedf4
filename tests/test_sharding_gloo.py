"""Multi-rank host logic of a search round with world size 2 on gloo, no GPU: the node partition is the
reference's coreNum[node]==corNum round-robin (:9619, :12164-12195), the exchange is ONE all-gather of fixed-size result
records, and every rank ends up with the same proposedMoves list, sorted as :12312 does.  The per-rank searches are
replaced by the CPU oracle here (this is a test: tests/ may use the oracle); on the GPU box the same functions carry the
records the CUDA search wrote."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maple_b200 import capi
from maple_b200.sharding import gather_records, moves_from_records, shard_nodes


def _problem():
    """A small synthetic tree with its four list families (built with the oracle: no GPU here) and deep-round parameters."""
    import math
    from host_recalc import recalc_lists
    from maple_b200.genome_list import pack_lists
    from maple_b200.synthetic import generate
    from oracle.oracle import Oracle
    d = generate(120, lRef=3000, mean_diffs=8.0, rate_variation=True, seed=7)
    orc = Oracle(d.model)
    n = len(d.up)
    children = [[int(d.child0[i]), int(d.child1[i])] if d.child0[i] >= 0 else [] for i in range(n)]
    upl = [None if u < 0 else int(u) for u in d.up]
    isTip = [not children[i] for i in range(n)]
    lower, upR, upL, tot = recalc_lists(orc, upl, children, [float(x) for x in d.dist], [[] for _ in range(n)], isTip, d.root,
                                        {int(t): d.tip_lists[i] for i, t in enumerate(d.tip_nodes)})
    lists = [lower[i] for i in range(n)] + [upR.get(i) for i in range(n)] + [upL.get(i) for i in range(n)] + [tot.get(i) for i in range(n)]
    packed = pack_lists(lists, d.model.lRef, d.model.usingErrorRate)
    ta = {"up": d.up, "child0": d.child0, "child1": d.child1, "dist": d.dist, "isTip": np.array(isTip, np.uint8), "root": d.root}
    L = math.log(d.model.lRef)
    params = {"strictTopologyStopRules": 0, "allowedFailsTopology": 4, "deeperSearchForLongBranches": 0, "thresholdLogLKtopology": 14.0 * L,
              "thresholdTopologyPlacement": -0.1, "thresholdLogLKoptimizationTopology": L, "thresholdLogLKconsecutivePlacement": 1.0,
              "effectivelyNon0BLen": 1.0 / (10 * d.model.lRef), "BLenThresholdDeeperSearch": (L + 5) / d.model.lRef, "defaultBLen": 0.000033}
    nodes = np.array([i for i in range(n) if d.up[i] >= 0], np.int32)
    return d.model, ta, packed, params, nodes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        model, ta, packed, params, nodes = _problem()
        mine = shard_nodes(nodes, rank, world)
        rec = Oracle(model).search_batch(ta, packed, params, mine, lazy_mode=1)
        raw = torch.from_numpy(np.ascontiguousarray(rec).view(np.uint8).reshape(len(mine), 64).copy())
        allrec = gather_records(raw, len(nodes), rank, world)
        moves = moves_from_records(nodes, allrec)
        # second round on the same tree: cost-balanced shards from what the first round measured (here the candidate counts)
        cost = allrec["phase1"].astype(np.float64)
        mine2 = shard_nodes(nodes, rank, world, cost)
        rec2 = Oracle(model).search_batch(ta, packed, params, mine2, lazy_mode=1)
        raw2 = torch.from_numpy(np.ascontiguousarray(rec2).view(np.uint8).reshape(len(mine2), 64).copy())
        allrec2 = gather_records(raw2, len(nodes), rank, world, cost=cost)
        assert allrec2.tobytes() == allrec.tobytes()
        q.put((rank, moves, allrec.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_agree_with_one():
    from oracle.oracle import Oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    model, ta, packed, params, nodes = _problem()
    one = Oracle(model).search_batch(ta, packed, params, nodes, lazy_mode=1)
    ref_moves = moves_from_records(nodes, one)
    assert got[0][1] == got[1][1] == ref_moves and len(ref_moves) > 0
    assert got[0][2] == got[1][2] == np.ascontiguousarray(one).tobytes()


def _place_problem():
    from maple_b200.genome_list import pack_lists
    model, ta, packed, params, nodes = _problem()
    from maple_b200.synthetic import generate
    d = generate(120, lRef=3000, mean_diffs=8.0, rate_variation=True, seed=7)
    samples = []
    for i, v in enumerate(d.tip_lists[:37]):  # new samples: placed tips with one more substitution (every third one unchanged)
        v = [tuple(e) for e in v]
        if i % 3:
            prev = 0
            for j, e in enumerate(v):
                end = e[1] if e[0] in (4, 5) else prev + 1
                if e[0] == 4 and len(e) == 2 and end - prev >= 3 + i:
                    pos = prev + 2 + i  # 1-based position of the new substitution, inside this R run
                    ref = int(d.model.refIdx[pos - 1])
                    v[j:j + 1] = [(4, pos - 1), ((ref + 1 + i % 3) % 4, ref), (4, end)]
                    break
                prev = end
        samples.append(v)
    pp = {"strictStopRules": 0, "allowedFails": 3, "deeperSearchForLongBranches": 0, "onlyFindIdentical": 0,
          "thresholdLogLK": params["thresholdLogLKtopology"], "thresholdLogLKoptimization": params["thresholdLogLKoptimizationTopology"],
          "thresholdLogLKconsecutivePlacement": 0.01, "effectivelyNon0BLen": params["effectivelyNon0BLen"],
          "BLenThresholdDeeperSearch": params["BLenThresholdDeeperSearch"], "oneMutBLen": 1.0 / d.model.lRef}
    return model, ta, packed, pp, samples


def _place_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from maple_b200.genome_list import pack_lists
        from oracle.oracle import Oracle
        model, ta, packed, pp, samples = _place_problem()
        mine = shard_nodes(np.arange(len(samples)), rank, world)
        rec = Oracle(model).place_batch(ta, packed, pp, pack_lists([samples[i] for i in mine], model.lRef, model.usingErrorRate))
        raw = torch.from_numpy(np.ascontiguousarray(rec).view(np.uint8).reshape(len(mine), 48).copy())
        allrec = gather_records(raw, len(samples), rank, world, fields=capi.PLACE_RESULT_FIELDS)
        q.put((rank, allrec.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_placement_batch_two_ranks_agree_with_one():
    """Samples of a placement batch are the sharded unit; same exchange, 48-byte records."""
    from maple_b200.genome_list import pack_lists
    from oracle.oracle import Oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_place_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    model, ta, packed, pp, samples = _place_problem()
    one = Oracle(model).place_batch(ta, packed, pp, pack_lists(samples, model.lRef, model.usingErrorRate))
    assert got[0][1] == got[1][1] == np.ascontiguousarray(one).tobytes()
    assert (one["status"] == 0).sum() > 10 and (one["status"] == 1).sum() > 5 and (one["phase1"][one["status"] == 0] > 0).all()


def test_cost_balanced_shards():
    rng = np.random.default_rng(3)
    nodes = rng.permutation(1000).astype(np.int32)
    cost = rng.pareto(1.2, 1000) * 1e6  # heavy tail, like search lengths
    for world in (1, 2, 4, 8):
        parts = [shard_nodes(nodes, r, world, cost) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == sorted(nodes.tolist())
        c = dict(zip(nodes.tolist(), cost.tolist()))
        sums = [sum(c[int(x)] for x in p) for p in parts]
        # the single longest search bounds what any deal can do; beyond it the shares are within a few per cent
        assert max(sums) <= max(cost.max(), 1.02 * sum(sums) / world), (world, sums)
        for p in parts:  # longest first inside a shard
            cc = [c[int(x)] for x in p]
            assert cc == sorted(cc, reverse=True)


def test_shard_nodes_is_the_round_robin_partition():
    nodes = np.arange(11, dtype=np.int32) * 3
    parts = [shard_nodes(nodes, r, 4) for r in range(4)]
    assert sorted(np.concatenate(parts).tolist()) == nodes.tolist()
    assert [len(p) for p in parts] == [3, 3, 3, 2]
    assert parts[1].tolist() == nodes[1::4].tolist()
    assert np.dtype(capi.SEARCH_RESULT_FIELDS).itemsize == 64
