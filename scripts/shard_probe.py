"""Ad-hoc: what one GPU of an N-GPU round sees -- a cost-balanced 1/N shard of the searches -- timed for several settings of
searches per warp."""
import math, sys
import numpy as np, torch
sys.path.insert(0, ".")
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists
from maple_b200.search import dirty_nodes, search_params
from maple_b200.sharding import shard_nodes
from maple_b200.synthetic import generate
from maple_b200.tree import DeviceTree

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lanes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 4, 8]
d = generate(nseq, rate_variation=True, seed=1, ml_like_blens=True)
eng = MapleEngine(d.model, 0)
tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, 0))
nodes = dirty_nodes(tree)
tree.prepare_search()
L = math.log(d.model.lRef)
p = search_params(d.model.lRef, False, 4, 14.0 * L)
cyc = torch.zeros(len(nodes), dtype=torch.int64, device=eng.device)
full = tree.search_records(tree.spr_search(nodes, p, cycles=cyc))
cost = cyc.cpu().numpy().astype(np.float64)
mine = shard_nodes(nodes, 0, world, cost)
print("full round: %d searches; shard 0 of %d: %d nodes, %d real searches, cost share %.3f" % (
    (full["status"] == 0).sum(), world, len(mine), (full["status"][np.isin(nodes, mine)] == 0).sum(), cost[np.isin(nodes, mine)].sum() / cost.sum()), flush=True)
for l in lanes:
    eng.set_lanes_per_warp(l)
    ts = []
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c2 = torch.zeros(len(mine), dtype=torch.int64, device=eng.device)
        a.record()
        out = tree.spr_search(mine, p, cycles=c2)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    c = c2.cpu().numpy()
    print("lanes/warp %2d: %s ms | longest search %.1f ms, p99 %.1f ms, sum of search times / 2368 warps = %.1f ms" % (
        l, " ".join("%.1f" % t for t in ts), c.max() / 1.965e6, np.percentile(c, 99) / 1.965e6, c.sum() / 1.965e6 / 2368), flush=True)
# the longest searches on SMs of their own (maple_ctx_set_critical_searches): explicit counts, then the automatic choice
eng.set_lanes_per_warp(0)
for k in ([int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 4, 8, 16, 32, -1]):
    ts = []
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c2 = torch.zeros(len(mine), dtype=torch.int64, device=eng.device)
        a.record()
        out = tree.spr_search(mine, p, cycles=c2, critical=None if k < 0 else k)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    rec = tree.search_records(out)
    pos = {int(nd): i for i, nd in enumerate(nodes)}
    sel = np.array([pos[int(nd)] for nd in mine])
    same = all(np.array_equal(rec[f], full[f][sel]) for f in ("status", "placement", "phase1", "bLenAppend"))
    c = c2.cpu().numpy()
    print("critical %3s (set %d): %s ms | longest search %.1f ms | records equal to the full round's: %s" % (
        "auto" if k < 0 else k, getattr(tree, "_critical_set", 0), " ".join("%.1f" % t for t in ts), c.max() / 1.965e6, same), flush=True)
# head of the list handed out one per warp (maple_ctx_set_head_searches)
for h in ([int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else []):
    eng.set_head_searches(h)
    ts = []
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c2 = torch.zeros(len(mine), dtype=torch.int64, device=eng.device)
        a.record()
        out = tree.spr_search(mine, p, cycles=c2, critical=0)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    rec = tree.search_records(out)
    pos = {int(nd): i for i, nd in enumerate(nodes)}
    sel = np.array([pos[int(nd)] for nd in mine])
    same = all(np.array_equal(rec[f], full[f][sel]) for f in ("status", "placement", "phase1", "bLenAppend"))
    print("head %6d: %s ms | records equal to the full round's: %s" % (h, " ".join("%.1f" % t for t in ts), same), flush=True)
eng.set_head_searches(0)
