"""Ad hoc: the head of the sorted list handed out one search per warp (maple_ctx_set_head_searches) against the plain launch, on a
whole round of a big tree.  usage: head_probe.py <nseq> <fast|deep> <head counts, comma separated>"""
import math, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists
from maple_b200.search import dirty_nodes, search_params
from maple_b200.synthetic import generate
from maple_b200.tree import DeviceTree

nseq = int(sys.argv[1])
deep = sys.argv[2] == "deep"
heads = [int(x) for x in sys.argv[3].split(",")]
t0 = time.time()
d = generate(nseq, rate_variation=True, seed=1, ml_like_blens=True)
eng = MapleEngine(d.model, 0)
tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, 0))
nodes = dirty_nodes(tree)
tree.prepare_search()
L = math.log(d.model.lRef)
p = search_params(d.model.lRef, False, 4, 14.0 * L) if deep else search_params(d.model.lRef, True, 2, 6.0 * L)
print("setup %.1fs, %d nodes" % (time.time() - t0, tree.n), flush=True)
ref = None
for h in [-1] + heads:  # -1: the first, unsorted round (it measures the lengths)
    eng.set_head_searches(max(h, 0))
    eng.search_stats(True, True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = tree.spr_search(nodes, p, critical=0)
    b.record()
    torch.cuda.synchronize()
    S = eng.search_stats(False, True)
    ms = a.elapsed_time(b)
    rec = tree.search_records(out)
    if ref is None:
        ref = rec
    same = all(np.array_equal(rec[f], ref[f]) for f in ("status", "placement", "phase1", "bLenAppend"))
    busy = float(sum(S[0:6])) / 1.9e9 / (2368 * ms / 1e3)
    print("head %8d: %9.1f ms, %d candidates, warps busy %.0f%%, records equal: %s" % (h, ms, int(rec["phase1"].sum()), 100 * busy, same), flush=True)
