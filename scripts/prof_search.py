"""Profiling target: build a synthetic tree and run the SPR search kernel N times (for ncu -k regex:k_spr_search)."""
import math, sys
import numpy as np, torch
sys.path.insert(0, ".")
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists
from maple_b200.search import dirty_nodes, search_params
from maple_b200.synthetic import generate
from maple_b200.tree import DeviceTree

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rnd = sys.argv[2] if len(sys.argv) > 2 else "deep"
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
scratch_keys = int(sys.argv[5]) if len(sys.argv) > 5 else 16384
d = generate(nseq, rate_variation=True, seed=1, ml_like_blens=True)
eng = MapleEngine(d.model, 0)
eng.set_search_variant(variant)
if len(sys.argv) > 6:
    eng.set_scan_service(int(sys.argv[6]))
tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, 0))
nodes = dirty_nodes(tree)
tree.prepare_search()
L = math.log(d.model.lRef)
p = search_params(d.model.lRef, True, 2, 6.0 * L) if rnd == "fast" else search_params(d.model.lRef, False, 4, 14.0 * L)
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = tree.spr_search(nodes, p, scratch_keys=scratch_keys)
    b.record()
    torch.cuda.synchronize()
    rec = tree.search_records(out)
    print("%s v%d nseq %d: %.1f ms, phase1 %d, %.3g cand/s, status %s" % (rnd, variant, nseq, a.elapsed_time(b), rec["phase1"].sum(), rec["phase1"].sum() / a.elapsed_time(b) * 1e3, np.bincount(rec["status"], minlength=4).tolist()), flush=True)
