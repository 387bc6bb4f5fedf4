"""Debug: run variant 4 (parallel + sequential replay compared per window) and dump the first mismatching window."""
import math, sys
import numpy as np, torch
sys.path.insert(0, ".")
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists
from maple_b200.search import dirty_nodes, search_params
from maple_b200.synthetic import generate
from maple_b200.tree import DeviceTree
d = generate(int(sys.argv[1]) if len(sys.argv) > 1 else 2000, rate_variation=True, seed=1, ml_like_blens=True)
eng = MapleEngine(d.model, 0)
tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, 0))
nodes = dirty_nodes(tree)
tree.prepare_search()
L = math.log(d.model.lRef)
p = search_params(d.model.lRef, False, 4, 14.0 * L)
eng.set_search_variant(4)
eng.search_stats(True, True)
out = tree.spr_search(nodes, p)
S = eng.search_stats(False, True)
D = S[32:]
print("mismatching windows:", D[0])
if D[0]:
    print("pos %d nWin %d jPar %d jSeq %d countedPar %d countedSeq %d bestPar %r bestSeq %r R %d preR %d nScore %d" % (
        D[1], D[2], D[3], D[4], D[5], D[6], np.uint64(D[7]).view(np.float64), np.uint64(D[8]).view(np.float64), D[9], D[10], D[11]))
    for q in range(int(D[2])):
        a, b = D[16 + 2 * q], D[17 + 2 * q]
        info, outw = a & 0xffffffff, a >> 32
        par, size = np.int32(np.uint32(b & 0xffffffff)), b >> 32
        print("w %2d info %s rel %d | out fail %d desc %d reached %d done %d counted %d | parent %d size %d" % (
            q, format(info & 15, "04b"), info >> 8, outw & 0xfffff, (outw >> 20) & 1, (outw >> 21) & 1, (outw >> 22) & 1, (outw >> 23) & 1, par, size))
eng.set_search_variant(2)
r2 = tree.search_records(tree.spr_search(nodes, p)).copy()
eng.set_search_variant(3)
r3 = tree.search_records(tree.spr_search(nodes, p)).copy()
bad = np.nonzero((r2["phase1"] != r3["phase1"]) | (r2["bestNode"] != r3["bestNode"]))[0]
print("variant 3 vs 2: mismatching searches", len(bad), "of", len(nodes), "status", np.bincount(r3["status"], minlength=4), np.bincount(r2["status"], minlength=4))
for i in bad[:5]:
    print("  node", nodes[i], "depth", tree.depth[nodes[i]], "phase1", r3["phase1"][i], r2["phase1"][i], "bestNode", r3["bestNode"][i], r2["bestNode"][i], "status", r3["status"][i], r2["status"][i])
for stats_on in (False, True):
    eng.set_search_variant(0)
    eng.search_stats(stats_on, False)
    r0 = tree.search_records(tree.spr_search(nodes, p)).copy()
    bad = np.nonzero((r0["phase1"] != r3["phase1"]) | (r0["bestNode"] != r3["bestNode"]))[0]
    print("stats", stats_on, "variant 0 vs 3: mismatching searches", len(bad), "of", len(nodes), "depth max", int(tree.depth.max()))
    for i in bad[:5]:
        print("  node", nodes[i], "depth", tree.depth[nodes[i]], "phase1", r0["phase1"][i], r3["phase1"][i], "bestNode", r0["bestNode"][i], r3["bestNode"][i])
