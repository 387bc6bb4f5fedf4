"""Ad hoc: where does the longest search of a deep round spend its time when it runs alone?"""
import math, sys
import numpy as np, torch
sys.path.insert(0, ".")
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists
from maple_b200.search import dirty_nodes, search_params
from maple_b200.synthetic import generate
from maple_b200.tree import DeviceTree
nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
d = generate(nseq, rate_variation=True, seed=1, ml_like_blens=True)
eng = MapleEngine(d.model, 0)
tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, 0))
nodes = dirty_nodes(tree)
tree.prepare_search()
L = math.log(d.model.lRef)
p = search_params(d.model.lRef, False, 4, 14.0 * L)
cyc = torch.zeros(len(nodes), dtype=torch.int64, device=eng.device)
rec = tree.search_records(tree.spr_search(nodes, p, scratch_keys=16384, cycles=cyc))
c = cyc.cpu().numpy()
order = np.argsort(-c)[:8]
print("longest searches (in the full round):")
for i in order:
    nd = nodes[i]
    print("  node %d depth %d subtree? phase1 %d cycles %.3g (%.0f ms)" % (nd, tree.depth[nd], rec["phase1"][i], c[i], c[i] / 1.965e6))
for i in order[:3]:
    one = np.array([nodes[i]], np.int32)
    eng.search_stats(True, True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r1 = tree.search_records(tree.spr_search(one, p, scratch_keys=16384))
    b.record(); torch.cuda.synchronize()
    S = eng.search_stats(False, True)
    wc = float(sum(S[0:6])) or 1.0
    print("node %d alone: %.1f ms, phase1 %d | control %.1f%% append %.1f%% merge %.1f%% blen %.1f%% differ %.1f%% scan %.1f%% (window+stage %.1f%% score %.1f%% replay %.1f%%) | ops a %d m %d b %d d %d, scan jobs %d batches %d counted %d queued %d" % (
        nodes[i], a.elapsed_time(b), r1["phase1"][0], 100 * S[0] / wc, 100 * S[1] / wc, 100 * S[2] / wc, 100 * S[3] / wc, 100 * S[4] / wc, 100 * S[5] / wc,
        100 * S[23] / wc, 100 * S[6] / wc, 100 * S[7] / wc, S[8], S[9], S[10], S[11], S[17], S[19], S[21], S[22]), flush=True)
