"""Ad-hoc: effect of resident-lane count and search ordering on the FSM search kernel."""
import math, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists
from maple_b200.search import dirty_nodes, search_params
from maple_b200.synthetic import generate
from maple_b200.tree import DeviceTree

nseq = int(sys.argv[1])
d = generate(nseq, rate_variation=True, seed=1, ml_like_blens=True)
eng = MapleEngine(d.model, 0)
tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, 0))
nodes = dirty_nodes(tree)
tree.prepare_search()
L = math.log(d.model.lRef)
pf = search_params(d.model.lRef, True, 2, 6.0 * L)
pd = search_params(d.model.lRef, False, 4, 14.0 * L)

def run(p, nd, conc):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = tree.spr_search(nd, p, max_concurrent=conc)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), tree.search_records(out)

ms, rec = run(pf, nodes, 0)
print("fast natural order, all lanes: %.1f ms  %.3g cand/s" % (ms, rec["phase1"].sum() / ms * 1e3), flush=True)
order_f = np.argsort(-rec["phase1"], kind="stable")
for conc in (0, 32768, 16384, 8192, 4096):
    ms2, _ = run(pf, nodes[order_f], conc)
    print("fast LPT(self) conc %6d: %.1f ms  %.3g cand/s" % (conc, ms2, rec["phase1"].sum() / ms2 * 1e3), flush=True)
for conc in (0, 32768, 16384, 8192):
    ms2, _ = run(pf, nodes, conc)
    print("fast natural conc %6d: %.1f ms  %.3g cand/s" % (conc, ms2, rec["phase1"].sum() / ms2 * 1e3), flush=True)
if len(sys.argv) > 2:
    for conc in (0, 16384, 8192):
        ms3, rd = run(pd, nodes[order_f], conc)
        print("deep LPT(fast) conc %6d: %.1f ms  %.3g cand/s" % (conc, ms3, rd["phase1"].sum() / ms3 * 1e3), flush=True)
