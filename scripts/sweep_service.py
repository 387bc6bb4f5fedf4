"""Ad-hoc: sweep the scan-service split (SMs that own searches vs SMs that serve scans) on a synthetic tree; records must not change."""
import math, sys
import numpy as np, torch
sys.path.insert(0, ".")
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists
from maple_b200.search import dirty_nodes, search_params
from maple_b200.synthetic import generate
from maple_b200.tree import DeviceTree

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
settings = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 12, 24, 36, 48, 74, 100]  # SMs that own searches; 0 = no service
rounds = sys.argv[3].split(",") if len(sys.argv) > 3 else ["deep", "fast"]
d = generate(nseq, rate_variation=True, seed=1, ml_like_blens=True)
eng = MapleEngine(d.model, 0)
tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, 0))
nodes = dirty_nodes(tree)
tree.prepare_search()
L = math.log(d.model.lRef)
for rnd in rounds:
    p = search_params(d.model.lRef, True, 2, 6.0 * L) if rnd == "fast" else search_params(d.model.lRef, False, 4, 14.0 * L)
    base = None
    for fsm in settings:
        eng.set_scan_service(fsm)
        times = []
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = tree.spr_search(nodes, p)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        rec = tree.search_records(out)
        raw = rec.tobytes()
        if base is None:
            base = raw
        print("%s fsmSMs %3d: %s ms, phase1 %d, %.3g cand/s, status %s, same as first setting: %s" % (
            rnd, fsm, " ".join("%.1f" % t for t in times), rec["phase1"].sum(), rec["phase1"].sum() / min(times[1:]) * 1e3,
            np.bincount(rec["status"], minlength=4).tolist(), raw == base), flush=True)
