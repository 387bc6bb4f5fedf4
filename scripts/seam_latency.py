"""The seam as the reference uses it outside the big parallel round:
 (a) latency of maple_spr_search_batch for batches of 1, 8, 64, 512 searches -- the reference's sub-rounds (:12348, :12374) and
     applySPRMovesParallel (:9476-9483) call the search for one node or a few at a time;
 (b) what INTEGRATION.md's gpuTopologyProposals pays per round when the tree lives in python: packing the reference's tuple lists of
     every node (pack_lists), uploading them (DeviceTree.from_lists) and binding (prepare_search), against keeping the tree resident.
Prints one JSON object."""
import json, math, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists, decode_stream
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
out = {"nseq": nseq, "nodes": int(tree.n)}
rng = np.random.default_rng(5)
for name, p in (("deep", search_params(d.model.lRef, False, 4, 14.0 * L)), ("fast", search_params(d.model.lRef, True, 2, 6.0 * L))):
    full = tree.search_records(tree.spr_search(nodes, p))
    real = nodes[full["status"] == 0]
    res = {}
    for k in (1, 8, 64, 512):
        lat = []
        for rep in range(40 if k <= 64 else 12):
            sel = rng.choice(real, k, replace=False).astype(np.int32)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            o = tree.spr_search(sel, p, schedule=False)
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.array(lat[2:])
        res["batch_%d" % k] = {"ms_p50": float(np.percentile(lat, 50)), "ms_p99": float(np.percentile(lat, 99)), "ms_max": float(lat.max())}
    out["search_latency_" + name] = res
# (b) the tree as the reference holds it: python tuple lists per node
host = tree.arena.to_host()
t0 = time.perf_counter()
py_lists = []
for i in range(4 * tree.n):
    ks = int(host.key_start[i])
    py_lists.append(None if ks < 0 else decode_stream(host.key, host.pay, ks, int(host.pay_start[i]), host.lRef, host.U, int(host.nkeys[i])))
t_decode = time.perf_counter() - t0
t0 = time.perf_counter()
packed = pack_lists(py_lists, d.model.lRef, 0)
t_pack = time.perf_counter() - t0
t0 = time.perf_counter()
t2 = DeviceTree.from_lists(eng, tree.up, tree.child0, tree.child1, tree.dist, tree.root, tree.isTip, packed)
torch.cuda.synchronize()
t_upload = time.perf_counter() - t0
t0 = time.perf_counter()
t2.prepare_search()
torch.cuda.synchronize()
t_bind = time.perf_counter() - t0
out["python_tree_to_device"] = {"lists": 4 * int(tree.n), "pack_lists_s": t_pack, "upload_from_lists_s": t_upload, "tree_bind_s": t_bind,
                                "packed_MB": (packed.key.nbytes + packed.pay.nbytes) / 1e6,
                                "note": "pack_lists walks every tuple in python; a resident DeviceTree pays none of this per round"}
print("SEAM_JSON " + json.dumps(out))
