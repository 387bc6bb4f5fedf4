"""Ad hoc: throughput of maple_place_batch (findBestParentForNewSample on a frozen tree) on a synthetic tree.
New samples = existing tips with one extra substitution each (so they are not simply absorbed as minor sequences)."""
import json, math, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from maple_b200 import capi
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists
from maple_b200.synthetic import generate
from maple_b200.tree import DeviceTree

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nnew = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]  # 0 = sample per thread, 1 = sample per warp, 2 = + MAT trees, 3 = + parallel window replay
d = generate(nseq, rate_variation=True, seed=1, ml_like_blens=True)
eng = MapleEngine(d.model, 0)
tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, 0))
tree.prepare_search()
lRef = d.model.lRef
rng = np.random.default_rng(3)
new = []
for i in rng.choice(len(d.tip_lists), nnew, replace=False):
    gl, out, pos, done = d.tip_lists[i], [], 0, False
    for e in gl:  # split the first long R run to insert a substitution in its middle
        end = e[1] if e[0] in (4, 5) else pos + 1
        if not done and e[0] == 4 and end - pos > 40:
            mid = pos + 20
            ref = int(d.model.refIdx[mid])
            out += [(4, mid), ((ref + 1 + int(rng.integers(3))) % 4, ref), (4, end)]
            done = True
        else:
            out.append(e)
        pos = end
    new.append(out)
samples = pack_lists(new, lRef, 0)
L = math.log(lRef)
p = capi.PlaceParams()
p.strictStopRules, p.allowedFails, p.deeperSearchForLongBranches, p.onlyFindIdentical = 0, 5, 0, 0
p.thresholdLogLK, p.thresholdLogLKoptimization, p.thresholdLogLKconsecutivePlacement = 18.0 * L, 1.0 * L, 1.0
p.effectivelyNon0BLen, p.BLenThresholdDeeperSearch, p.oneMutBLen = 1.0 / (10 * lRef), (L + 5) / lRef, 1.0 / lRef
results, summary = {}, {"nseq": nseq, "new_samples": nnew, "variants": {}}
for variant in variants:
    eng.set_place_variant(variant)
    best_dt = None
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        rec = tree.place_samples(samples, p)
        dt = time.time() - t0
        print("variant %d rep %d: %d samples in %.2f s (incl. upload and re-bind) = %.3g samples/s; %d candidate branches (%.0f / sample) = "
              "%.3g placements/s; status %s" % (variant, rep, nnew, dt, nnew / dt, rec["phase1"].sum(), rec["phase1"].mean(),
                                                rec["phase1"].sum() / dt, np.bincount(rec["status"], minlength=4).tolist()), flush=True)
        best_dt = dt if best_dt is None else min(best_dt, dt)
    results[variant] = rec
    summary["variants"][str(variant)] = {"seconds": best_dt, "samples_per_s": nnew / best_dt, "candidate_branches": int(rec["phase1"].sum()),
                                         "placements_per_s": float(rec["phase1"].sum()) / best_dt,
                                         "status_counts": np.bincount(rec["status"], minlength=4).tolist()}
for v in variants[1:]:  # the kernels must agree record by record
    a, b = results[variants[0]], results[v]
    same = all(np.array_equal(a[f], b[f]) for f in ("bestNode", "status", "phase1", "missedMinors", "bLenTop", "bLenBottom", "bLenAppend"))
    print("variants %d and %d agree:" % (variants[0], v), same, "max |score diff| =", float(np.nanmax(np.abs(a["bestScore"] - b["bestScore"]))))
    summary["variants"][str(v)]["records_identical_to_variant_%d" % variants[0]] = bool(same)
summary["timing"] = "wall clock around DeviceTree.place_samples: upload of the sample lists, re-bind, kernel, records back; best of 2"
print("PLACE_JSON " + json.dumps(summary), flush=True)
