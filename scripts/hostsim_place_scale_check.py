"""Ad hoc, CPU only: the warp placement kernels (variants 1 and 3, lanes emulated in turn by tests/hostsim) on the bench's own tree
-- 100 000 sequences by default -- with the samples and stop rules of scripts/time_place.py, against the oracle.  Round 1: the first
96 samples (58 108 candidate branches each on average, tree 42 levels deep) identical for both variants with the device's default
scratch sizing.  Usage: python scripts/hostsim_place_scale_check.py [nseq] [nsamples]"""
import math
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from hostsim import KernelSourceOnHost  # noqa: E402
from maple_b200.genome_list import pack_lists  # noqa: E402
from maple_b200.synthetic import generate  # noqa: E402
from oracle.host_tree import build_tree_lists  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nsamples = int(sys.argv[2]) if len(sys.argv) > 2 else 96
d = generate(nseq, lRef=29903, mean_diffs=10.0, rate_variation=True, seed=1, ml_like_blens=True)
model = d.model
orc, hs = Oracle(model), KernelSourceOnHost(model)
lists, dist, isTip = build_tree_lists(orc, d.up, d.child0, d.child1, d.dist, d.root, d.tip_nodes, d.tip_lists, model.lRef, 0)
ta = {"up": d.up, "child0": d.child0, "child1": d.child1, "dist": dist, "isTip": isTip, "root": d.root}
lRef = model.lRef
L = math.log(lRef)
rng = np.random.default_rng(3)
new = []
for i in rng.choice(len(d.tip_lists), min(4000, len(d.tip_lists)), replace=False)[:nsamples]:
    gl, out, pos, done = d.tip_lists[i], [], 0, False
    for e in gl:
        end = e[1] if e[0] in (4, 5) else pos + 1
        if not done and e[0] == 4 and end - pos > 40:
            mid = pos + 20
            ref = int(model.refIdx[mid])
            out += [(4, mid), ((ref + 1 + int(rng.integers(3))) % 4, ref), (4, end)]
            done = True
        else:
            out.append(e)
        pos = end
    new.append(out)
samples = pack_lists(new, lRef, 0)
pp = {"strictStopRules": 0, "allowedFails": 5, "deeperSearchForLongBranches": 0, "onlyFindIdentical": 0, "thresholdLogLK": 18.0 * L,
      "thresholdLogLKoptimization": 1.0 * L, "thresholdLogLKconsecutivePlacement": 1.0, "effectivelyNon0BLen": 1.0 / (10 * lRef),
      "BLenThresholdDeeperSearch": (L + 5) / lRef, "oneMutBLen": 1.0 / lRef}
ref = orc.place_batch(ta, lists, pp, samples)
print("oracle: %.0f candidate branches per sample, status counts %s" % (ref["phase1"].mean(), np.bincount(ref["status"], minlength=4).tolist()))
for mat in (0, 2):
    got = hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=4096, mat=mat)
    ok = got["status"] != 3
    same = all(np.array_equal(got[f][ok], ref[f][ok]) for f in ("status", "bestNode", "phase1", "missedMinors", "bLenTop", "bLenBottom", "bLenAppend"))
    same = same and float(np.max(np.abs(got["bestScore"][ok] - ref["bestScore"][ok]), initial=0.0)) <= 1e-9
    print("variant %d: status counts %s, identical to the oracle where not status 3: %s" % (mat + 1, np.bincount(got["status"], minlength=4).tolist(), same))
