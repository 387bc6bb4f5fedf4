"""Ad-hoc timing of the device SPR search on a synthetic tree (not the bench)."""
import math, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import os
from maple_b200 import capi
if os.environ.get('MAPLE_LIB'):  # A/B of two builds of the library in one job
    capi.LIB_PATH = os.path.abspath(os.environ['MAPLE_LIB'])
from maple_b200.engine import MapleEngine
from maple_b200.genome_list import pack_lists
from maple_b200.search import dirty_nodes, search_params
from maple_b200.synthetic import generate
from maple_b200.tree import DeviceTree

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
t0 = time.time()
ml = len(sys.argv) > 2 and sys.argv[2] == 'ml'
d = generate(nseq, rate_variation=True, seed=1, ml_like_blens=ml)
eng = MapleEngine(d.model, 0)
tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, 0))
torch.cuda.synchronize()
print("setup %.1fs nodes %d arena %.1f MB" % (time.time() - t0, tree.n, tree.arena.used_bytes() / 1e6), flush=True)
nodes = dirty_nodes(tree)
tree.prepare_search()
L = math.log(d.model.lRef)
variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
if len(sys.argv) > 5:
    eng.set_scan_service(int(sys.argv[5]))
rounds = sys.argv[4].split(",") if len(sys.argv) > 4 else ["fast", "deep"]
for variant in variants:
  eng.set_search_variant(variant)
  for name, strict, fails, thr in (("fast", True, 2, 6.0 * L), ("deep", False, 4, 14.0 * L)):
    if name not in rounds:
        continue
    name = "v%d %s" % (variant, name)
    p = search_params(d.model.lRef, strict, fails, thr)
    for rep in range(2):
        if rep == 1 and variant != 1:
            eng.search_stats(True, True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        cyc = torch.zeros(len(nodes), dtype=torch.int64, device=eng.device)
        out = tree.spr_search(nodes, p, cycles=cyc)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
    rec = tree.search_records(out)
    if variant != 1:
        S = eng.search_stats(False, True)
        wc = float(sum(S[0:6])) or 1.0
        print("   total warp time in the loop: %.1f warp-seconds @1.9GHz = %.0f%% of 2368 warps x %.1f ms" % (wc / 1.9e9, 100 * wc / 1.9e9 / (2368 * ms / 1e3), ms), flush=True)
        print("   warp cycles: control %.1f%% append %.1f%% merge %.1f%% blen %.1f%% differ %.1f%% scan %.1f%% (scan: window+stage %.1f%% score %.1f%% replay %.1f%%) | "
              "iterations %.3g; lanes/op-iteration: append %.1f merge %.1f blen %.1f differ %.1f | op counts a %.3g m %.3g b %.3g d %.3g" % (
                  100 * S[0] / wc, 100 * S[1] / wc, 100 * S[2] / wc, 100 * S[3] / wc, 100 * S[4] / wc, 100 * S[5] / wc, 100 * S[23] / wc, 100 * S[6] / wc, 100 * S[7] / wc,
                  S[16], S[8] / max(S[12], 1), S[9] / max(S[13], 1), S[10] / max(S[14], 1), S[11] / max(S[15], 1), S[8], S[9], S[10], S[11]), flush=True)
        print("   scan jobs %d, nodes in their ranges %.3g, batches %.3g, lanes scored %.3g (%.1f / batch, window %.1f nodes), counted %.3g, phase-2 entries queued %d, windows replayed node by node %d" % (
            S[17], S[18], S[19], S[20], S[20] / max(S[19], 1), S[24] / max(S[19], 1), S[21], S[22], S[25]), flush=True)
    if variant != 1 and len(S) > 33:
        print("   queued phase-2 entries evaluated by the warp: %d, left to the owning lane: %d" % (S[32], S[33]), flush=True)
    if variant != 1 and S[27]:
        print("   scan service: jobs posted (lane 0 only) %d, served %d, server warps serving %.1f warp-s, idle %.1f warp-s" % (S[26], S[27], S[28] / 1.9e9, S[29] / 1.9e9), flush=True)
    st = np.bincount(rec["status"], minlength=4)
    ph = rec["phase1"].sum()
    print("%s: %.1f ms, searches %d, status %s, phase1 %d (%.1f/search, max %d), %.3g cand/s, proposals %d" % (
        name, ms, len(nodes), st.tolist(), ph, ph / len(nodes), rec["phase1"].max(), ph / ms * 1e3, (rec["placement"] >= 0).sum()), flush=True)
    c = cyc.cpu().numpy().astype(np.float64)
    print("   cycles/search: mean %.3g p50 %.3g p99 %.3g max %.3g | sum/56832 threads = %.1f ms @1.9GHz, longest = %.1f ms | cycles per candidate %.0f | corr(phase1,cycles)=%.2f" % (
        c.mean(), np.percentile(c, 50), np.percentile(c, 99), c.max(), c.sum() / 56832 / 1.9e6, c.max() / 1.9e6, c.sum() / max(ph, 1),
        np.corrcoef(rec["phase1"], c)[0, 1]), flush=True)
