"""Placement variant 1 (one new sample per warp: windowed scans + one-lane replay + per-lane refinement, place_scan.cuh) run
on the host with the lanes of every phase emulated in turn (tests/hostsim), against the oracle's findBestParentForNewSample and
the straight-line kernel source:
  * on synthetic MAT-free trees (the scan path proper: rate variation, error model, strict and permissive stop rules);
  * on the MAT-free trees the reference built from its own Newick output (tests/golden/extras 'loaded': real data with minor
    sequences), with the new samples recorded in the fixtures;
  * on the reference's frozen MAT trees, where variant 1 must hand every sample to the straight-line walk and reproduce the
    recorded placements exactly.
Bar: identical node, status, candidate counts, missed-minor counts and branch lengths; scores within 1e-9."""
import math

import numpy as np
import pytest

from golden_io import golden_names, load_extras, load_golden
from hostsim import KernelSourceOnHost
from maple_b200.genome_list import pack_lists
from maple_b200.model import MapleModel
from oracle.host_tree import build_tree_lists
from oracle.oracle import Oracle
from test_oracle_placement_golden import check_placements, place_params
from tree_fixture import tree_arrays, tree_lists


def _same(a, b):
    for f in ("status", "bestNode", "phase1", "missedMinors", "bLenTop", "bLenBottom", "bLenAppend"):
        assert np.array_equal(a[f], b[f]), (f, [(i, x, y) for i, (x, y) in enumerate(zip(a[f], b[f])) if x != y][:5])
    x, y = a["bestScore"], b["bestScore"]
    fin = np.isfinite(y)
    assert np.array_equal(x[~fin], y[~fin]) and np.max(np.abs(x[fin] - y[fin]), initial=0.0) <= 1e-9


def _mutated(tip_lists, refIdx, count, every=3):
    """New samples: tips with one more substitution inside their first long enough R run (every third one unchanged)."""
    out = []
    for i, v in enumerate(tip_lists[:count]):
        v = [tuple(e) for e in v]
        if i % every:
            prev = 0
            for j, e in enumerate(v):
                end = e[1] if e[0] in (4, 5) else prev + 1
                if e[0] == 4 and len(e) == 2 and end - prev >= 3 + i % 50:
                    pos = prev + 2 + i % 50
                    ref = int(refIdx[pos - 1])
                    v[j:j + 1] = [(4, pos - 1), ((ref + 1 + i % 3) % 4, ref), (4, end)]
                    break
                prev = end
        out.append(v)
    return out


@pytest.mark.parametrize("rv,err,strict,nseq", [(False, False, True, 300), (True, False, False, 300), (True, True, False, 200),
                                                (False, False, False, 1200)])
def test_scan_placement_matches_oracle_on_synthetic_trees(rv, err, strict, nseq):
    from maple_b200.synthetic import generate
    d = generate(nseq, lRef=4000, mean_diffs=8.0, rate_variation=rv, error_model=err, site_specific_errors=err, seed=11)
    model = d.model
    orc, hs = Oracle(model), KernelSourceOnHost(model)
    lists, dist, isTip = build_tree_lists(orc, d.up, d.child0, d.child1, d.dist, d.root, d.tip_nodes, d.tip_lists, model.lRef,
                                          int(model.usingErrorRate))
    ta = {"up": d.up, "child0": d.child0, "child1": d.child1, "dist": dist, "isTip": isTip, "root": d.root}
    L = math.log(model.lRef)
    pp = {"strictStopRules": int(strict), "allowedFails": 2 if strict else 4, "deeperSearchForLongBranches": 0, "onlyFindIdentical": int(err),
          "thresholdLogLK": (2.0 if strict else 14.0) * L, "thresholdLogLKoptimization": L, "thresholdLogLKconsecutivePlacement": 0.01,
          "effectivelyNon0BLen": 1.0 / (10 * model.lRef), "BLenThresholdDeeperSearch": (L + 5) / model.lRef, "oneMutBLen": 1.0 / model.lRef}
    samples = pack_lists(_mutated(d.tip_lists, model.refIdx, 90), model.lRef, model.usingErrorRate)
    ref = orc.place_batch(ta, lists, pp, samples)
    got = hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16)
    _same(got, ref)
    _same(hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16, mat=1), ref)
    _same(hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16, mat=2), ref)
    _same(hs.place_batch(ta, lists, pp, samples, scratch_keys=1 << 16), ref)
    assert (ref["status"] == 0).sum() > 20 and (ref["status"] == 1).sum() > 5
    assert ref["phase1"].max() > 96  # more than one window


@pytest.mark.parametrize("name", ["ex_unrest", "ex_gtr", "ex_unrest_rv", "ex_unrest_err", "ay_unrest_300", "ay_unrest_deep_200"])
def test_scan_placement_on_reference_built_trees_without_mat(name):
    ex, g = load_extras(name), load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    t = dict(ex["read"]["binary"]["loaded"])
    t["numMinor"] = [len(m) for m in t["minorSequences"]]
    t["children"] = [c or [] for c in t["children"]]
    shim = {"tree": t, "lists": ex["lists"], "env": g["env"]}
    ta, lists = tree_arrays(shim), tree_lists(shim)
    assert int(ta["mutStart"][-1]) == 0
    pp = place_params(g)
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    ref = Oracle(model).place_batch(ta, lists, pp, samples)
    hs = KernelSourceOnHost(model)
    _same(hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16), ref)
    _same(hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16, mat=1), ref)
    _same(hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16, mat=2), ref)
    if not pp["deeperSearchForLongBranches"]:
        assert (ref["status"] == 0).sum() > 10


@pytest.mark.parametrize("name", golden_names())
def test_scan_variant_falls_back_on_mat_trees_and_matches_reference(name):
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    rec = KernelSourceOnHost(model).place_batch_scan(tree_arrays(g), tree_lists(g), place_params(g), samples, scratch_keys=1 << 16)
    check_placements(g, rec)


@pytest.mark.parametrize("mat", [1, 2])
@pytest.mark.parametrize("name", golden_names())
def test_mat_covering_variant_matches_reference_on_mat_trees(name, mat):
    """place_sample_warp_mat on the reference's frozen trees WITH their local references: lane 0 walks the part of the tree
    above mutation-carrying nodes, mutation-free subtrees are scan jobs; identical to the 448 recorded placements."""
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    ta = tree_arrays(g)
    rec = KernelSourceOnHost(model).place_batch_scan(ta, tree_lists(g), place_params(g), samples, scratch_keys=1 << 16, mat=mat)
    check_placements(g, rec)
    ref = Oracle(model).place_batch(ta, tree_lists(g), place_params(g), samples)
    _same(rec, ref)
    assert int(ta["mutStart"][-1]) > 0  # these trees do carry MAT mutations


@pytest.mark.parametrize("strict,fails,thr,cons", [(1, 0, 0.5, 0.01), (1, 1, 1.0, 0.5), (0, 0, 0.5, 0.01), (1, 0, 3.0, 2.0)])
def test_parallel_replay_under_aggressive_stop_rules(strict, fails, thr, cons):
    """Stop rules that prune hard leave better-scoring nodes in subtrees the walk never enters: the parallel replay's running best
    (taken over every earlier score of the window) is then wrong, it notices, and the window is replayed serially (an instrumented
    build counted 77 such windows next to 20 042 committed ones for these four settings).  Results must not change."""
    from maple_b200.synthetic import generate
    d = generate(1500, lRef=4000, mean_diffs=8.0, rate_variation=False, seed=21)
    model = d.model
    orc, hs = Oracle(model), KernelSourceOnHost(model)
    lists, dist, isTip = build_tree_lists(orc, d.up, d.child0, d.child1, d.dist, d.root, d.tip_nodes, d.tip_lists, model.lRef, 0)
    ta = {"up": d.up, "child0": d.child0, "child1": d.child1, "dist": dist, "isTip": isTip, "root": d.root}
    L = math.log(model.lRef)
    pp = {"strictStopRules": strict, "allowedFails": fails, "deeperSearchForLongBranches": 0, "onlyFindIdentical": 0, "thresholdLogLK": thr * L,
          "thresholdLogLKoptimization": L, "thresholdLogLKconsecutivePlacement": cons, "effectivelyNon0BLen": 1.0 / (10 * model.lRef),
          "BLenThresholdDeeperSearch": (L + 5) / model.lRef, "oneMutBLen": 1.0 / model.lRef}
    samples = pack_lists(_mutated(d.tip_lists, model.refIdx, 300, every=7), model.lRef, 0)
    ref = orc.place_batch(ta, lists, pp, samples)
    for mat in (0, 1, 2):
        _same(hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16, mat=mat), ref)


@pytest.mark.parametrize("deep_child", [0, 1])
def test_warp_placement_on_a_caterpillar_tree(deep_child):
    """A ladder of 160 tips: the tree is 159 levels deep, so the per-depth states go past the 48 kept in the warp block into the
    global array, a window holds an ancestor chain of 96 nodes (the longest the pointer jumping can meet: 7 rounds) when the deep
    child is explored first, and subtrees are skipped from far above."""
    from maple_b200.synthetic import generate
    d = generate(160, lRef=4000, mean_diffs=8.0, rate_variation=True, seed=31)
    model = d.model
    n_tips = len(d.tip_lists)
    # nodes 0..n_tips-1 tips, n_tips.. internals; internal k has the tip k and the next internal (the last one: two tips)
    n = 2 * n_tips - 1
    up, c0, c1 = np.full(n, -1, np.int32), np.full(n, -1, np.int32), np.full(n, -1, np.int32)
    rng = np.random.default_rng(5)
    dist = np.where(rng.random(n) < 0.3, 0.0, rng.exponential(1.5 / model.lRef, n))
    for k in range(n_tips - 1):
        node = n_tips + k
        deep = n_tips + k + 1 if k < n_tips - 2 else n_tips - 1
        pair = (k, deep) if deep_child == 1 else (deep, k)
        c0[node], c1[node] = pair
        up[pair[0]] = up[pair[1]] = node
    root = n_tips
    dist[root] = 0.0
    orc, hs = Oracle(model), KernelSourceOnHost(model)
    lists, dist2, isTip = build_tree_lists(orc, up, c0, c1, dist, root, np.arange(n_tips), d.tip_lists, model.lRef, int(model.usingErrorRate))
    ta = {"up": up, "child0": c0, "child1": c1, "dist": dist2, "isTip": isTip, "root": root}
    L = math.log(model.lRef)
    pp = {"strictStopRules": 0, "allowedFails": 6, "deeperSearchForLongBranches": 0, "onlyFindIdentical": 0, "thresholdLogLK": 18.0 * L,
          "thresholdLogLKoptimization": L, "thresholdLogLKconsecutivePlacement": 0.01, "effectivelyNon0BLen": 1.0 / (10 * model.lRef),
          "BLenThresholdDeeperSearch": (L + 5) / model.lRef, "oneMutBLen": 1.0 / model.lRef}
    samples = pack_lists(_mutated(d.tip_lists, model.refIdx, 60), model.lRef, model.usingErrorRate)
    ref = orc.place_batch(ta, lists, pp, samples)
    assert ref["phase1"].max() > 150 and (ref["status"] == 0).sum() > 20
    for mat in (0, 1, 2):
        _same(hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16, mat=mat), ref)
    pp["strictStopRules"], pp["allowedFails"], pp["thresholdLogLK"] = 1, 1, 2.0 * L
    ref = orc.place_batch(ta, lists, pp, samples)
    for mat in (0, 1, 2):
        _same(hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16, mat=mat), ref)
