"""Placement variant 1 on the device (k_place_samples_warp: one new sample per warp, place_scan.cuh) against the CPU oracle and
the one-sample-per-thread kernel: synthetic MAT-free trees (the scan path proper), the MAT-free trees the reference built from
its own Newick output, and the reference's frozen MAT trees (every sample takes the straight-line walk inside the kernel and
must reproduce the recorded placements).  The same source is held to the same data on the host, lanes emulated in turn, by
tests/test_place_scan_host.py.  Needs a GPU."""
import math

import numpy as np
import pytest

from golden_io import golden_names, load_extras, load_golden
from maple_b200.genome_list import pack_lists
from maple_b200.model import MapleModel
from test_gpu_placement import _capi_params
from test_oracle_placement_golden import check_placements, place_params
from test_place_scan_host import _mutated, _same
from tree_fixture import tree_arrays, tree_lists

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rv,err,strict,nseq", [(False, False, True, 300), (True, True, False, 200), (False, False, False, 1200)])
def test_warp_placement_matches_oracle_on_synthetic_trees(rv, err, strict, nseq):
    from maple_b200.engine import MapleEngine
    from maple_b200.synthetic import generate
    from maple_b200.tree import DeviceTree
    from oracle.oracle import Oracle
    d = generate(nseq, lRef=4000, mean_diffs=8.0, rate_variation=rv, error_model=err, site_specific_errors=err, seed=11)
    model = d.model
    eng = MapleEngine(model, 0)
    tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
    tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, model.lRef, model.usingErrorRate))
    tree.prepare_search()
    L = math.log(model.lRef)
    pp = {"strictStopRules": int(strict), "allowedFails": 2 if strict else 4, "deeperSearchForLongBranches": 0, "onlyFindIdentical": int(err),
          "thresholdLogLK": (2.0 if strict else 14.0) * L, "thresholdLogLKoptimization": L, "thresholdLogLKconsecutivePlacement": 0.01,
          "effectivelyNon0BLen": 1.0 / (10 * model.lRef), "BLenThresholdDeeperSearch": (L + 5) / model.lRef, "oneMutBLen": 1.0 / model.lRef}
    samples = pack_lists(_mutated(d.tip_lists, model.refIdx, 90), model.lRef, model.usingErrorRate)
    ta = {"up": tree.up, "child0": tree.child0, "child1": tree.child1, "dist": tree.dist, "isTip": tree.isTip, "root": tree.root}
    host = tree.arena.to_host()
    ref = Oracle(model).place_batch(ta, host, pp, samples)
    eng.set_place_variant(1)
    got = tree.place_samples(samples, _capi_params(pp))
    _same(got, ref)
    eng.set_place_variant(0)
    _same(tree.place_samples(samples, _capi_params(pp)), ref)
    assert (ref["status"] == 0).sum() > 20 and ref["phase1"].max() > 96


@pytest.mark.parametrize("name", ["ex_unrest", "ex_unrest_rv", "ay_unrest_300"])
def test_warp_placement_on_reference_built_trees_without_mat(name):
    from maple_b200.engine import MapleEngine
    from maple_b200.tree import DeviceTree
    from oracle.oracle import Oracle
    ex, g = load_extras(name), load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    t = dict(ex["read"]["binary"]["loaded"])
    t["numMinor"] = [len(m) for m in t["minorSequences"]]
    t["children"] = [c or [] for c in t["children"]]
    shim = {"tree": t, "lists": ex["lists"], "env": g["env"]}
    ta, lists = tree_arrays(shim), tree_lists(shim)
    eng = MapleEngine(model, 0)
    tree = DeviceTree.from_lists(eng, ta["up"], ta["child0"], ta["child1"], ta["dist"], ta["root"], ta["isTip"], lists, numMinor=ta["numMinor"])
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    eng.set_place_variant(1)
    got = tree.place_samples(samples, _capi_params(place_params(g)))
    _same(got, Oracle(model).place_batch(ta, lists, place_params(g), samples))


@pytest.mark.parametrize("name", ["ex_unrest", "ay_unrest_deep_200"])
def test_warp_variant_falls_back_on_mat_trees(name):
    from maple_b200.engine import MapleEngine
    from maple_b200.tree import DeviceTree
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    eng = MapleEngine(model, 0)
    ta, lists = tree_arrays(g), tree_lists(g)
    tree = DeviceTree.from_lists(eng, ta["up"], ta["child0"], ta["child1"], ta["dist"], ta["root"], ta["isTip"], lists,
                                 ta["mutStart"], ta["mut"], ta["numMinor"])
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    eng.set_place_variant(1)
    check_placements(g, tree.place_samples(samples, _capi_params(place_params(g))))
