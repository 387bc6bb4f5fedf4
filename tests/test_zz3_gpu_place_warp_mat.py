"""Placement variants 2 and 3 on the device (k_place_samples_warp_mat: one sample per warp with MAT trees covered; 3 = with the
parallel window replay) against the
reference's recorded placements on its frozen MAT trees, and against the oracle on MAT-free trees.  The source is identical to
the reference on the host with its lanes emulated (tests/test_place_scan_host.py); variant 3 ran on a B200 in the round-1
bench (records identical to variant 0)."""

import numpy as np
import pytest

from golden_io import golden_names, load_golden
from maple_b200.genome_list import pack_lists
from maple_b200.model import MapleModel
from test_gpu_placement import _capi_params
from test_oracle_placement_golden import check_placements, place_params
from test_place_scan_host import _same
from tree_fixture import tree_arrays, tree_lists

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("name", golden_names())
def test_warp_mat_placement_matches_reference(name, variant):
    from maple_b200.engine import MapleEngine
    from maple_b200.tree import DeviceTree
    from oracle.oracle import Oracle
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    eng = MapleEngine(model, 0)
    ta, lists = tree_arrays(g), tree_lists(g)
    tree = DeviceTree.from_lists(eng, ta["up"], ta["child0"], ta["child1"], ta["dist"], ta["root"], ta["isTip"], lists,
                                 ta["mutStart"], ta["mut"], ta["numMinor"])
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    eng.set_place_variant(variant)
    rec = tree.place_samples(samples, _capi_params(place_params(g)), scratch_keys=1 << 14)
    check_placements(g, rec)
    _same(rec, Oracle(model).place_batch(ta, lists, place_params(g), samples))


@pytest.mark.parametrize("variant", [2, 3])
def test_warp_mat_placement_matches_oracle_on_a_synthetic_tree(variant):
    import math
    from maple_b200.engine import MapleEngine
    from maple_b200.synthetic import generate
    from maple_b200.tree import DeviceTree
    from oracle.oracle import Oracle
    from test_place_scan_host import _mutated
    d = generate(1200, lRef=4000, mean_diffs=8.0, rate_variation=True, seed=11)
    model = d.model
    eng = MapleEngine(model, 0)
    tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
    tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, model.lRef, model.usingErrorRate))
    tree.prepare_search()
    L = math.log(model.lRef)
    pp = {"strictStopRules": 0, "allowedFails": 4, "deeperSearchForLongBranches": 0, "onlyFindIdentical": 0, "thresholdLogLK": 14.0 * L,
          "thresholdLogLKoptimization": L, "thresholdLogLKconsecutivePlacement": 0.01, "effectivelyNon0BLen": 1.0 / (10 * model.lRef),
          "BLenThresholdDeeperSearch": (L + 5) / model.lRef, "oneMutBLen": 1.0 / model.lRef}
    samples = pack_lists(_mutated(d.tip_lists, model.refIdx, 90), model.lRef, model.usingErrorRate)
    ta = {"up": tree.up, "child0": tree.child0, "child1": tree.child1, "dist": tree.dist, "isTip": tree.isTip, "root": tree.root}
    ref = Oracle(model).place_batch(ta, tree.arena.to_host(), pp, samples)
    eng.set_place_variant(variant)
    _same(tree.place_samples(samples, _capi_params(pp)), ref)
