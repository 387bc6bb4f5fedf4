"""Placement variant 2 on the device (k_place_samples_warp_mat: one sample per warp with MAT trees covered) against the
reference's recorded placements on its frozen MAT trees, and against the oracle on MAT-free trees.  The source is identical to
the reference on the host with its lanes emulated (tests/test_place_scan_host.py) but the kernel was written after the GPU
budget of round 1 was spent: set MAPLE_RUN_HW_UNVERIFIED=1 to run it on a B200 (first thing to do in round 2)."""
import os

import numpy as np
import pytest

from golden_io import golden_names, load_golden
from maple_b200.genome_list import pack_lists
from maple_b200.model import MapleModel
from test_gpu_placement import _capi_params
from test_oracle_placement_golden import check_placements, place_params
from test_place_scan_host import _same
from tree_fixture import tree_arrays, tree_lists

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("MAPLE_RUN_HW_UNVERIFIED"), reason="kernel not yet run on hardware (round 2, first call)")]


@pytest.mark.parametrize("name", golden_names())
def test_warp_mat_placement_matches_reference(name):
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
    eng.set_place_variant(2)
    rec = tree.place_samples(samples, _capi_params(place_params(g)), scratch_keys=1 << 14)
    check_placements(g, rec)
    _same(rec, Oracle(model).place_batch(ta, lists, place_params(g), samples))
