"""Device placement of new samples (maple_place_batch = findBestParentForNewSample, :7912) against the placements the unmodified
reference computed on its frozen trees and against the CPU oracle -- needs a GPU.  Same bar as the oracle's own pin: node,
minor-sequence verdict, number of candidate branches and branch lengths identical, scores within 1e-9."""
import numpy as np
import pytest

from golden_io import hw_names as golden_names, load_golden
from maple_b200.genome_list import pack_lists
from maple_b200.model import MapleModel
from test_oracle_placement_golden import check_placements, place_params
from tree_fixture import tree_arrays, tree_lists

pytestmark = pytest.mark.gpu


def _capi_params(d):
    from maple_b200 import capi
    p = capi.PlaceParams()
    for k, v in d.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("name", [n for n in golden_names() if "placements" in load_golden(n)])
def test_device_placement_matches_reference_and_oracle(name):
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
    rec = tree.place_samples(samples, _capi_params(place_params(g)))
    check_placements(g, rec)
    ref = Oracle(model).place_batch(ta, lists, place_params(g), samples)
    for f in ("bestNode", "status", "phase1", "missedMinors", "bLenTop", "bLenBottom", "bLenAppend"):
        assert np.array_equal(rec[f], ref[f]), f
    with np.errstate(invalid="ignore"):  # -inf scores compare equal, their difference is nan
        same = (rec["bestScore"] == ref["bestScore"]) | (np.abs(rec["bestScore"] - ref["bestScore"]) <= 1e-9)
    assert same.all()
