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


@pytest.mark.parametrize("variant", [0, 3])
def test_place_search_place_on_one_context(variant):
    """Placement, then an SPR round, then the same placement again on ONE context: the three launches share the context's scratch
    allocations (a search once freed the placement scratch under the placement kernel's feet), and a placement batch must leave
    the arena as it found it (its sample lists are temporaries)."""
    from maple_b200 import capi
    from maple_b200.engine import MapleEngine
    from maple_b200.tree import DeviceTree
    from tree_fixture import search_params as fixture_params, searched_nodes
    g = load_golden("ex_unrest")
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    eng = MapleEngine(model, 0)
    eng.set_place_variant(variant)
    ta, lists = tree_arrays(g), tree_lists(g)
    tree = DeviceTree.from_lists(eng, ta["up"], ta["child0"], ta["child1"], ta["dist"], ta["root"], ta["isTip"], lists,
                                 ta["mutStart"], ta["mut"], ta["numMinor"])
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    pp = _capi_params(place_params(g))
    state0 = (tree.arena.n, tree.arena.key_tail, tree.arena.pay_tail)
    first = tree.place_samples(samples, pp)
    assert (tree.arena.n, tree.arena.key_tail, tree.arena.pay_tail) == state0
    sp = capi.SearchParams()
    for k, v in fixture_params(g).items():
        setattr(sp, k, v)
    nodes = np.array(searched_nodes(g), np.int32)
    s1 = tree.search_records(tree.spr_search(nodes, sp))
    second = tree.place_samples(samples, pp)
    s2 = tree.search_records(tree.spr_search(nodes, sp))
    assert first.tobytes() == second.tobytes()
    assert s1.tobytes() == s2.tobytes()
    check_placements(g, second)
