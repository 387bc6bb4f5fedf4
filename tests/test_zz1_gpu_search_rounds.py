"""The device search under the stop rules bench.py measures (the reference's DEEP rounds) and on rearranged trees, against rounds
recorded from the reference itself (tests/golden/extras 'rounds'): frozen tree with the deep rules, and a copy with perturbed
branch lengths under both rule sets (28-40 accepted proposals per round).  Bar: equal to the oracle in everything; the
reference's record of every search whose outcome does not depend on its lazy fill order (at least 98 % of them) and the same
proposedMoves (tree_fixture.compare_with_reference_searches).
The CPU twin (oracle, and the CUDA source compiled for the host) is tests/test_search_rounds_golden.py.  Written after the GPU
budget of round 1 was spent: first run on hardware is the round-end test run.  Needs a GPU."""
import numpy as np
import pytest

from test_gpu_search import _capi_params, _compare
from test_search_rounds_golden import ROUNDS, round_shim
from tree_fixture import compare_with_reference_searches, search_params, searched_nodes, tree_arrays, tree_lists

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [0, 2, 4])
@pytest.mark.parametrize("rnd", ROUNDS)
@pytest.mark.parametrize("name", ["ex_unrest", "ex_unrest_rv_sse", "ay_unrest_300", "ay_unrest_deep_200"])
def test_device_reproduces_the_reference_round(name, rnd, variant):
    from maple_b200.engine import MapleEngine
    from maple_b200.model import MapleModel
    from maple_b200.tree import DeviceTree
    from oracle.oracle import Oracle
    g, s = round_shim(name, rnd)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    eng = MapleEngine(model, 0)
    eng.set_search_variant(variant)
    ta, lists = tree_arrays(s), tree_lists(s)
    tree = DeviceTree.from_lists(eng, ta["up"], ta["child0"], ta["child1"], ta["dist"], ta["root"], ta["isTip"], lists,
                                 ta["mutStart"], ta["mut"], ta["numMinor"])
    nodes = np.array(searched_nodes(s), np.int32)
    tree.prepare_search()
    rec = tree.search_records(tree.spr_search(nodes, _capi_params(search_params(s))))
    orc = Oracle(model)
    pre = orc.search_batch(ta, lists, search_params(s), nodes, lazy_mode=1)
    _compare(rec, pre, nodes)
    compare_with_reference_searches(s, nodes, rec, orc.search_batch(ta, lists, search_params(s), nodes, lazy_mode=0), pre)


@pytest.mark.parametrize("key", ["frozen_nonstrict", "frozen_tight", "perturbed_default", "perturbed_nonstrict", "perturbed_tight"])
@pytest.mark.parametrize("name", ["ex_unrest", "ex_unrest_err", "ay_unrest_300"])
def test_device_placements_under_other_rules(name, key):
    """maple_place_batch (default kernel) against placements recorded from the reference under non-strict and tight stop rules, on
    frozen and perturbed trees (CPU twin: tests/test_placement_rules_golden.py)."""
    from maple_b200.engine import MapleEngine
    from maple_b200.genome_list import pack_lists
    from maple_b200.model import MapleModel
    from maple_b200.tree import DeviceTree
    from test_gpu_placement import _capi_params as place_capi
    from test_oracle_placement_golden import check_placements, place_params
    from test_placement_rules_golden import _shim
    g, tree_shim, s = _shim(name, key)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    eng = MapleEngine(model, 0)
    ta, lists = tree_arrays(tree_shim), tree_lists(tree_shim)
    tree = DeviceTree.from_lists(eng, ta["up"], ta["child0"], ta["child1"], ta["dist"], ta["root"], ta["isTip"], lists,
                                 ta["mutStart"], ta["mut"], ta["numMinor"])
    samples = pack_lists([[tuple(e) for e in s["lists"][c["diffs"]]] for c in s["placements"]], model.lRef, model.usingErrorRate)
    check_placements(s, tree.place_samples(samples, place_capi(place_params(s))))
