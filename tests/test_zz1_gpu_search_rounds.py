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


@pytest.mark.parametrize("variant", [0, 2])
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
