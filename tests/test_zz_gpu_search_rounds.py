"""The device search under the stop rules bench.py measures (the reference's DEEP rounds) and on rearranged trees, against rounds
recorded from the reference itself (tests/golden/extras 'rounds'): frozen tree with the deep rules, and a copy with perturbed
branch lengths under both rule sets (28-40 accepted proposals per round).  Bar: every search's best node, branch lengths and score
as the reference recorded them, the same proposedMoves; the candidate count may be one higher in searches that reach a zero-length
child of the root (filled before the round here, lazily by the reference: DESIGN section 5); equal to the oracle in everything.
The CPU twin (oracle, and the CUDA source compiled for the host) is tests/test_search_rounds_golden.py.  Written after the GPU
budget of round 1 was spent: first run on hardware is the round-end test run.  Needs a GPU."""
import numpy as np
import pytest

from test_gpu_search import _capi_params, _compare
from test_search_rounds_golden import ROUNDS, round_shim
from tree_fixture import search_params, searched_nodes, tree_arrays, tree_lists

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
    _compare(rec, Oracle(model).search_batch(ta, lists, search_params(s), nodes, lazy_mode=1), nodes)
    by_node = {int(n): r for n, r in zip(nodes, rec)}
    t = s["tree"]
    for q in s["searches"]:
        r = by_node[t["children"][q["node"]][q["child"]]]
        assert r["status"] == 0 and r["bestNode"] == q["bestNode"], (q, r)
        assert [r["bLenTop"], r["bLenBottom"], r["bLenAppend"]] == [float(x) for x in q["blens"]], (q, r)
        assert r["bestScore"] == q["bestScore"] or abs(r["bestScore"] - q["bestScore"]) <= 1e-9, (q, r)
        assert r["phase1"] - q["phase1"] in (0, 1), (q, r)
    got = sorted((n, int(r["placement"])) for n, r in by_node.items() if r["placement"] >= 0)
    assert got == sorted((m[0], m[1]) for core in s["proposed"] for m in core)
