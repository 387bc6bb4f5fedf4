"""The oracle's SPR search (startTopologyUpdatesParallel body + findBestParentTopology) against every search the
reference ran on its frozen trees: identical best node, branch lengths and number of phase-1 candidates, scores
within 1e-9, identical proposed moves."""
import numpy as np
import pytest

from golden_io import golden_names, load_golden
from maple_b200.model import MapleModel
from oracle.oracle import Oracle
from tree_fixture import search_params, searched_nodes, tree_arrays, tree_lists


@pytest.mark.parametrize("name", golden_names())
def test_search_matches_reference(name):
    g = load_golden(name)
    orc = Oracle(MapleModel.from_reference_snapshot(g["env"], g["model"]))
    ta, lists = tree_arrays(g), tree_lists(g)
    nodes = searched_nodes(g)
    res = orc.search_batch(ta, lists, search_params(g), nodes, lazy_mode=0)
    by_node = {n: r for n, r in zip(nodes, res)}
    t = g["tree"]
    assert g["searches"]
    nphase1 = 0
    for s in g["searches"]:
        pruned = t["children"][s["node"]][s["child"]]
        r = by_node[pruned]
        assert r["status"] == 0, (s, r)
        assert abs(r["bestCurrentLK"] - s["bestLKdiff"]) <= 1e-9 or r["bestCurrentLK"] == s["bestLKdiff"]
        assert r["phase1"] == s["phase1"], (s, r)
        assert r["bestNode"] == s["bestNode"], (s, r)
        assert abs(r["bestScore"] - s["bestScore"]) <= 1e-9 or r["bestScore"] == s["bestScore"], (s, r)
        assert [r["bLenTop"], r["bLenBottom"], r["bLenAppend"]] == [float(x) for x in s["blens"]], (s, r)
        nphase1 += r["phase1"]
    assert nphase1 == g["phase1Total"]
    # nodes the reference did not search (current cost above the threshold and zero branch length)
    searched = {t["children"][s["node"]][s["child"]] for s in g["searches"]}
    for n, r in by_node.items():
        if n not in searched:
            assert r["status"] == 1 and r["placement"] == -1
    got = sorted((int(n), int(r["placement"])) for n, r in by_node.items() if r["placement"] >= 0)
    exp = sorted((m[0], m[1]) for core in g["proposed"] for m in core)
    assert got == exp
    imp = {m[0]: m[2] for core in g["proposed"] for m in core}
    for n, r in by_node.items():
        if r["placement"] >= 0:
            assert abs(r["improvement"] - imp[n]) <= 1e-9
