"""Search rounds of the reference beyond the one the main fixtures hold (tests/golden/extras 'rounds', make_golden.py:
harvest_rounds): the DEEP stop rules of its later rounds -- the rules bench.py measures -- on the frozen tree, and both rule sets
on a copy with perturbed branch lengths, recalculated lists and part of the nodes dirty (28-40 accepted proposals per round).
The oracle must reproduce every search (best node, branch lengths, candidate count identical, scores within 1e-9, the same
proposedMoves); the CUDA source on the host (straight-line search and the per-lane state machine of the default kernel) must equal
the oracle search by search, and the reference's record of every search whose outcome does not depend on the order in which the
reference fills probVectTotUp of zero-length children of the root (tree_fixture.compare_with_reference_searches)."""
import numpy as np
import pytest

from golden_io import golden_names, load_extras, load_golden
from hostsim import KernelSourceOnHost
from maple_b200.model import MapleModel
from oracle.oracle import Oracle
from test_kernel_source_host import _prefilled_lists
from tree_fixture import compare_with_reference_searches, search_params, searched_nodes, tree_arrays, tree_lists

ROUNDS = ["frozen_deep", "perturbed_deep", "perturbed_fast"]


def round_shim(name, rnd):
    g, ex = load_golden(name), load_extras(name)
    r = ex["rounds"][rnd]
    if rnd.startswith("frozen"):
        shim = dict(g)
    else:
        t = dict(ex["perturbed"])
        t["numMinor"] = [len(m) for m in t["minorSequences"]]
        shim = {"tree": t, "lists": ex["lists"], "env": g["env"]}
    shim["params"] = r["params"]
    shim["searches"], shim["proposed"], shim["phase1Total"] = r["searches"], r["proposed"], r["phase1Total"]
    return g, shim


@pytest.mark.parametrize("rnd", ROUNDS)
@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_the_reference_round(name, rnd):
    g, s = round_shim(name, rnd)
    orc = Oracle(MapleModel.from_reference_snapshot(g["env"], g["model"]))
    ta, lists, nodes = tree_arrays(s), tree_lists(s), searched_nodes(s)
    res = orc.search_batch(ta, lists, search_params(s), nodes, lazy_mode=0)
    by_node = {n: r for n, r in zip(nodes, res)}
    t = s["tree"]
    assert len(s["searches"]) > 50
    for q in s["searches"]:
        r = by_node[t["children"][q["node"]][q["child"]]]
        assert r["status"] == 0 and r["phase1"] == q["phase1"] and r["bestNode"] == q["bestNode"], (q, r)
        assert abs(r["bestScore"] - q["bestScore"]) <= 1e-9 or r["bestScore"] == q["bestScore"], (q, r)
        assert [r["bLenTop"], r["bLenBottom"], r["bLenAppend"]] == [float(x) for x in q["blens"]], (q, r)
    assert int(res["phase1"].sum()) == s["phase1Total"]
    got = sorted((int(n), int(r["placement"])) for n, r in by_node.items() if r["placement"] >= 0)
    assert got == sorted((m[0], m[1]) for core in s["proposed"] for m in core)
    imp = {m[0]: m[2] for core in s["proposed"] for m in core}
    assert all(abs(r["improvement"] - imp[n]) <= 1e-9 for n, r in by_node.items() if r["placement"] >= 0)
    if rnd != "frozen_deep":
        assert len(got) > 10  # the perturbed tree does get rearranged


@pytest.mark.parametrize("kind", ["straight", "fsm"])
@pytest.mark.parametrize("rnd", ROUNDS)
@pytest.mark.parametrize("name", ["ex_unrest", "ex_unrest_rv_sse", "ex_unrest_err", "ay_unrest_300", "ay_unrest_deep_200", "ay_unrest_1000"])
def test_cuda_source_reproduces_the_round(name, rnd, kind):
    g, s = round_shim(name, rnd)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    orc, hs = Oracle(model), KernelSourceOnHost(model)
    ta, nodes = tree_arrays(s), np.array(searched_nodes(s), np.int32)
    lists = _prefilled_lists(s, orc)
    rec = (hs.search_batch_fsm if kind == "fsm" else hs.search_batch)(ta, lists, search_params(s), nodes, scratch_keys=1 << 17)
    ref = orc.search_batch(ta, lists, search_params(s), nodes, lazy_mode=1)
    for f in ("status", "placement", "bestNode", "phase1", "bLenTop", "bLenBottom", "bLenAppend"):
        assert np.array_equal(rec[f], ref[f]), f
    for f in ("bestCurrentLK", "bestScore", "improvement"):
        assert np.max(np.abs(rec[f] - ref[f])) <= 1e-9, f
    lazy = orc.search_batch(ta, tree_lists(s), search_params(s), nodes, lazy_mode=0)
    compare_with_reference_searches(s, nodes, rec, lazy, ref)
