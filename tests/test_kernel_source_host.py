"""The CUDA source itself against the reference, without a GPU: the lane-level device code (likelihood.cuh, search.cuh,
place.cuh, and the per-lane state machine of search_fsm.cuh -- what k_append / k_merge / k_blen / ... / k_spr_search (variant 1) /
k_spr_search_fsm without its warp scans (variant 2) / k_place_samples execute per thread) is compiled for the host by
tests/hostsim and run through the same golden-vector checks as the oracle, plus oracle comparisons of whole searches and
placements.  The shuffle-based subtree scan of the default search kernel (warp_scan_job) needs the hardware and is covered by
the -m gpu tests; its two per-lane append forms (site-converged, queued-site) are checked here too."""
import math

import numpy as np
import pytest

import test_oracle_golden as og
from golden_io import golden_names, load_golden
from hostsim import KernelSourceOnHost
from maple_b200.genome_list import pack_lists
from maple_b200.model import MapleModel
from oracle.oracle import Oracle
from test_oracle_placement_golden import check_placements, place_params
from tree_fixture import FAMILIES, compare_with_reference_searches, search_params, searched_nodes, tree_arrays, tree_lists

NAMES = golden_names()


@pytest.fixture(scope="module", params=NAMES)
def fx(request):
    g = load_golden(request.param)
    return g, KernelSourceOnHost(MapleModel.from_reference_snapshot(g["env"], g["model"]), with_root_tables=True)


@pytest.mark.parametrize("check", ["test_append", "test_merge", "test_blen", "test_differ", "test_pass_branch", "test_shorten",
                                   "test_root_vector", "test_prob_root", "test_tree_likelihood_from_parts"])
def test_primitives_against_reference_vectors(fx, check):
    getattr(og, check)(fx)  # the oracle's own golden checks, computing with the kernel source


@pytest.mark.parametrize("which", ["sitewise", "q4"])
def test_scan_append_forms_against_reference_vectors(fx, which):
    g, hs = fx
    L = g["lists"]
    n = 0
    for c in g["calls"]["appendProbNode"]:
        got = hs.append_variant(which, L[c["P"]], L[c["C"]], c["isTipC"], c["bLen"])
        if c["out"] == float("-inf"):
            assert got == float("-inf")
        else:
            assert abs(got - c["out"]) <= 1e-9
        # the three forms run the same arithmetic in the same order: the lane path and the warp scans of one search or
        # placement may mix them without changing a single `>` / `>=` decision
        assert got == hs.append(L[c["P"]], L[c["C"]], c["isTipC"], c["bLen"])
        n += 1
    assert n >= 40


def _prefilled_lists(g, orc):
    """probVectTotUp of zero-length children of the root filled ahead of the round (DeviceTree.prepare_search)."""
    t, L = g["tree"], g["lists"]
    fam = {f: [None if j is None else L[j] for j in t[f]] for f in FAMILIES}
    root = t["root"]
    if t["children"][root]:
        for c, up in ((t["children"][root][0], "probVectUpRight"), (t["children"][root][1], "probVectUpLeft")):
            if t["dist"][c] == 0.0 and fam["probVectTotUp"][c] is None and fam[up][root] is not None:
                fam["probVectTotUp"][c] = orc.merge(fam[up][root], 0.0, False, fam["probVect"][c], 0.0, False, isUpDown=True)
    flat = []
    for f in FAMILIES:
        flat.extend(fam[f])
    return pack_lists(flat, g["env"]["lRef"], g["env"]["usingErrorRate"])


def _search(hs, kind, ta, lists, params, nodes, scratch_keys):
    """kind 'straight': search_node (k_spr_search, variant 1); 'fsm': fsm_step / fsm_finish, the per-lane state machine of the
    default kernel k_spr_search_fsm, driven like the kernel drives one lane, warp scans off (variant 2)."""
    if kind == "fsm":
        return hs.search_batch_fsm(ta, lists, params, nodes, scratch_keys=scratch_keys)
    return hs.search_batch(ta, lists, params, nodes, scratch_keys=scratch_keys)


@pytest.mark.parametrize("kind", ["straight", "fsm"])
@pytest.mark.parametrize("name", NAMES)
def test_straight_line_search_source_matches_oracle_and_reference(name, kind):
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    orc, hs = Oracle(model), KernelSourceOnHost(model)
    ta, nodes = tree_arrays(g), np.array(searched_nodes(g), np.int32)
    lists = _prefilled_lists(g, orc)
    rec = _search(hs, kind, ta, lists, search_params(g), nodes, 1 << 16)
    ref = orc.search_batch(ta, lists, search_params(g), nodes, lazy_mode=1)
    for f in ("status", "placement", "bestNode", "phase1", "bLenTop", "bLenBottom", "bLenAppend"):
        assert np.array_equal(rec[f], ref[f]), f
    for f in ("bestCurrentLK", "bestScore", "improvement"):
        assert np.max(np.abs(rec[f] - ref[f])) <= 1e-9, f
    lazy = orc.search_batch(ta, tree_lists(g), search_params(g), nodes, lazy_mode=0)
    compare_with_reference_searches(g, nodes, rec, lazy, ref)


@pytest.mark.parametrize("name", NAMES)
def test_placement_source_matches_reference(name):
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    hs = KernelSourceOnHost(model)
    rec = hs.place_batch(tree_arrays(g), _prefilled_lists(g, Oracle(model)), place_params(g), samples, scratch_keys=8192)
    check_placements(g, rec)
    assert int(rec["missedMinors"].sum()) >= 0


def test_scratch_exhaustion_is_reported_not_overrun():
    g = load_golden("ex_unrest")
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    hs = KernelSourceOnHost(model)
    ta, nodes = tree_arrays(g), np.array(searched_nodes(g), np.int32)
    rec = hs.search_batch(ta, _prefilled_lists(g, Oracle(model)), search_params(g), nodes, scratch_keys=64)
    assert (rec["status"] == 3).any() and not math.isnan(float(rec["bestCurrentLK"].sum()))


@pytest.mark.parametrize("kind", ["straight", "fsm"])
@pytest.mark.parametrize("rv,err,strict,ml", [(False, False, True, False), (True, False, False, True), (True, True, False, False)])
def test_straight_line_search_source_matches_oracle_on_synthetic_trees(rv, err, strict, ml, kind):
    """Same configurations as the device test (tests/test_gpu_search.py), on the host: rate variation, site-specific error model,
    strict and deep stop rules, ML-like branch lengths."""
    from maple_b200.synthetic import generate
    from oracle.host_tree import build_tree_lists
    d = generate(400, lRef=4000, mean_diffs=8.0, rate_variation=rv, error_model=err, site_specific_errors=err, seed=5, ml_like_blens=ml)
    model = d.model
    orc, hs = Oracle(model), KernelSourceOnHost(model)
    lists, dist, isTip = build_tree_lists(orc, d.up, d.child0, d.child1, d.dist, d.root, d.tip_nodes, d.tip_lists, model.lRef,
                                          int(model.usingErrorRate))
    ta = {"up": d.up, "child0": d.child0, "child1": d.child1, "dist": dist, "isTip": isTip, "root": d.root}
    L = math.log(model.lRef)
    sp = {"strictTopologyStopRules": int(strict), "allowedFailsTopology": 2 if strict else 4, "deeperSearchForLongBranches": 0,
          "thresholdLogLKtopology": (2.0 if strict else 14.0) * L, "thresholdTopologyPlacement": -0.1,
          "thresholdLogLKoptimizationTopology": L, "thresholdLogLKconsecutivePlacement": 1.0,
          "effectivelyNon0BLen": 1.0 / (10 * model.lRef), "BLenThresholdDeeperSearch": (L + 5) / model.lRef, "defaultBLen": 0.000033}
    nodes = np.array([i for i in range(len(d.up)) if d.up[i] >= 0], np.int32)
    rec = _search(hs, kind, ta, lists, sp, nodes, 1 << 15)
    ref = orc.search_batch(ta, lists, sp, nodes, lazy_mode=1)
    for f in ("status", "placement", "bestNode", "phase1", "bLenTop", "bLenBottom", "bLenAppend"):
        assert np.array_equal(rec[f], ref[f]), f
    for f in ("bestCurrentLK", "bestScore", "improvement"):
        fin = np.isfinite(ref[f])
        assert np.array_equal(rec[f][~fin], ref[f][~fin]) and np.max(np.abs(rec[f][fin] - ref[f][fin]), initial=0.0) <= 1e-9, f
    assert (ref["status"] == 0).sum() > 100 and (ref["placement"] >= 0).sum() > 0 and int(ref["phase1"].sum()) > 20000
