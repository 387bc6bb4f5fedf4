"""The body of the default search kernel (k_spr_search_fsm: per-lane state machines + warp-cooperative subtree scans) run on the
host with its 32 lanes emulated as coroutines that meet at the *_sync intrinsics (tests/hostsim/hostwarp.cpp,
shim_warp/cuda_runtime.h).  Both forms of the scans -- warp_scan_job (search_fsm.cuh) and warp_scan_job2 (scan2.cuh: scan-format
lists, prefix-form replay) -- must give the records of the straight-line search of the same source, bit for bit (same libm),
on every fixture and on the rounds recorded under the other stop rules; and they must equal the reference's own record of every
search whose outcome does not depend on the order in which the reference fills probVectTotUp of zero-length children of the root.
Shapes the default sizing does not produce are forced: tiny pools (windows cut by capacity, windows scored from the arena lists),
one and many searches per warp, every subtree scanned (scan_min_size 1)."""
import numpy as np
import pytest

from golden_io import golden_names, load_golden
from hostsim import KernelSourceOnHost, WarpKernelOnHost
from maple_b200.model import MapleModel
from oracle.oracle import Oracle
from test_kernel_source_host import _prefilled_lists
from test_search_rounds_golden import ROUNDS, round_shim
from tree_fixture import compare_with_reference_searches, search_params, searched_nodes, tree_arrays, tree_lists

FIELDS = ("status", "placement", "bestNode", "phase1", "bLenTop", "bLenBottom", "bLenAppend", "bestCurrentLK", "bestScore", "improvement")


def _same(a, b):
    for f in FIELDS:
        x, y = a[f], b[f]
        assert np.array_equal(x, y) or np.all((x == y) | ((x != x) & (y != y))), f


@pytest.mark.parametrize("form", [1, 2])
@pytest.mark.parametrize("name", [n for n in golden_names() if n != "ay_unrest_1000"])
def test_warp_body_equals_straight_line_search(name, form):
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    hs, hw = KernelSourceOnHost(model), WarpKernelOnHost(model)
    ta, nodes = tree_arrays(g), np.array(searched_nodes(g), np.int32)
    if len(nodes) > 1000:  # the emulation runs ~50 searches a second: a spread sample of the big fixtures
        nodes = nodes[::7]
    lists = _prefilled_lists(g, Oracle(model))
    want = hs.search_batch(ta, lists, search_params(g), nodes, scratch_keys=1 << 15)
    st = np.zeros(40, np.uint64)
    got = hw.search_batch_warp(ta, lists, search_params(g), nodes, scan_form=form, stats=st, big_slots=64)
    _same(got, want)
    if not g["env"]["deeperSearchForLongBranches"]:  # (that option keeps every node on the lane path)
        assert st[17] > 0 and st[21] > 0  # scan jobs ran and counted candidates


@pytest.mark.parametrize("rnd", ROUNDS)
@pytest.mark.parametrize("name", ["ex_unrest", "ex_unrest_rv_sse", "ay_unrest_deep_200"])
def test_second_form_on_the_other_rounds(name, rnd):
    g, s = round_shim(name, rnd)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    orc, hs, hw = Oracle(model), KernelSourceOnHost(model), WarpKernelOnHost(model)
    ta, nodes = tree_arrays(s), np.array(searched_nodes(s), np.int32)
    lists = _prefilled_lists(s, orc)
    want = hs.search_batch(ta, lists, search_params(s), nodes, scratch_keys=1 << 17)
    got = hw.search_batch_warp(ta, lists, search_params(s), nodes, scan_form=2, scratch_keys=1 << 15)
    _same(got, want)
    ref = orc.search_batch(ta, lists, search_params(s), nodes, lazy_mode=1)
    lazy = orc.search_batch(ta, tree_lists(s), search_params(s), nodes, lazy_mode=0)
    compare_with_reference_searches(s, nodes, got, lazy, ref)


@pytest.mark.parametrize("shape", [dict(pool_bytes=1024, lanes_per_warp=1), dict(pool_bytes=256, lanes_per_warp=32),
                                   dict(pool_bytes=4096, scan_min_size=1, lanes_per_warp=7), dict(pool_bytes=0, lanes_per_warp=2),
                                   dict(scan_flags=2, lanes_per_warp=4)])
@pytest.mark.parametrize("name", ["ex_unrest_rv", "ay_unrest_300"])
def test_second_form_forced_shapes(name, shape):
    """Pools of 256 - 4096 bytes cut most windows short (a 256-byte pool holds about one list; with none at all every window is
    scored from the arena lists); scan_flags=2 replays every window node by node, the path a window takes when the stop rule
    prunes a node that holds a new best; the records must not change."""
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    hs, hw = KernelSourceOnHost(model), WarpKernelOnHost(model)
    ta, nodes = tree_arrays(g), np.array(searched_nodes(g), np.int32)
    if len(nodes) > 150:
        nodes = nodes[::2]
    lists = _prefilled_lists(g, Oracle(model))
    want = hs.search_batch(ta, lists, search_params(g), nodes, scratch_keys=1 << 15)
    _same(hw.search_batch_warp(ta, lists, search_params(g), nodes, scan_form=2, **shape), want)


@pytest.mark.parametrize("shape", [dict(owner_warps=1, server_warps=1, lanes_per_warp=32), dict(owner_warps=2, server_warps=3, lanes_per_warp=32),
                                   dict(owner_warps=3, server_warps=1, lanes_per_warp=5, scan_min_size=1),
                                   dict(owner_warps=1, server_warps=2, lanes_per_warp=32, pool_bytes=192)])
@pytest.mark.parametrize("name", ["ex_unrest_rv", "ay_unrest_300"])
def test_scan_service(name, shape):
    """The scan service: the warps that own the searches post their subtree scans in global-memory slots, other warps take them
    from a ticket ring, run them and hand the results back (scan2.cuh: ScanQueue, scan_server_loop; fsm_warp_loop).  All warps are
    emulated together, a spinning lane lets the others run.  With a 192-byte pool the removed list's copy fits no server: every
    job is declined and run by the warp that owns it.  Records as without the service."""
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    hs, hw = KernelSourceOnHost(model), WarpKernelOnHost(model)
    ta, nodes = tree_arrays(g), np.array(searched_nodes(g), np.int32)
    if len(nodes) > 150:
        nodes = nodes[::2]
    lists = _prefilled_lists(g, Oracle(model))
    want = hs.search_batch(ta, lists, search_params(g), nodes, scratch_keys=1 << 15)
    st = np.zeros(40, np.uint64)
    got = hw.search_batch_warp(ta, lists, search_params(g), nodes, scan_form=2, big_slots=64, stats=st, **shape)
    _same(got, want)
    assert st[27] > 0  # jobs went through the servers


@pytest.mark.parametrize("rv,err,strict,ml,rows", [(True, False, False, True, 100000), (True, True, False, False, 37), (False, False, True, False, 100000)])
def test_dense_scoring_pass(rv, err, strict, ml, rows):
    """The dense scoring pass (scan2.cuh: DenseScores): every scorable node scored against the removed list of every search that
    will run, before the searches; their subtree scans then only read.  With room for 37 rows the other searches scan as usual.
    Records as the straight-line search.  Synthetic trees (rate variation, site-specific error model, both stop-rule settings): the
    pass applies to trees without MAT mutations, and the reference's own trees carry them."""
    import math
    from maple_b200.synthetic import generate
    from oracle.host_tree import build_tree_lists
    d = generate(400, lRef=4000, mean_diffs=8.0, rate_variation=rv, error_model=err, site_specific_errors=err, seed=5, ml_like_blens=ml)
    model = d.model
    orc, hs, hw = Oracle(model), KernelSourceOnHost(model), WarpKernelOnHost(model)
    lists, dist, isTip = build_tree_lists(orc, d.up, d.child0, d.child1, d.dist, d.root, d.tip_nodes, d.tip_lists, model.lRef,
                                          int(model.usingErrorRate))
    ta = {"up": d.up, "child0": d.child0, "child1": d.child1, "dist": dist, "isTip": isTip, "root": d.root}
    L = math.log(model.lRef)
    sp = {"strictTopologyStopRules": int(strict), "allowedFailsTopology": 2 if strict else 4, "deeperSearchForLongBranches": 0,
          "thresholdLogLKtopology": (2.0 if strict else 14.0) * L, "thresholdTopologyPlacement": -0.1,
          "thresholdLogLKoptimizationTopology": L, "thresholdLogLKconsecutivePlacement": 1.0,
          "effectivelyNon0BLen": 1.0 / (10 * model.lRef), "BLenThresholdDeeperSearch": (L + 5) / model.lRef, "defaultBLen": 0.000033}
    nodes = np.array([i for i in range(len(d.up)) if d.up[i] >= 0], np.int32)[::5]
    want = hs.search_batch(ta, lists, sp, nodes, scratch_keys=1 << 15)
    st = np.zeros(40, np.uint64)
    got = hw.search_batch_warp(ta, lists, sp, nodes, scan_form=2, big_slots=64, stats=st, dense_rows=rows, lanes_per_warp=6)
    _same(got, want)
    assert 0 < st[30] <= rows and st[21] > 1000


def test_head_of_the_list_one_search_per_warp():
    """BigScratch::headEnd (maple_ctx_set_head_searches): the first entries of the list pulled by lane 0 only, the other lanes
    waiting until they are gone -- every head length from 'one entry' to 'the whole list' gives the records of the plain run."""
    g = load_golden("ay_unrest_300")
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    hs, hw = KernelSourceOnHost(model), WarpKernelOnHost(model)
    ta, nodes = tree_arrays(g), np.array(searched_nodes(g), np.int32)[::3]
    lists = _prefilled_lists(g, Oracle(model))
    want = hs.search_batch(ta, lists, search_params(g), nodes, scratch_keys=1 << 15)
    for head in (1, 7, len(nodes) // 2, len(nodes)):
        got = hw.search_batch_warp(ta, lists, search_params(g), nodes, scan_form=2, lanes_per_warp=5, big_slots=8, head_end=head)
        _same(got, want)


@pytest.mark.parametrize("eval_slice", [1024, 96, 0])
def test_queued_phase2_entries_evaluated_by_the_warp(eval_slice):
    """The phase-2 entries a subtree scan queues are evaluated one per lane (search_fsm.cuh: warp_eval_queue) and folded in the
    reference's order; with 96 scratch entries per lane most batches do not fit and the owning lane goes through its queue
    itself, with none it always does.  Deep rules on a perturbed tree (queues of many entries).  Records as the straight-line search."""
    g, s = round_shim("ex_unrest_rv", "perturbed_deep")
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    orc, hs, hw = Oracle(model), KernelSourceOnHost(model), WarpKernelOnHost(model)
    ta, nodes = tree_arrays(s), np.array(searched_nodes(s), np.int32)[::2]
    lists = _prefilled_lists(s, orc)
    want = hs.search_batch(ta, lists, search_params(s), nodes, scratch_keys=1 << 17)
    st = np.zeros(40, np.uint64)
    got = hw.search_batch_warp(ta, lists, search_params(s), nodes, scan_form=2, scratch_keys=1 << 15, stats=st, eval_slice=eval_slice, lanes_per_warp=4)
    _same(got, want)
    if eval_slice == 1024:
        assert st[32] > 20 and st[33] == 0, (st[32], st[33])
    elif eval_slice == 96:
        assert st[33] > 0
    else:
        assert st[32] == 0


def test_search_that_exhausts_its_scratch_starts_over_in_a_large_slot():
    """With 512 entries of scratch per lane many searches of this tree run out; each takes one of the launch's large slots (8x)
    and starts over inside the same launch.  Records as with ample scratch; with too few slots the rest report status 3."""
    g = load_golden("ex_unrest")
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    hs, hw = KernelSourceOnHost(model), WarpKernelOnHost(model)
    ta, nodes = tree_arrays(g), np.array(searched_nodes(g), np.int32)
    lists = _prefilled_lists(g, Oracle(model))
    want = hs.search_batch(ta, lists, search_params(g), nodes, scratch_keys=1 << 15)
    st = np.zeros(40, np.uint64)
    got = hw.search_batch_warp(ta, lists, search_params(g), nodes, scan_form=2, scratch_keys=512, big_slots=256, stats=st, lanes_per_warp=5)
    assert 0 < st[31] <= 256, st[31]
    _same(got, want)
    few = hw.search_batch_warp(ta, lists, search_params(g), nodes, scan_form=2, scratch_keys=512, big_slots=2, lanes_per_warp=5)
    over = few["status"] == 3
    assert over.sum() == st[31] - 2
    _same(few[~over], want[~over])


def test_second_form_big_fixture_deep_round():
    """1 000 sequences, deep stop rules: long scan jobs, deep paths."""
    g, s = round_shim("ay_unrest_1000", "frozen_deep")
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    orc, hs, hw = Oracle(model), KernelSourceOnHost(model), WarpKernelOnHost(model)
    ta, nodes = tree_arrays(s), np.array(searched_nodes(s), np.int32)[::6]
    lists = _prefilled_lists(s, orc)
    want = hs.search_batch(ta, lists, search_params(s), nodes, scratch_keys=1 << 17)
    st = np.zeros(40, np.uint64)
    got = hw.search_batch_warp(ta, lists, search_params(s), nodes, scan_form=2, scratch_keys=1 << 15, stats=st)
    _same(got, want)
    assert st[21] > 10000


@pytest.mark.parametrize("name", ["ex_unrest", "ex_unrest_err", "ex_unrest_rv_sse", "ay_unrest_300", "syn_unrest_rv_2000", "syn_unrest_rv_sse_1500"])
def test_scan_format_append_equals_dev_append(name):
    """appendProbNode through the scan-format copies (precomputed Q*rate, precomputed removed-side factors) against dev_append of
    the same source: recorded calls and random pairs of the fixture's lists with branch lengths from the corners; bit for bit."""
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    hs, hw = KernelSourceOnHost(model), WarpKernelOnHost(model)
    L = g["lists"]
    for c in g["calls"]["appendProbNode"]:
        a = hw.scan_append(L[c["P"]], L[c["C"]], c["isTipC"], c["bLen"])
        b = hs.append(L[c["P"]], L[c["C"]], c["isTipC"], c["bLen"])
        assert a == b, c
        assert a == c["out"] or abs(a - c["out"]) <= 1e-9
        assert hw.scan_append(L[c["P"]], L[c["C"]], c["isTipC"], c["bLen"], convert_slow=True) == b, c
    rng = np.random.default_rng(11)
    lower = [i for i in set(g["tree"]["probVect"]) if i is not None]
    cand = [i for i in set(g["tree"]["probVectTotUp"]) if i is not None]
    blens = [0.0, 1e-9, 0.3 / model.lRef, 1.0 / model.lRef, 3.3 / model.lRef, 0.1]
    for _ in range(1500):
        P, Cc = L[cand[rng.integers(len(cand))]], L[lower[rng.integers(len(lower))]]
        tip, bl = bool(rng.integers(2)), blens[rng.integers(len(blens))]
        a, b = hw.scan_append(P, Cc, tip, bl), hs.append(P, Cc, tip, bl)
        assert a == b, (P, Cc, tip, bl, a, b)
        a = hw.scan_append(P, Cc, tip, bl, convert_slow=True)  # the job's own factors for the O entries below the shortcut
        assert a == b, (P, Cc, tip, bl, a, b)
