"""The python orchestration above the C ABI (maple_b200/tree.py) run on CPU tensors over tests/fake_device.py, i.e. the same
code the GPU tests run with the CUDA library swapped for the oracle: input-tree set-up -> list building -> likelihood, and
the fast branch-length sweep, against the reference's recorded results.  Guards the host logic where no GPU is available;
the kernels themselves are covered by the -m gpu tests."""
import numpy as np
import pytest

from fake_device import FakeEngine
from golden_io import load_extras, load_golden
from maple_b200.genome_list import lists_equal
from maple_b200.model import MapleModel
from maple_b200.newick import load_input_tree
from maple_b200.tree import DeviceTree
from tree_fixture import tree_arrays, tree_lists

FAMILIES = ("probVect", "probVectUpRight", "probVectUpLeft", "probVectTotUp")


@pytest.mark.parametrize("name", ["ex_unrest", "ex_unrest_err", "ay_unrest_300"])
def test_input_tree_through_device_tree_code(name, tmp_path):
    ex, g = load_extras(name), load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    nwk, aln = tmp_path / "t.nwk", tmp_path / "a.txt"
    nwk.write_text(ex["newick"]["binary"] + "\n")
    aln.write_text(ex["alignmentText"])
    t, root, names, tip_nodes, tip_lists = load_input_tree(str(nwk), str(aln), model, default_blen=g["env"]["defaultBLen"],
                                                           only_find_identical=g["placeEnv"]["onlyFindIdentical"])
    tree = DeviceTree.from_host_tree(FakeEngine(model), t, root, tip_nodes, tip_lists)
    want = ex["read"]["binary"]["loaded"]
    assert [float(x) for x in tree.dist] == want["dist"]
    for i in t.reachable(root):
        got = tree.lists_of(i)
        for f, fam in enumerate(FAMILIES):
            if fam == "probVectTotUp" and tree.dist[i] == 0 and tree.up[i] == root:
                continue
            j = want[fam][i]
            assert lists_equal(got[f], None if j is None else ex["lists"][j]), (fam, i)
    assert abs(tree.tree_likelihood() - ex["read"]["binary"]["loadedLK"]) <= 1e-6


@pytest.mark.parametrize("name", ["ex_unrest", "ex_unrest_rv_sse", "ay_unrest_300"])
@pytest.mark.parametrize("which", ["frozen", "perturbed"])
def test_fast_sweep_through_device_tree_code(name, which):
    ex, g = load_extras(name), load_golden(name)
    if which == "frozen":
        shim, want = g, ex["sweeps"]["fastPass"]
    else:
        t = dict(ex["perturbed"])
        t["numMinor"] = [len(m) for m in t["minorSequences"]]
        shim, want = {"tree": t, "lists": ex["lists"], "env": g["env"]}, ex["sweeps"]["perturbed_fastPass"]
    a = tree_arrays(shim)
    eng = FakeEngine(MapleModel.from_reference_snapshot(g["env"], g["model"]))
    tree = DeviceTree.from_lists(eng, a["up"], a["child0"], a["child1"], a["dist"], a["root"], a["isTip"], tree_lists(shim),
                                 mutStart=a["mutStart"], mut=a["mut"], numMinor=a["numMinor"])
    n_ids = tree.arena.n
    updates, dirty = tree.optimize_branch_lengths(g["env"]["effectivelyNon0BLen"], dirty=shim["tree"]["dirty"])
    assert updates == want["updates"]
    assert [float(x) for x in tree.dist] == want["dist"]
    assert np.array_equal(tree.d_dist.numpy(), tree.dist)
    assert [bool(x) for x in dirty] == want["dirty"]
    assert tree.arena.n == n_ids
    # the frozen tree's likelihood is still the reference's (lists untouched by the sweep)
    if which == "frozen":
        assert abs(tree.tree_likelihood() - g["treeLK"]) <= 1e-6


@pytest.mark.parametrize("variant", [0, 1, 3])
def test_place_samples_wrapper_and_its_retries(variant):
    """DeviceTree.place_samples over the stand-in library (the kernel source on the host behind it): samples are appended to the
    arena, the tree is re-bound, and samples whose scratch (here: a bestNodes table of 1 024 entries under very permissive rules) is
    exhausted come back with status 3 and are re-run with 8x the scratch through variant 0, whatever variant the batch started on."""
    import math
    from maple_b200 import capi
    from maple_b200.genome_list import pack_lists
    from maple_b200.synthetic import generate
    from oracle.oracle import Oracle
    from test_place_scan_host import _mutated
    d = generate(1500, lRef=4000, mean_diffs=8.0, rate_variation=False, seed=21)
    model = d.model
    eng = FakeEngine(model)
    tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
    tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, model.lRef, 0))
    L = math.log(model.lRef)
    pp = {"strictStopRules": 0, "allowedFails": 8, "deeperSearchForLongBranches": 0, "onlyFindIdentical": 0, "thresholdLogLK": 40.0 * L,
          "thresholdLogLKoptimization": 6.0 * L, "thresholdLogLKconsecutivePlacement": 0.01, "effectivelyNon0BLen": 1.0 / (10 * model.lRef),
          "BLenThresholdDeeperSearch": (L + 5) / model.lRef, "oneMutBLen": 1.0 / model.lRef}
    p = capi.PlaceParams()
    for k, v in pp.items():
        setattr(p, k, v)
    samples = pack_lists(_mutated(d.tip_lists, model.refIdx, 24), model.lRef, 0)
    ta = {"up": tree.up, "child0": tree.child0, "child1": tree.child1, "dist": tree.dist, "isTip": tree.isTip, "root": tree.root}
    ref = Oracle(model).place_batch(ta, tree.arena.to_host(), pp, samples)
    eng.set_place_variant(variant)
    first = []
    orig = eng.lib.maple_place_batch

    def spy(ctx, params, n, sampleLists, out, scratch_keys, stream):
        rc = orig(ctx, params, n, sampleLists, out, scratch_keys, stream)
        first.append((n, scratch_keys, getattr(eng.lib, "place_variant", 0)))
        return rc

    eng.lib.maple_place_batch = spy
    rec = tree.place_samples(samples, p)
    for f in ("bestNode", "status", "phase1", "bLenTop", "bLenBottom", "bLenAppend"):
        assert np.array_equal(rec[f], ref[f]), f
    assert len(first) >= 2 and first[0][0] == 24 and first[1][0] < 24 and first[1][1] == 8 * 4096  # some samples were re-run with 8x
    assert first[0][2] == variant and all(c[2] == 0 for c in first[1:])  # the re-runs go through variant 0
    assert eng.place_variant == variant  # and the batch variant is restored


@pytest.mark.parametrize("name", ["ex_unrest", "ay_unrest_300"])
def test_search_seam_wrapper_on_reference_trees(name):
    """start_topology_updates_parallel (the drop-in for Pool.map(startTopologyUpdatesParallel) + sort, :12283-12312) over the
    stand-in library: prepare_search fills the zero-length root children, binds the tree, the searches run (here: the kernel's
    per-lane state machine on the host), scratch-exhausted searches are re-run with more, and the moves come back sorted as the
    reference sorts them -- equal to the oracle's, and to the reference's own proposedMoves where its lazy fill plays no role."""
    from maple_b200 import capi
    from maple_b200.search import dirty_nodes, start_topology_updates_parallel
    from oracle.oracle import Oracle
    from tree_fixture import search_params, searched_nodes
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    a = tree_arrays(g)
    eng = FakeEngine(model)
    tree = DeviceTree.from_lists(eng, a["up"], a["child0"], a["child1"], a["dist"], a["root"], a["isTip"], tree_lists(g),
                                 mutStart=a["mutStart"], mut=a["mut"], numMinor=a["numMinor"])
    p = capi.SearchParams()
    for k, v in search_params(g).items():
        setattr(p, k, v)
    t = g["tree"]
    nodes = dirty_nodes(tree, t["dirty"], t["replacements"], g["env"]["maxReplacements"])
    assert sorted(nodes.tolist()) == sorted(searched_nodes(g))
    tree.tree_likelihood()  # moves the arena tables (temporary lists): the search must notice and bind again
    moves, rec = start_topology_updates_parallel(tree, p, nodes, scratch_keys=256)  # small scratch: some searches are re-run
    ref = Oracle(model).search_batch(a, tree_lists(g), search_params(g), nodes, lazy_mode=1)
    for f in ("status", "placement", "bestNode", "phase1", "bLenTop", "bLenBottom", "bLenAppend"):
        assert np.array_equal(rec[f], ref[f]), f
    want = sorted(((int(n), int(r["placement"]), float(r["improvement"])) for n, r in zip(nodes, ref) if r["placement"] >= 0), key=lambda m: m[2])
    assert [m[:2] for m in moves] == [m[:2] for m in want] and all(abs(x[2] - y[2]) <= 1e-9 for x, y in zip(moves, want))
    assert moves == sorted(moves, key=lambda m: m[2])
    assert sorted(m[:2] for m in moves) == sorted((m[0], m[1]) for core in g["proposed"] for m in core)  # the reference's own proposedMoves
    assert (rec["status"] == 0).sum() > 50


def test_launch_shape_is_chosen_by_measurement(monkeypatch):
    """DeviceTree._critical_auto (tree.py): for a batch size the first sorted round runs plain and is timed, the next runs with the
    longest searches on SMs of their own and is timed, the faster shape is kept; a batch of another size starts over; searches
    without a recorded length only get measured.  The timers are stubbed -- the decision logic is what is tested here."""
    import torch
    g = load_golden("ex_unrest")
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    a = tree_arrays(g)
    tree = DeviceTree.from_lists(FakeEngine(model), a["up"], a["child0"], a["child1"], a["dist"], a["root"], a["isTip"], tree_lists(g),
                                 mutStart=a["mutStart"], mut=a["mut"], numMinor=a["numMinor"])
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)

    class Ev:
        def __init__(self, ms):
            self.ms = ms

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return other.ms

    def run(n, cost, ms):
        k, rec = tree._critical_auto(torch.tensor(cost, dtype=torch.int64), n)
        if rec is not None:
            rec["ev"] = (Ev(0.0), Ev(ms))  # what spr_search records around the launch
        return k, rec is not None

    big = torch.iinfo(torch.int64).max
    n = 1000
    cost = [900, 800, 500, 460, 300] + [10] * (n - 5)
    assert run(n, [big] * n, 0.0) == (0, False)            # nothing measured yet: plain, untimed
    assert run(n, cost, 400.0) == (0, True)                # plain, timed
    k, timed = run(n, cost, 300.0)                         # critical, timed: the searches at least half as long as the longest
    assert (k, timed) == (4, True)
    assert run(n, cost, 0.0) == (4, False) and tree._critical_state["choice"] == "on"   # 300 < 0.97 * 400: kept
    assert run(n, cost, 0.0) == (4, False)
    assert run(2 * n, cost + [10] * n, 500.0) == (0, True)  # another batch size: start over
    assert run(2 * n, cost + [10] * n, 499.0)[1] is True
    assert run(2 * n, cost + [10] * n, 0.0) == (0, False) and tree._critical_state["choice"] == "off"  # 499 is not < 0.97 * 500
    assert tree._critical_count(torch.tensor([5] * 10, dtype=torch.int64), 10) == 0  # too few searches for a second launch
