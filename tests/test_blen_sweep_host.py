"""traverseTreeToOptimizeBranchLengths(fastPass=True) (:8727) as a batch plan (maple_b200/blen_sweep.py) run over the CPU oracle,
against sweeps recorded from the reference (tests/golden/extras): on the frozen tree of each fixture (already optimised: nothing
moves, every node turns clean) and on a copy with perturbed lengths and recalculated lists (60-100 branches move)."""
import numpy as np
import pytest

from golden_io import golden_names, load_extras, load_golden
from maple_b200 import blen_sweep
from maple_b200.model import MapleModel
from oracle.oracle import Oracle


def host_sweep(orc, env, tree, lists, dirty):
    """The plan of DeviceTree.optimize_branch_lengths with the oracle's single-call functions."""
    n = len(tree["up"])
    child0 = np.array([c[0] if c else -1 for c in tree["children"]], np.int64)
    child1 = np.array([c[1] if c else -1 for c in tree["children"]], np.int64)
    dist = np.array(tree["dist"], np.float64)
    numMinor = [len(m) for m in tree["minorSequences"]]
    isTip = [(not tree["children"][i]) and numMinor[i] == 0 for i in range(n)]
    mut = tree["mutations"]
    L = lambda fam, i: None if tree[fam][i] is None else lists[tree[fam][i]]  # noqa: E731
    root = tree["root"]
    if child0[root] >= 0:
        c1, c2 = int(child0[root]), int(child1[root])
        cand = blen_sweep.root_split_candidates(dist[c1], dist[c2], env["lRef"], env["effectivelyNon0BLen"])
        if cand is not None:
            v1 = orc.pass_branch(L("probVect", c1), mut[c1], True) if mut[c1] else L("probVect", c1)
            v2 = orc.pass_branch(L("probVect", c2), mut[c2], True) if mut[c2] else L("probVect", c2)
            cost = []
            for b1, b2 in zip(*cand):
                rv, lk = orc.merge(v1, float(b1), isTip[c1], v2, float(b2), isTip[c2], returnLK=True)
                if mut[root]:
                    rv = orc.pass_branch(rv, mut[root], True)
                cost.append(lk + orc.prob_root(rv))
            dist[c1], dist[c2] = blen_sweep.choose_root_split(np.array(cost), cand[0], dist[c1], dist[c2])
    nodes = blen_sweep.sweep_nodes(tree["up"], child0, child1, root, dirty)
    best, false = np.zeros(len(nodes)), np.zeros(len(nodes), bool)
    for k, nd in enumerate(nodes):
        p = tree["up"][nd]
        upv = L("probVectUpRight", p) if tree["children"][p][0] == nd else L("probVectUpLeft", p)
        if mut[nd]:
            upv = orc.pass_branch(upv, mut[nd], False)
        r = orc.blen(upv, L("probVect", nd), isTip[nd])
        false[k], best[k] = r is None, (0.0 if r is None else r)
    new, changed, still = blen_sweep.accept(dist[nodes], best, false)
    dist[nodes] = new
    dirty = np.array(dirty, bool)
    dirty[nodes] = still
    return dist, dirty, int(changed.sum())


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("which", ["frozen", "perturbed"])
def test_fast_sweep_matches_reference(name, which):
    ex, g = load_extras(name), load_golden(name)
    orc = Oracle(MapleModel.from_reference_snapshot(g["env"], g["model"]), with_root_tables=True)
    if which == "frozen":
        tree = dict(g["tree"])
        tree["minorSequences"] = ex["frozen"]["minorSequences"]
        lists, want = g["lists"], ex["sweeps"]["fastPass"]
    else:
        tree, lists, want = ex["perturbed"], ex["lists"], ex["sweeps"]["perturbed_fastPass"]
    dist, dirty, updates = host_sweep(orc, g["env"], tree, lists, tree["dirty"])
    assert updates == want["updates"]
    assert [float(x) for x in dist] == want["dist"]
    assert [bool(x) for x in dirty] == want["dirty"]
    if which == "perturbed":
        assert updates > 50 and (np.array(tree["dist"]) != dist).sum() >= updates


def test_root_split_candidates_rounding():
    b1, b2 = blen_sweep.root_split_candidates(2.5 / 1000, 0.0, 1000, 1e-8)  # round(2.5) == 2 in python: 5 candidates
    assert len(b1) == 5 and b1[-1] == 2.0 / 1000 and abs(b2[-1] - 0.5 / 1000) < 1e-18
    assert blen_sweep.root_split_candidates(0.0, 1e-9, 1000, 1e-8) is None
    b1, b2 = blen_sweep.root_split_candidates(0.0002, 0.0001, 1000, 1e-8)  # tot = 0.3 mutations: 0, 0.3 (clamped), 0.3
    assert len(b1) == 3 and b1[0] == 0.0 and b1[1] == b1[2] and b2[1] == 0.0
    assert blen_sweep.choose_root_split(np.array([1.0, 3.0, 3.0]), b1, 0.0002, 0.0001)[0] == b1[1]


def test_accept_rule():
    dist = np.array([0.0, 0.0, 1e-4, 1e-4, 1e-4, 1e-4, 2e-4])
    best = np.array([0.0, 1e-4, 0.0, 1.005e-4, 1.02e-4, 0.5e-4, 9.9e9])
    false = np.array([True, False, True, False, False, False, True])
    new, changed, still = blen_sweep.accept(dist, best, false)
    assert changed.tolist() == [False, True, True, False, True, True, True] and still.tolist() == changed.tolist()
    assert new.tolist() == [0.0, 1e-4, 0.0, 1e-4, 1.02e-4, 0.5e-4, 0.0]
