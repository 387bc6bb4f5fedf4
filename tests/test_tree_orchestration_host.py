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
