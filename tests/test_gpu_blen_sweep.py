"""DeviceTree.optimize_branch_lengths (traverseTreeToOptimizeBranchLengths with fastPass=True, :8727, as one maple_blen_batch
launch + the root-split scan) against sweeps recorded from the reference on the frozen tree of each fixture and on a copy with
perturbed lengths (tests/golden/extras).  MAT trees included (lists crossing a local-reference branch are re-referenced on the
device).  The CPU twin over the oracle is tests/test_blen_sweep_host.py.  Needs a GPU."""
import numpy as np
import pytest

from golden_io import load_extras, load_golden
from tree_fixture import tree_arrays, tree_lists

pytestmark = pytest.mark.gpu


def device_tree(g, shim):
    from maple_b200.engine import MapleEngine
    from maple_b200.model import MapleModel
    from maple_b200.tree import DeviceTree
    eng = MapleEngine(MapleModel.from_reference_snapshot(g["env"], g["model"]), 0)
    a = tree_arrays(shim)
    return DeviceTree.from_lists(eng, a["up"], a["child0"], a["child1"], a["dist"], a["root"], a["isTip"], tree_lists(shim),
                                 mutStart=a["mutStart"], mut=a["mut"], numMinor=a["numMinor"])


@pytest.mark.parametrize("name", ["ex_unrest", "ex_jc", "ex_unrest_rv_sse", "ex_unrest_err", "ay_unrest_300"])
@pytest.mark.parametrize("which", ["frozen", "perturbed"])
def test_device_fast_sweep_matches_reference(name, which):
    ex, g = load_extras(name), load_golden(name)
    if which == "frozen":
        shim, want = g, ex["sweeps"]["fastPass"]
    else:
        t = dict(ex["perturbed"])
        t["numMinor"] = [len(m) for m in t["minorSequences"]]
        shim, want = {"tree": t, "lists": ex["lists"], "env": g["env"]}, ex["sweeps"]["perturbed_fastPass"]
    tree = device_tree(g, shim)
    n_ids = tree.arena.n
    updates, dirty = tree.optimize_branch_lengths(g["env"]["effectivelyNon0BLen"], dirty=shim["tree"]["dirty"])
    assert updates == want["updates"]
    assert [float(x) for x in tree.dist] == want["dist"]  # bit-identical lengths
    assert np.array_equal(tree.d_dist.cpu().numpy(), tree.dist)
    assert [bool(x) for x in dirty] == want["dirty"]
    assert tree.arena.n == n_ids  # temporary lists released
    if which == "perturbed":
        assert updates > 50
