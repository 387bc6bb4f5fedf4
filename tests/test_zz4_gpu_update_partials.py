"""Device twin of tests/test_update_partials_host.py: maple_update_partials / maple_blen_sweep_sequential through the C ABI
(DeviceTree.update_partials, DeviceTree.optimize_branch_lengths_sequential) against the sequential branch-length sweeps recorded
from the reference (traverseTreeToOptimizeBranchLengths with fastPass=False: every accepted change followed by updatePartials):
lengths and dirty flags bit for bit, the number of updates, and a search round on the lists the sweep leaves behind."""
import numpy as np
import pytest

from golden_io import hw_names as golden_names, load_extras, load_golden
from maple_b200.model import MapleModel
from tree_fixture import tree_arrays, tree_lists

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["sequential", "perturbed_sequential"])
@pytest.mark.parametrize("name", golden_names())
def test_device_sequential_sweep_matches_reference(name, which):
    from maple_b200.engine import MapleEngine
    from maple_b200.tree import DeviceTree
    ex, g = load_extras(name), load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    eng = MapleEngine(model, 0)
    if which == "sequential":
        tree = dict(g["tree"])
        tree["minorSequences"] = ex["frozen"]["minorSequences"]
        lists = g["lists"]
    else:
        tree, lists = dict(ex["perturbed"]), ex["lists"]
    tree["numMinor"] = [len(m) for m in tree["minorSequences"]]
    shim = {"tree": tree, "lists": lists, "env": g["env"]}
    ta = tree_arrays(shim)
    dt = DeviceTree.from_lists(eng, ta["up"], ta["child0"], ta["child1"], ta["dist"], ta["root"], ta["isTip"], tree_lists(shim),
                               ta["mutStart"], ta["mut"], ta["numMinor"])
    want = ex["sweeps"][which]
    updates, dirty = dt.optimize_branch_lengths_sequential(g["env"]["effectivelyNon0BLen"], tree["dirty"])
    assert updates == want["updates"]
    assert [float(x) for x in dt.dist] == want["dist"]
    assert [bool(x) for x in dirty] == want["dirty"]
    lk = dt.tree_likelihood()  # the lists the sweep left are consistent lists of the new lengths
    assert np.isfinite(lk)
    if want.get("treeLK") is not None:
        assert abs(lk - want["treeLK"]) <= 1e-6, (lk, want["treeLK"])


def test_update_partials_after_one_edit():
    """One branch length changed by hand, then updatePartials from both ends of the branch as the reference does (:8875): the
    lists it leaves give the tree likelihood that lists rebuilt from scratch give (to the tolerance updatePartials stops at)."""
    from maple_b200.engine import MapleEngine
    from maple_b200.genome_list import pack_lists
    from maple_b200.synthetic import generate
    from maple_b200.tree import DeviceTree
    d = generate(600, rate_variation=True, seed=9, ml_like_blens=True)
    eng = MapleEngine(d.model, 0)
    tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
    tips = pack_lists(d.tip_lists, d.model.lRef, d.model.usingErrorRate)
    tree.recalculate_all_lists(d.tip_nodes, tips)
    lk0 = tree.tree_likelihood()
    node = int(np.nonzero((tree.up >= 0) & (tree.up != tree.root) & (tree.dist > 0))[0][17])
    tree.dist[node] *= 3.0
    tree.d_dist.copy_(__import__("torch").from_numpy(tree.dist))
    parent = int(tree.up[node])
    tree.update_partials([(node, 2), (parent, 0 if tree.child0[parent] == node else 1)])
    lk1 = tree.tree_likelihood()
    tree.recalculate_all_lists(d.tip_nodes, tips)
    lk2 = tree.tree_likelihood()
    assert lk1 != lk0 and abs(lk1 - lk2) <= 1e-4 * max(1.0, abs(lk2) * 1e-6), (lk0, lk1, lk2)
