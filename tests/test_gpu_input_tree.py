"""--inputTree + --input end to end with the likelihood work on the device: files -> host set-up (maple_b200.newick) ->
DeviceTree.from_host_tree (reCalculateAllGenomeLists as level-synchronous merge batches) -> tree_likelihood, against what the
reference's own set-up of the same files produced (tests/golden/extras).  The CPU twin of this test, over the oracle, is
tests/test_input_tree_host.py.  Needs a GPU."""
import pytest

from golden_io import load_extras, load_golden
from maple_b200.genome_list import lists_equal

pytestmark = pytest.mark.gpu

FAMILIES = ("probVect", "probVectUpRight", "probVectUpLeft", "probVectTotUp")


@pytest.mark.parametrize("name", ["ex_unrest", "ex_jc", "ex_unrest_rv", "ex_unrest_err", "ay_unrest_300"])
def test_device_input_tree_matches_reference(name, tmp_path):
    from maple_b200.engine import MapleEngine
    from maple_b200.model import MapleModel
    from maple_b200.newick import load_input_tree
    from maple_b200.tree import DeviceTree
    ex, g = load_extras(name), load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    nwk, aln = tmp_path / "t.nwk", tmp_path / "a.txt"
    nwk.write_text(ex["newick"]["binary"] + "\n")
    aln.write_text(ex["alignmentText"])
    t, root, names, tip_nodes, tip_lists = load_input_tree(str(nwk), str(aln), model, default_blen=g["env"]["defaultBLen"],
                                                           only_find_identical=g["placeEnv"]["onlyFindIdentical"],
                                                           only_n_ambiguities=g["tipInputs"]["onlyNambiguities"])
    eng = MapleEngine(model, 0)
    tree = DeviceTree.from_host_tree(eng, t, root, tip_nodes, tip_lists)
    want = ex["read"]["binary"]["loaded"]
    assert [float(x) for x in tree.dist] == want["dist"]
    bad = []
    for i in t.reachable(root):
        got = tree.lists_of(i)
        for f, fam in enumerate(FAMILIES):
            j = want[fam][i]
            if fam == "probVectTotUp" and tree.dist[i] == 0 and tree.up[i] == root:
                continue
            if not lists_equal(got[f], None if j is None else ex["lists"][j]):
                bad.append((fam, i))
    assert not bad, bad[:10]
    lk = tree.tree_likelihood()
    assert abs(lk - ex["read"]["binary"]["loadedLK"]) <= 1e-6  # north-star tolerance on log-likelihoods
