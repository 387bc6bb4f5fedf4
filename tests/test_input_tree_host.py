"""--inputTree + --input end to end on the CPU side: Newick and alignment files -> tips and minor-sequence collapse
(maple_b200.newick) -> the four list families and the tree likelihood built by the oracle in level-synchronous batches
(oracle/host_tree.py, the orchestration DeviceTree runs on the device) == what the reference's own set-up of the same files
gave (reCalculateAllGenomeLists(firstSetUp=True) + calculateTreeLikelihood, recorded in tests/golden/extras)."""
import numpy as np
import pytest

from golden_io import load_extras, load_golden
from maple_b200.genome_list import lists_equal
from maple_b200.model import MapleModel
from maple_b200.newick import is_minor_sequence, load_input_tree
from oracle.host_tree import build_tree_lists, tree_likelihood
from oracle.oracle import Oracle

FAMILIES = ("probVect", "probVectUpRight", "probVectUpLeft", "probVectTotUp")


def load(name, tmp_path):
    ex, g = load_extras(name), load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    nwk, aln = tmp_path / "t.nwk", tmp_path / "a.txt"
    nwk.write_text(ex["newick"]["binary"] + "\n")
    aln.write_text(ex["alignmentText"])
    out = load_input_tree(str(nwk), str(aln), model, default_blen=g["env"]["defaultBLen"],
                          only_find_identical=g["placeEnv"]["onlyFindIdentical"], only_n_ambiguities=g["tipInputs"]["onlyNambiguities"])
    return ex, g, model, out


# ex_unrest_rv_sse is left out: there the reference's collapse depends on the history of its ambiguity table (test_newick_io)
@pytest.mark.parametrize("name", ["ex_unrest", "ex_gtr", "ex_jc", "ex_unrest_rv", "ex_unrest_err", "ay_unrest_300", "ay_unrest_deep_200"])
def test_input_tree_lists_and_likelihood_match_reference(name, tmp_path):
    ex, g, model, (t, root, names, tip_nodes, tip_lists) = load(name, tmp_path)
    want = ex["read"]["binary"]["loaded"]
    a = t.arrays()
    orc = Oracle(model, with_root_tables=True)
    pl, dist, isTip = build_tree_lists(orc, a["up"], a["child0"], a["child1"], a["dist"], root, tip_nodes, tip_lists, model.lRef,
                                       int(model.usingErrorRate), isTip=a["isTip"])
    n = len(t)
    assert [float(x) for x in dist] == want["dist"]  # including the zero-length repairs of the set-up (oneMutBLen/2, :6181)
    bad = []
    for i in t.reachable(root):
        for f, fam in enumerate(FAMILIES):
            j = want[fam][i]
            if fam == "probVectTotUp" and a["dist"][i] == 0 and a["up"][i] == root:
                continue  # filled ahead of the search here, lazily by the reference
            if not lists_equal(pl.get(f * n + i), None if j is None else ex["lists"][j]):
                bad.append((fam, i))
    assert not bad, bad[:10]
    lk = tree_likelihood(orc, pl, a["child0"], a["child1"], dist, root, isTip, a["numMinor"])
    assert abs(lk - ex["read"]["binary"]["loadedLK"]) <= 1e-9 * abs(lk)
    # the frozen tree it was written from had MAT local references and its own minor-sequence history: close, not equal
    assert abs(lk - g["treeLK"]) <= 1e-4 * abs(lk)


@pytest.mark.parametrize("name", ["ex_unrest", "ex_unrest_err", "ay_unrest_300"])
def test_host_is_minor_sequence_matches_oracle(name):
    """The loader's python isMinorSequence against the pinned C oracle on all pairs of recorded tips."""
    g = load_golden(name)
    orc = Oracle(MapleModel.from_reference_snapshot(g["env"], g["model"]))
    tips = [g["lists"][t["list"]] for t in g["tipInputs"]["tips"]][:40]
    seen = set()
    for only in (False, True):
        for v1 in tips:
            for v2 in tips:
                r = is_minor_sequence(v1, v2, g["env"]["lRef"], only)
                assert r == orc.is_minor(v1, v2, only)
                seen.add(r)
    assert seen == {0, 1, 2}
