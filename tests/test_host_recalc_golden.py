"""reCalculateAllGenomeLists restated over the oracle reproduces the four list families of the
reference's frozen trees bit-exactly (this pins the tree-level orchestration the device builder mirrors)."""
import pytest

from golden_io import golden_names, load_golden
from host_recalc import recalc_lists
from maple_b200.genome_list import lists_equal
from maple_b200.model import MapleModel
from oracle.oracle import Oracle


# The error-model fixtures are left out: under usingErrorRate the reference rewrites tip O-vectors in place
# (updateProbVectTerminalNode, :3966-4008, called from :6131 with the node's CURRENT minor-sequence list), so a
# snapshot of its tree is not a fixed point of its own recalculation at leaves that carry minor sequences.
@pytest.mark.parametrize("name", ["ay_unrest_300", "ex_unrest", "ex_unrest_rv", "ex_jc", "ex_gtr"])
def test_recalc_matches_reference_tree(name):
    g = load_golden(name)
    orc = Oracle(MapleModel.from_reference_snapshot(g["env"], g["model"]))
    t, L = g["tree"], g["lists"]
    n = len(t["up"])
    isTip = [len(t["children"][i]) == 0 and t["numMinor"][i] == 0 for i in range(n)]
    live, stack = [], [t["root"]]
    while stack:
        x = stack.pop()
        live.append(x)
        stack.extend(t["children"][x])
    tips = {i: L[t["probVect"][i]] for i in live if not t["children"][i]}
    lower, upR, upL, tot = recalc_lists(orc, t["up"], t["children"], t["dist"], t["mutations"], isTip, t["root"], tips)

    def ref(fam, i):
        j = t[fam][i]
        return None if j is None else L[j]

    bad = []
    for i in live:
        if t["children"][i]:
            if not lists_equal(lower[i], ref("probVect", i)):
                bad.append(("lower", i))
            if not lists_equal(upR.get(i), ref("probVectUpRight", i)):
                bad.append(("upR", i))
            if not lists_equal(upL.get(i), ref("probVectUpLeft", i)):
                bad.append(("upL", i))
        if i != t["root"] and not lists_equal(tot.get(i), ref("probVectTotUp", i)):
            bad.append(("tot", i))
    assert not bad, bad[:10]
