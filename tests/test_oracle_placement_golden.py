"""The oracle's findBestParentForNewSample (oracle/maple_oracle.c: or_place_sample, or_is_minor) against placements the
unmodified reference computed on its frozen trees (tests/golden/make_golden.py: harvest_placements) -- new samples derived
from placed ones (copies, one difference fewer / more, mixes of two samples).  Bar: same node, same "absorbed as a minor
sequence" verdicts, same number of candidate branches scored, same branch lengths, scores within 1e-9."""
import numpy as np
import pytest

from golden_io import golden_names, load_golden
from maple_b200.genome_list import pack_lists
from maple_b200.model import MapleModel
from tree_fixture import tree_arrays, tree_lists


def place_params(g):
    e, pe = g["env"], g["placeEnv"]
    return {"strictStopRules": int(pe["strictStopRules"]), "allowedFails": int(pe["allowedFails"]),
            "deeperSearchForLongBranches": int(bool(e["deeperSearchForLongBranches"])), "onlyFindIdentical": int(pe["onlyFindIdentical"]),
            "thresholdLogLK": pe["thresholdLogLK"], "thresholdLogLKoptimization": pe["thresholdLogLKoptimization"],
            "thresholdLogLKconsecutivePlacement": e["thresholdLogLKconsecutivePlacement"], "effectivelyNon0BLen": e["effectivelyNon0BLen"],
            "BLenThresholdDeeperSearch": e["BLenThresholdDeeperSearch"], "oneMutBLen": pe["oneMutBLen"]}


def check_placements(g, rec):
    for r, c in zip(rec, g["placements"]):
        assert r["status"] == (1 if c["minor"] else 0), (c["label"], int(r["status"]))
        assert r["bestNode"] == c["bestNode"], (c["label"], int(r["bestNode"]), c["bestNode"])
        if c["minor"]:
            assert r["bestScore"] == 1.0
            continue
        assert r["phase1"] == c["phase1"], (c["label"], int(r["phase1"]), c["phase1"])
        assert [r["bLenTop"], r["bLenBottom"], r["bLenAppend"]] == c["blens"], (c["label"], r, c["blens"])
        assert r["bestScore"] == c["bestScore"] or abs(r["bestScore"] - c["bestScore"]) <= 1e-9, (c["label"], float(r["bestScore"]), c["bestScore"])


@pytest.mark.parametrize("name", [n for n in golden_names() if "placements" in load_golden(n)])
def test_oracle_placement_matches_reference(name):
    from oracle.oracle import Oracle
    g = load_golden(name)
    assert len(g["placements"]) >= 20
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    rec = Oracle(model).place_batch(tree_arrays(g), tree_lists(g), place_params(g), samples)
    check_placements(g, rec)
    assert sum(c["minor"] for c in g["placements"]) > 0 and sum(not c["minor"] for c in g["placements"]) > 0


def test_is_minor_sequence_cases():
    """isMinorSequence (:5919) on hand-made lists (lRef from a fixture)."""
    from oracle.oracle import Oracle
    g = load_golden("ex_unrest")
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    L = model.lRef
    orc = Oracle(model)
    a = [(4, 100), (1, 0), (4, L)]                      # one substitution at 101
    b = [(4, 100), (5, 101), (4, L)]                    # same site unknown
    c = [(4, 100), (6, 0, [0.5, 0.5, 0.0, 0.0]), (4, L)]  # ambiguous A/C there
    d = [(4, 100), (2, 0), (4, L)]                      # a different substitution
    assert orc.is_minor(a, a) == 1 and orc.is_minor(a, a, True) == 1
    assert orc.is_minor(a, b) == 1 and orc.is_minor(b, a) == 2 and orc.is_minor(a, b, True) == 0
    assert orc.is_minor(a, c) == 1 and orc.is_minor(c, a) == 2
    assert orc.is_minor(a, d) == 0 and orc.is_minor(d, c) == 0
