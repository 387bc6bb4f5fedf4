"""findBestParentForNewSample under stop rules the main fixtures lack, recorded from the reference (tests/golden/extras
'placements_more', make_golden.py: harvest_more_placements): the non-strict form of its default rule and a tight one, on the frozen
tree and on the copy with perturbed branch lengths (5 x 56 placements per configuration).  The oracle, the straight-line CUDA
source and the warp forms (variants 2 and 3: these trees carry MAT mutations) must reproduce every placement: node, absorbed /
placed verdict, candidate count, branch lengths; scores within 1e-9."""
import pytest

from golden_io import golden_names, load_extras, load_golden
from hostsim import KernelSourceOnHost
from maple_b200.genome_list import pack_lists
from maple_b200.model import MapleModel
from oracle.oracle import Oracle
from test_oracle_placement_golden import check_placements, place_params
from tree_fixture import tree_arrays, tree_lists

KEYS = ["frozen_nonstrict", "frozen_tight", "perturbed_default", "perturbed_nonstrict", "perturbed_tight"]


def _shim(name, key):
    g, ex = load_golden(name), load_extras(name)
    pm = ex["placements_more"]
    if key.startswith("frozen"):
        tree_shim = g
    else:
        t = dict(ex["perturbed"])
        t["numMinor"] = [len(m) for m in t["minorSequences"]]
        tree_shim = {"tree": t, "lists": ex["lists"], "env": g["env"]}
    s = {"env": g["env"], "placeEnv": pm[key]["placeEnv"], "placements": pm[key]["placements"], "lists": pm["lists"]}
    return g, tree_shim, s


@pytest.mark.parametrize("key", KEYS)
@pytest.mark.parametrize("name", golden_names())
def test_placements_under_other_rules(name, key):
    g, tree_shim, s = _shim(name, key)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    ta, lists = tree_arrays(tree_shim), tree_lists(tree_shim)
    samples = pack_lists([[tuple(e) for e in s["lists"][c["diffs"]]] for c in s["placements"]], model.lRef, model.usingErrorRate)
    pp = place_params(s)
    check_placements(s, Oracle(model).place_batch(ta, lists, pp, samples))
    if name in ("ex_unrest", "ex_unrest_err", "ay_unrest_300", "ay_unrest_deep_200", "ay_unrest_1000"):
        hs = KernelSourceOnHost(model)
        check_placements(s, hs.place_batch(ta, lists, pp, samples, scratch_keys=1 << 16))
        for mat in (1, 2):
            check_placements(s, hs.place_batch_scan(ta, lists, pp, samples, scratch_keys=1 << 16, mat=mat))
    assert sum(not c["minor"] for c in s["placements"]) > 10
