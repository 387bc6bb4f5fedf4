"""The 1 000-sequence reference tree (tests/golden/ay_unrest_1000: 556 searches and 128 k candidates per strict round, 295 k under the
deep rules on the perturbed copy) on the device: searches of all four kernel variants, both extra rounds with the default kernel,
and the recorded placements.  Same bars as tests/test_gpu_search.py / test_gpu_placement.py.  The fixture was recorded after the
GPU budget of round 1 was spent (CPU twins: test_oracle_*_golden.py, test_kernel_source_host.py, test_search_rounds_golden.py), so
this file is collected last.  Needs a GPU."""
import numpy as np
import pytest

from golden_io import load_golden
from maple_b200.genome_list import pack_lists
from maple_b200.model import MapleModel
from test_gpu_placement import _capi_params as _place_params
from test_gpu_search import _capi_params, _compare
from test_oracle_placement_golden import check_placements, place_params
from test_search_rounds_golden import round_shim
from tree_fixture import compare_with_reference_searches, search_params, searched_nodes, tree_arrays, tree_lists

pytestmark = pytest.mark.gpu
NAME = "ay_unrest_1000"


def _tree(g, s):
    from maple_b200.engine import MapleEngine
    from maple_b200.tree import DeviceTree
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    eng = MapleEngine(model, 0)
    ta, lists = tree_arrays(s), tree_lists(s)
    tree = DeviceTree.from_lists(eng, ta["up"], ta["child0"], ta["child1"], ta["dist"], ta["root"], ta["isTip"], lists,
                                 ta["mutStart"], ta["mut"], ta["numMinor"])
    return model, eng, tree, ta, lists


@pytest.mark.parametrize("variant,rnd", [(0, "main"), (1, "main"), (2, "main"), (3, "main"), (4, "main"), (4, "frozen_deep"), (0, "frozen_deep"), (0, "perturbed_deep"),
                                         (0, "perturbed_fast")])
def test_searches_on_the_big_tree(variant, rnd):
    from oracle.oracle import Oracle
    if rnd == "main":
        g = s = load_golden(NAME)
    else:
        g, s = round_shim(NAME, rnd)
    model, eng, tree, ta, lists = _tree(g, s)
    eng.set_search_variant(variant)
    nodes = np.array(searched_nodes(s), np.int32)
    tree.prepare_search()
    big = variant == 1  # the straight-line kernel does not re-run searches that exhaust their scratch
    rec = tree.search_records(tree.spr_search(nodes, _capi_params(search_params(s)), scratch_keys=(1 << 15) if big else 0,
                                              max_concurrent=4096 if big else 0))
    orc = Oracle(model)
    pre = orc.search_batch(ta, lists, search_params(s), nodes, lazy_mode=1)
    _compare(rec, pre, nodes)
    compare_with_reference_searches(s, nodes, rec, orc.search_batch(ta, lists, search_params(s), nodes, lazy_mode=0), pre)


def test_placements_on_the_big_tree():
    g = load_golden(NAME)
    model, eng, tree, ta, lists = _tree(g, g)
    samples = pack_lists([g["lists"][c["diffs"]] for c in g["placements"]], model.lRef, model.usingErrorRate)
    check_placements(g, tree.place_samples(samples, _place_params(place_params(g))))
