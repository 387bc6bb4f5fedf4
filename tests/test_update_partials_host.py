"""updatePartials (:5479) and the default, sequential mode of traverseTreeToOptimizeBranchLengths (:8727, fastPass=False) as the
device runs them (maple_b200/csrc/update.cuh: one lane, the reference's order, lists re-pointed in the arena), compiled for the
host, against sweeps RECORDED FROM THE REFERENCE (tests/golden/extras 'sweeps': 'sequential' on the frozen tree, 'perturbed_sequential'
on the copy with perturbed lengths and recalculated lists, 57-1 900 accepted changes each followed by updatePartials).  Every
estimate of the sweep reads the lists the previous updatePartials calls left behind, so equal final lengths -- bit for bit -- pin
the whole chain; the dirty flags and the number of updates must match too, and the lists the sweep leaves must give the same tree
log-likelihood as lists rebuilt from scratch (to the tolerance updatePartials stops propagating at)."""
import ctypes as C

import numpy as np
import pytest

from golden_io import golden_names, load_extras, load_golden
from hostsim import KernelSourceOnHost
from maple_b200 import blen_sweep
from maple_b200.genome_list import PackedLists, pack_lists
from maple_b200.model import MapleModel
from oracle.oracle import Oracle, _p
from tree_fixture import tree_arrays, tree_lists


def _writable(lists: PackedLists, slack_keys: int):
    """Arena arrays with room behind the tails."""
    capK, capP = int(lists.key.size) + slack_keys, int(lists.pay.size) + 6 * slack_keys
    key, pay = np.zeros(capK, np.uint32), np.zeros(capP, np.float64)
    key[: lists.key.size], pay[: lists.pay.size] = lists.key, lists.pay
    tails = np.array([(lists.key.size + 3) // 4 * 4, (lists.pay.size + 1) // 2 * 2], np.int64)
    return key, pay, lists.key_start.copy(), lists.pay_start.copy(), lists.nkeys.copy(), lists.npay.copy(), tails, capK, capP


def run_update(hs, ta, lists, dist, dirty, mode, entries=(), slack_keys=1 << 20):
    key, pay, ks, ps, nk, npay, tails, capK, capP = _writable(lists, slack_keys)
    cur = PackedLists(key, pay, ks, ps, nk, npay, lists.lRef, lists.U)
    t, keep = hs._tree_struct(ta, cur)
    dist = np.ascontiguousarray(dist, np.float64).copy()
    dirty = np.ascontiguousarray(dirty, np.uint8).copy()
    ent = np.ascontiguousarray(np.array(entries, np.int32).reshape(-1))
    upd = C.c_int32(0)
    f = hs.L.hs_update
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 9 + [C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    st = f(hs.mp, C.addressof(t), _p(key), _p(pay), _p(ks), _p(ps), _p(nk), _p(npay), _p(tails), capK, capP, _p(dist), _p(dirty), mode,
           len(ent) // 2, _p(ent), C.addressof(upd))
    return st, int(upd.value), dist, dirty, cur


def sequential_sweep(hs, orc, env, g_tree, lists_py, which_lists):
    """Root split by the plan of blen_sweep (as DeviceTree does it with batch calls), its two updatePartials, then the device loop."""
    shim = {"tree": g_tree, "lists": lists_py, "env": env}
    ta, packed = tree_arrays(shim), tree_lists(shim)
    dist, dirty = np.array(ta["dist"], np.float64), np.array(g_tree["dirty"], np.uint8)
    root, c0, c1 = ta["root"], ta["child0"], ta["child1"]
    cur = packed
    if c0[root] >= 0:
        a, b = int(c0[root]), int(c1[root])
        cand = blen_sweep.root_split_candidates(dist[a], dist[b], env["lRef"], env["effectivelyNon0BLen"])
        if cand is not None:
            mut = g_tree["mutations"]
            L = lambda i: lists_py[g_tree["probVect"][i]]  # noqa: E731
            v1 = orc.pass_branch(L(a), mut[a], True) if mut[a] else L(a)
            v2 = orc.pass_branch(L(b), mut[b], True) if mut[b] else L(b)
            cost = []
            for b1, b2 in zip(*cand):
                rv, lk = orc.merge(v1, float(b1), bool(ta["isTip"][a]), v2, float(b2), bool(ta["isTip"][b]), returnLK=True)
                if mut[root]:
                    rv = orc.pass_branch(rv, mut[root], True)
                cost.append(lk + orc.prob_root(rv))
            nb1, nb2 = blen_sweep.choose_root_split(np.array(cost), cand[0], dist[a], dist[b])
            for child, new in ((a, nb1), (b, nb2)):  # :8788-8797 (both lists name (root, 0), as the reference does)
                dist[child] = new
                ta["dist"] = dist
                st, _, dist, dirty, cur = run_update(hs, ta, cur, dist, dirty, 0, [(child, 2), (root, 0)])
                assert st == 0
    ta["dist"] = dist
    st, updates, dist, dirty, cur = run_update(hs, ta, cur, dist, dirty, 1)
    assert st == 0
    return updates, dist, dirty, cur, ta


@pytest.mark.parametrize("which", ["sequential", "perturbed_sequential"])
@pytest.mark.parametrize("name", golden_names())
def test_sequential_sweep_matches_reference(name, which):
    ex, g = load_extras(name), load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    orc, hs = Oracle(model, with_root_tables=True), KernelSourceOnHost(model, with_root_tables=True)
    if which == "sequential":
        tree = dict(g["tree"])
        tree["minorSequences"] = ex["frozen"]["minorSequences"]
        lists = g["lists"]
    else:
        tree = dict(ex["perturbed"])
        lists = ex["lists"]
    tree["numMinor"] = [len(m) for m in tree["minorSequences"]]
    want = ex["sweeps"][which]
    updates, dist, dirty, cur, ta = sequential_sweep(hs, orc, g["env"], tree, lists, which)
    assert updates == want["updates"]  # (the root split's own changes are counted by neither)
    assert [float(x) for x in dist] == want["dist"]
    assert [bool(x) for x in dirty] == want["dirty"]
    if which == "perturbed_sequential":
        assert updates > 50
