"""CUDA kernels (through the C ABI) against the golden vectors and the CPU oracle -- needs a GPU.

Bar: merged lists, branch lengths, list comparisons bit-exact; log-likelihood scores within
1e-9 absolute (device log() vs libm log(): <= 1 ulp of each folded factor).
"""
import numpy as np
import pytest

from golden_io import hw_names as golden_names, load_golden
from maple_b200.genome_list import lists_equal, pack_lists
from maple_b200.model import MapleModel

pytestmark = pytest.mark.gpu
LK_TOL = 1e-9


@pytest.fixture(scope="module", params=golden_names())
def env(request):
    import torch
    from maple_b200.engine import MapleEngine
    from oracle.oracle import Oracle
    assert torch.cuda.is_available()
    g = load_golden(request.param)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    eng = MapleEngine(model, 0)
    packed = pack_lists(g["lists"], model.lRef, model.usingErrorRate)
    eng.bind(packed)
    return g, eng, packed, Oracle(model)


def test_append_vs_golden_and_oracle(env):
    g, eng, packed, orc = env
    calls = g["calls"]["appendProbNode"]
    p = np.array([c["P"] for c in calls], np.int32)
    c_ = np.array([c["C"] for c in calls], np.int32)
    tip = np.array([c["isTipC"] for c in calls], np.uint8)
    bl = np.array([c["bLen"] for c in calls], np.float64)
    exp = np.array([c["out"] for c in calls], np.float64)
    got = eng.append_prob_batch(p, c_, tip, bl).cpu().numpy()
    got_h = eng.append_prob_batch_host(p, c_, tip, bl)
    ref = orc.append_batch(packed, p, c_, tip, bl)
    assert np.array_equal(got, got_h)
    inf = np.isinf(exp)
    assert np.array_equal(np.isinf(got), inf) and np.all(got[inf] < 0)
    assert np.array_equal(np.isinf(ref), inf)
    assert np.max(np.abs(got[~inf] - exp[~inf])) <= LK_TOL
    assert np.max(np.abs(got[~inf] - ref[~inf])) <= LK_TOL


def test_append_random_pairs_vs_oracle(env):
    g, eng, packed, orc = env
    rng = np.random.default_rng(7)
    n = 20000
    nl = len(packed)
    p = rng.integers(0, nl, n).astype(np.int32)
    c_ = rng.integers(0, nl, n).astype(np.int32)
    tip = rng.integers(0, 2, n).astype(np.uint8)
    bl = np.where(rng.random(n) < 0.2, 0.0, rng.random(n) * 3e-4)
    got = eng.append_prob_batch(p, c_, tip, bl).cpu().numpy()
    ref = orc.append_batch(packed, p, c_, tip, bl)
    inf = np.isinf(ref)
    assert np.array_equal(np.isinf(got), inf)
    assert np.max(np.abs(got[~inf] - ref[~inf])) <= LK_TOL * np.maximum(1.0, np.abs(ref[~inf]) * 1e-3).max()


def test_merge_vs_golden(env):
    g, eng, packed, orc = env
    L = g["lists"]
    calls = g["calls"]["mergeVectors"]
    fl = np.array([(1 if c["isUpDown"] else 0) | (2 if c["returnLK"] else 0) for c in calls], np.uint8)
    r = eng.merge_batch([c["v1"] for c in calls], [c["b1"] for c in calls], [c["t1"] for c in calls], [c["v2"] for c in calls],
                        [c["b2"] for c in calls], [c["t2"] for c in calls], fl, [c["numMinor1"] for c in calls],
                        [c["numMinor2"] for c in calls])
    outs = r.to_lists()
    lk = r.lk.cpu().numpy()
    for i, c in enumerate(calls):
        exp = None if c["out"] is None else L[c["out"]]
        assert lists_equal(outs[i], exp), (i, c)
        if c["returnLK"]:
            assert abs(lk[i] - c["lk"]) <= LK_TOL


def test_merge_random_pairs_vs_oracle(env):
    g, eng, packed, orc = env
    rng = np.random.default_rng(11)
    n = 4000
    nl = len(packed)
    i1 = rng.integers(0, nl, n).astype(np.int32)
    i2 = rng.integers(0, nl, n).astype(np.int32)
    b1 = np.where(rng.random(n) < 0.25, 0.0, rng.random(n) * 2e-4)
    b2 = np.where(rng.random(n) < 0.25, 0.0, rng.random(n) * 2e-4)
    t1 = rng.integers(0, 2, n).astype(np.uint8)
    t2 = rng.integers(0, 2, n).astype(np.uint8)
    fl = rng.integers(0, 2, n).astype(np.uint8)  # isUpDown on/off, no LK (LK needs lower lists)
    for shorten in (False, True):
        r = eng.merge_batch(i1, b1, t1, i2, b2, t2, fl, shorten=shorten)
        o = orc.merge_batch(packed, i1, b1, t1, i2, b2, t2, fl)
        st = r.status.cpu().numpy()
        assert np.array_equal(st != 0, o["status"] != 0)
        key, pay = r.key.cpu().numpy().view(np.uint32), r.pay.cpu().numpy()
        ks, ps, nk, npay = (x.cpu().numpy() for x in (r.key_start, r.pay_start, r.nkeys, r.npay))
        if not shorten:
            assert np.array_equal(nk, o["nkeys"]) and np.array_equal(npay, o["npay"])
            for i in range(n):
                if st[i] == 0:
                    assert np.array_equal(key[ks[i]:ks[i] + nk[i]], o["key"][o["key_start"][i]:o["key_start"][i] + nk[i]])
                    assert np.array_equal(pay[ps[i]:ps[i] + npay[i]], o["pay"][o["pay_start"][i]:o["pay_start"][i] + npay[i]])
        else:
            outs = r.to_lists()
            for i in range(0, n, 9):
                if st[i] == 0:
                    from maple_b200.genome_list import decode_stream
                    un = decode_stream(o["key"], o["pay"], o["key_start"][i], o["pay_start"][i], packed.lRef, packed.U)
                    assert lists_equal(outs[i], orc.shorten(un))


def test_blen_vs_golden_and_oracle(env):
    g, eng, packed, orc = env
    calls = g["calls"]["estimateBranchLengthWithDerivative"]
    p = np.array([c["P"] for c in calls], np.int32)
    c_ = np.array([c["C"] for c in calls], np.int32)
    tip = np.array([c["fromTipC"] for c in calls], np.uint8)
    out, st = eng.blen_batch(p, c_, tip)
    out, st = out.cpu().numpy(), st.cpu().numpy()
    for i, c in enumerate(calls):
        if c["out"] is None:
            assert st[i] == 1
        else:
            assert st[i] == 0 and out[i] == c["out"], (i, c, out[i])
    rng = np.random.default_rng(5)
    n = 5000
    p = rng.integers(0, len(packed), n).astype(np.int32)
    c_ = rng.integers(0, len(packed), n).astype(np.int32)
    tip = rng.integers(0, 2, n).astype(np.uint8)
    out, st = eng.blen_batch(p, c_, tip)
    ro, rs = orc.blen_batch(packed, p, c_, tip)
    assert np.array_equal(st.cpu().numpy(), rs)
    assert np.array_equal(out.cpu().numpy(), ro)


def test_differ_vs_golden_and_oracle(env):
    g, eng, packed, orc = env
    calls = [c for c in g["calls"]["areVectorsDifferent"] if c["v2"] is not None]
    got = eng.vectors_differ_batch([c["v1"] for c in calls], [c["v2"] for c in calls]).cpu().numpy()
    assert np.array_equal(got.astype(bool), np.array([c["out"] for c in calls]))
    rng = np.random.default_rng(3)
    n = 20000
    a = rng.integers(0, len(packed), n).astype(np.int32)
    b = np.where(rng.random(n) < 0.3, a, rng.integers(0, len(packed), n)).astype(np.int32)
    got = eng.vectors_differ_batch(a, b).cpu().numpy()
    assert np.array_equal(got, orc.differ_batch(packed, a, b))


def test_reference_named_single_calls(env):
    g, eng, packed, orc = env
    L = g["lists"]
    c = g["calls"]["appendProbNode"][0]
    v = eng.appendProbNode(L[c["P"]], L[c["C"]], c["isTipC"], c["bLen"])
    assert (v == c["out"]) or abs(v - c["out"]) <= LK_TOL
    c = g["calls"]["mergeVectors"][0]
    out = eng.mergeVectors(L[c["v1"]], c["b1"], c["t1"], L[c["v2"]], c["b2"], c["t2"], returnLK=c["returnLK"], isUpDown=c["isUpDown"])
    if c["returnLK"]:
        out = out[0]
    assert lists_equal(out, None if c["out"] is None else L[c["out"]])
    c = g["calls"]["estimateBranchLengthWithDerivative"][0]
    r = eng.estimateBranchLengthWithDerivative(L[c["P"]], L[c["C"]], c["fromTipC"])
    assert (r is False and c["out"] is None) or r == c["out"]
    c = g["calls"]["areVectorsDifferent"][0]
    assert eng.areVectorsDifferent(L[c["v1"]], None if c["v2"] is None else L[c["v2"]]) == c["out"]
    eng.bind(packed)


def test_prob_root_vs_golden_and_oracle(env):
    g, eng, packed, orc = env
    from oracle.oracle import Oracle
    orc_r = Oracle(eng.model, with_root_tables=True)
    calls = g["calls"]["findProbRoot"]
    assert calls
    idx = np.array([c["v"] for c in calls], np.int32)
    got = eng.prob_root_batch(idx).cpu().numpy()
    for c, x in zip(calls, got):
        assert abs(x - c["out"]) <= 1e-6, (c, x)  # north star: 1e-6 absolute on log-likelihoods
        assert abs(x - orc_r.prob_root(g["lists"][c["v"]])) <= LK_TOL * max(1.0, abs(c["out"]) * 1e-3)
    assert abs(eng.findProbRoot(g["lists"][calls[0]["v"]]) - got[0]) == 0.0


def test_root_vector_vs_golden(env):
    """maple_root_vector_batch against every rootVector call the reference recorded (:4916-4996): lists bit-identical."""
    g, eng, packed, orc = env
    calls = g["calls"]["rootVector"]
    assert calls
    r = eng.root_vector_batch([c["v"] for c in calls], [float(c["bLen"]) if c["bLen"] else 0.0 for c in calls],
                              [1 if c["isFromTip"] else 0 for c in calls])
    for got, c in zip(r.to_lists(), calls):
        assert lists_equal(got, g["lists"][c["out"]]), c
        assert lists_equal(got, orc.shorten(orc.root_vector(g["lists"][c["v"]], c["bLen"], c["isFromTip"]))), c
    c = calls[0]
    assert lists_equal(eng.rootVector(g["lists"][c["v"]], c["bLen"], c["isFromTip"]), g["lists"][c["out"]])


def test_pass_branch_vs_golden(env):
    g, eng, packed, orc = env
    calls = g["calls"]["passGenomeListThroughBranch"]
    for c in calls[:40]:
        got = eng.passGenomeListThroughBranch(g["lists"][c["v"]], c["mutations"], c["dirIsUp"])
        assert lists_equal(got, g["lists"][c["out"]]), c


def test_tree_likelihood_vs_reference(env):
    """calculateTreeLikelihood of the reference's frozen tree (:9721-9779), computed on the device from the stored lists."""
    g, eng, packed, orc = env
    from maple_b200.engine import MapleEngine
    from maple_b200.tree import DeviceTree
    from tree_fixture import tree_arrays, tree_lists
    eng2 = MapleEngine(eng.model, 0)
    ta = tree_arrays(g)
    tree = DeviceTree.from_lists(eng2, ta["up"], ta["child0"], ta["child1"], ta["dist"], ta["root"], ta["isTip"], tree_lists(g),
                                 ta["mutStart"], ta["mut"], ta["numMinor"])
    lk = tree.tree_likelihood()
    assert abs(lk - g["treeLK"]) <= 1e-6, (lk, g["treeLK"])
