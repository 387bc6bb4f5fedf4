"""Randomised cross-check of the CUDA source (compiled for the host, tests/hostsim) against the C oracle: two independent
restatements of the reference's list functions must agree on inputs the recorded vectors do not contain.  Inputs grow organically:
start from the lists of a reference tree, merge random pairs with branch lengths drawn from the corners (0, 1e-9, a fraction of a
mutation, a few mutations, 0.1) in both modes, feed the results back into the pools, and run every function on random pairs from the
pools.  Lists bit-identical, scalars within 1e-9 (equal when infinite)."""
import json
import random

import pytest

from golden_io import load_golden
from hostsim import KernelSourceOnHost
from maple_b200.genome_list import lists_equal
from maple_b200.model import MapleModel
from oracle.oracle import Oracle


def _close(a, b):
    if a is None or b is None:
        return a is None and b is None
    return a == b or abs(a - b) <= 1e-9


@pytest.mark.parametrize("name,seed", [("ex_unrest", 1), ("ex_unrest_rv_sse", 2), ("ex_unrest_err", 3), ("ay_unrest_300", 4), ("ex_jc", 5)])
def test_random_operation_chains_agree(name, seed):
    g = load_golden(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    orc, hs = Oracle(model, with_root_tables=True), KernelSourceOnHost(model, with_root_tables=True)
    rng = random.Random(seed)
    t, L = g["tree"], g["lists"]
    clean = [i for i in range(len(t["up"])) if not t["mutations"][i]]  # lists in the reference genome's coordinates
    lower = [L[t["probVect"][i]] for i in clean if t["probVect"][i] is not None][:120]
    upper = [L[t[f][i]] for f in ("probVectUpRight", "probVectUpLeft", "probVectTotUp") for i in clean if t[f][i] is not None][:200]
    lRef = model.lRef

    def blen():
        return rng.choice([0.0, 0.0, 1e-9, rng.random() / lRef, 3 * rng.random() / lRef, 0.1 * rng.random(), 0.1])

    n_none = n_inf = n_false = 0
    for step in range(2500):
        op = rng.randrange(8)
        if op == 0:  # lower x lower
            a, b = rng.choice(lower), rng.choice(lower)
            args = (a, blen(), rng.random() < 0.3, b, blen(), rng.random() < 0.3)
            kw = {"returnLK": rng.random() < 0.3, "numMinor1": rng.choice([0, 0, 2]), "numMinor2": rng.choice([0, 0, 1])}
            r1, r2 = orc.merge(*args, **kw), hs.merge(*args, **kw)
            if kw["returnLK"] and r1 is not None:
                assert r2 is not None and lists_equal(r1[0], r2[0]) and _close(r1[1], r2[1]), (step, args, kw)
                out = r1[0]
            else:
                assert lists_equal(r1, r2), (step, args, kw)
                out = r1
            if out is None:
                n_none += 1
            elif len(lower) < 400:
                lower.append(orc.shorten(out))
        elif op == 1:  # upper x lower
            a, b = rng.choice(upper), rng.choice(lower)
            args = (a, blen(), False, b, blen(), rng.random() < 0.3)
            r1, r2 = orc.merge(*args, isUpDown=True), hs.merge(*args, isUpDown=True)
            assert lists_equal(r1, r2), (step, args)
            if r1 is None:
                n_none += 1
            elif len(upper) < 500:
                upper.append(orc.shorten(r1))
        elif op == 2:
            a, b, tip, bl = rng.choice(upper), rng.choice(lower), rng.random() < 0.5, blen()
            r1, r2 = orc.append(a, b, tip, bl), hs.append(a, b, tip, bl)
            assert _close(r1, r2), (step, r1, r2)
            assert r2 == hs.append_variant("sitewise", a, b, tip, bl) == hs.append_variant("q4", a, b, tip, bl)
            n_inf += r1 == float("-inf")
        elif op == 3:
            a, b, tip = rng.choice(upper), rng.choice(lower), rng.random() < 0.5
            r1, r2 = orc.blen(a, b, tip), hs.blen(a, b, tip)
            assert r1 == r2, (step, r1, r2)  # bit-identical lengths
            n_false += r1 is None
        elif op == 4:
            pool = rng.choice([lower, upper])
            a, b = rng.choice(pool), rng.choice(pool)
            assert orc.differ(a, b) == hs.differ(a, b) and not hs.differ(a, a)
        elif op == 5:
            a, bl, tip = rng.choice(lower), blen(), rng.random() < 0.5
            assert lists_equal(orc.root_vector(a, bl, tip), hs.root_vector(a, bl, tip))
        elif op == 6:
            a = rng.choice(lower)
            assert _close(orc.prob_root(a), hs.prob_root(a))
        else:
            a = rng.choice(rng.choice([lower, upper]))
            assert lists_equal(orc.shorten(a), hs.shorten(a))
    assert len(lower) > 200 and len(upper) > 300  # the pools did grow


def _check_against(res, want):
    assert len(res) == len(want)
    for k, (a, b) in enumerate(zip(res, want)):
        assert a[0] == b[0], k
        for x, y in zip(a[1:], b[1:]):
            if isinstance(y, float) and isinstance(x, float):
                assert x == y or abs(x - y) <= 1e-9, (k, a, b)
            else:
                assert x == y, (k, a, b)  # digests of lists (bit-exact content), booleans, None, "-inf"
        if a[0] == 3 and a[1] is not None:
            assert a[1] == b[1], (k, a, b)  # branch lengths are bit-identical


@pytest.mark.parametrize("backend", ["oracle", "cuda-source"])
@pytest.mark.parametrize("name", ["ex_unrest", "ex_gtr", "ex_jc", "ex_unrest_rv", "ex_unrest_rv_sse", "ex_unrest_err", "ay_unrest_300",
                                  "ay_unrest_deep_200", "ay_unrest_1000"])
def test_random_chain_matches_the_reference(name, backend):
    """The same chain was run with the REFERENCE's own functions when the fixtures were made (make_golden.py: harvest_fuzz): 2 500
    operations per configuration on lists the reference's run never produced.  The oracle and the CUDA source must reproduce every
    output list bit for bit and every scalar within 1e-9."""
    import fuzz_chain
    from golden_io import load_extras
    g, ex = load_golden(name), load_extras(name)
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    be = fuzz_chain.OracleBackend((Oracle if backend == "oracle" else KernelSourceOnHost)(model, with_root_tables=True))
    t, L = g["tree"], g["lists"]
    lower, upper = fuzz_chain.initial_pools(t, lambda fam, i: None if t[fam][i] is None else L[t[fam][i]])
    fz = ex["fuzz"]
    res = fuzz_chain.run_chain(be, lower, upper, model.lRef, fz["seed"], fz["steps"])
    res = json.loads(json.dumps(res))  # tuples -> lists, like the stored results
    _check_against(res, fz["results"])
    assert sum(1 for r in fz["results"] if r[0] in (0, 1) and r[1] is None) > 5  # impossible merges were met too
