"""Device-resident list building (reCalculateAllGenomeLists as level-synchronous merge batches) against the
same orchestration run over the CPU oracle, on a synthetic tree -- needs a GPU."""
import numpy as np
import pytest

from host_recalc import recalc_lists
from maple_b200.genome_list import lists_equal, pack_lists

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rv,err", [(False, False), (True, False), (True, True)])
def test_device_recalc_matches_oracle(rv, err):
    import torch
    from maple_b200.engine import MapleEngine
    from maple_b200.synthetic import generate
    from maple_b200.tree import DeviceTree
    from maple_b200.workloads import neighbourhood_pairs
    from oracle.oracle import Oracle
    d = generate(150, lRef=4000, mean_diffs=8.0, rate_variation=rv, error_model=err, site_specific_errors=err, seed=5)
    eng = MapleEngine(d.model, 0)
    tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
    tips = pack_lists(d.tip_lists, d.model.lRef, d.model.usingErrorRate)
    tree.recalculate_all_lists(d.tip_nodes, tips)
    orc = Oracle(d.model)
    n = len(d.up)
    children = [[int(d.child0[i]), int(d.child1[i])] if d.child0[i] >= 0 else [] for i in range(n)]
    upl = [None if u < 0 else int(u) for u in d.up]
    isTip = [not children[i] for i in range(n)]
    lower, upR, upL, tot = recalc_lists(orc, upl, children, [float(x) for x in d.dist], [[] for _ in range(n)], isTip, d.root,
                                        {int(t): d.tip_lists[i] for i, t in enumerate(d.tip_nodes)})
    for i in range(n):
        got = tree.lists_of(i)
        assert lists_equal(got[0], lower[i]), ("lower", i)
        assert lists_equal(got[1], upR.get(i)), ("upR", i)
        assert lists_equal(got[2], upL.get(i)), ("upL", i)
        assert lists_equal(got[3], tot.get(i)), ("tot", i)
    # the candidate batch built on top of it scores identically on GPU and oracle
    s, p, c, tip, bl = neighbourhood_pairs(tree, 5)
    got = eng.append_prob_batch(p, c, tip, bl).cpu().numpy()
    host = tree.arena.to_host()
    ref = orc.append_batch(host, p.cpu().numpy(), c.cpu().numpy(), tip.cpu().numpy(), bl.cpu().numpy())
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin) and fin.sum() > 100
    assert np.max(np.abs(got[fin] - ref[fin])) <= 1e-9
