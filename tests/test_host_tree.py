"""oracle/host_tree.py (batched CPU list building, used by bench.py's CPU arm) against the per-merge restatement of
reCalculateAllGenomeLists in tests/host_recalc.py, on synthetic trees with and without the error model."""
import pytest

from host_recalc import recalc_lists
from maple_b200.genome_list import lists_equal
from maple_b200.synthetic import generate


@pytest.mark.parametrize("rv,err,ml", [(False, False, False), (True, False, True), (True, True, False)])
def test_batched_builder_matches_per_merge_builder(rv, err, ml):
    from oracle.host_tree import build_tree_lists
    from oracle.oracle import Oracle
    d = generate(180, lRef=4000, mean_diffs=8.0, rate_variation=rv, error_model=err, site_specific_errors=err, seed=9, ml_like_blens=ml)
    orc = Oracle(d.model)
    pl, dist, isTip = build_tree_lists(orc, d.up, d.child0, d.child1, d.dist, d.root, d.tip_nodes, d.tip_lists, d.model.lRef,
                                       d.model.usingErrorRate)
    n = len(d.up)
    children = [[int(d.child0[i]), int(d.child1[i])] if d.child0[i] >= 0 else [] for i in range(n)]
    upl = [None if u < 0 else int(u) for u in d.up]
    lower, upR, upL, tot = recalc_lists(orc, upl, children, [float(x) for x in dist], [[] for _ in range(n)], [not c for c in children],
                                        d.root, {int(t): d.tip_lists[i] for i, t in enumerate(d.tip_nodes)})
    for i in range(n):
        for fam, ref in ((0, lower.get(i)), (1, upR.get(i)), (2, upL.get(i)), (3, tot.get(i))):
            got = pl.get(fam * n + i) if pl.key_start[fam * n + i] >= 0 else None
            if ref is None and fam == 3 and got is not None and dist[i] == 0.0:
                continue  # zero-length child of the root, pre-filled for the search like DeviceTree.prepare_search does
            assert lists_equal(got, ref), (fam, i)
