"""Turn the frozen tree of a golden fixture into host arrays + one packed arena (list id = family*nNodes + node)."""
import numpy as np

from maple_b200.genome_list import pack_lists

FAMILIES = ("probVect", "probVectUpRight", "probVectUpLeft", "probVectTotUp")


def tree_arrays(g):
    t = g["tree"]
    n = len(t["up"])
    up = np.array([-1 if u is None else u for u in t["up"]], np.int32)
    child0 = np.array([c[0] if c else -1 for c in t["children"]], np.int32)
    child1 = np.array([c[1] if c else -1 for c in t["children"]], np.int32)
    dist = np.array(t["dist"], np.float64)
    isTip = np.array([(not t["children"][i]) and t["numMinor"][i] == 0 for i in range(n)], np.uint8)
    mutStart = np.zeros(n + 1, np.int32)
    mut = []
    for i in range(n):
        mut.extend(t["mutations"][i])
        mutStart[i + 1] = len(mut)
    mut = np.array(mut, np.int32).reshape(-1, 3) if mut else np.zeros((0, 3), np.int32)
    return {"up": up, "child0": child0, "child1": child1, "dist": dist, "isTip": isTip, "root": t["root"],
            "mutStart": mutStart, "mut": mut, "numMinor": np.array(t["numMinor"], np.int32)}


def tree_lists(g):
    t, L = g["tree"], g["lists"]
    lists = []
    for fam in FAMILIES:
        lists.extend(None if j is None else L[j] for j in t[fam])
    return pack_lists(lists, g["env"]["lRef"], g["env"]["usingErrorRate"])


def search_params(g):
    e, p = g["env"], g["params"]
    return {"strictTopologyStopRules": int(p["strict"]), "allowedFailsTopology": int(p["fails"]),
            "deeperSearchForLongBranches": int(bool(e["deeperSearchForLongBranches"])),
            "thresholdLogLKtopology": p["thr"], "thresholdTopologyPlacement": p["thrPlace"],
            "thresholdLogLKoptimizationTopology": e["thresholdLogLKoptimizationTopology"],
            "thresholdLogLKconsecutivePlacement": e["thresholdLogLKconsecutivePlacement"],
            "effectivelyNon0BLen": e["effectivelyNon0BLen"], "BLenThresholdDeeperSearch": e["BLenThresholdDeeperSearch"],
            "defaultBLen": e["defaultBLen"]}


def searched_nodes(g):
    """Nodes startTopologyUpdatesParallel visits (:9615-9619), in the order the reference's in-process run visited
    them: worker 0's nodes in its stack order, then worker 1's, ... (reachable, dirty, replacements<=max)."""
    t = g["tree"]
    out = []
    for core in range(g["params"]["numCores"]):
        stack = [t["root"]]
        while stack:
            n = stack.pop()
            stack.extend(t["children"][n])
            if (t["dirty"][n] and t["replacements"][n] <= g["env"]["maxReplacements"] and t["up"][n] is not None
                    and t["coreNum"][n] == core):
                out.append(n)
    return out


def compare_with_reference_searches(g, nodes, rec, lazy, pre, max_count_diff=3, min_fraction=0.98):
    """rec: records of the path under test (device, or the CUDA source on the host) for `nodes`, computed with probVectTotUp of the
    zero-length children of the root filled BEFORE the round; lazy / pre: the oracle's records in the reference's lazy mode and in
    the pre-filled mode.  The reference fills those lists lazily and order-dependently during the round (:7198-7200).  Over all
    recorded rounds (5 374 searches) the deviation shows as: a different number of scored candidates in searches that reach such a
    child (1 306 searches, by at most 3), and in ONE search (ex_gtr, pruned node 53, best node = that child) a refinement of the
    branch lengths the reference skips; proposedMoves never differ.  So: every search whose outcome does not depend on the fill
    order must equal the reference's record in node, lengths and score; those are at least 98 % of every round; the proposals must be identical."""
    import numpy as np
    fields = ("status", "placement", "bestNode", "bLenTop", "bLenBottom", "bLenAppend", "bestScore", "improvement")
    t = g["tree"]
    by_node = {int(n): (r, all(a[f] == b[f] or (a[f] != a[f] and b[f] != b[f]) for f in fields)) for n, r, a, b in zip(nodes, rec, lazy, pre)}
    checked = 0
    for s in g["searches"]:
        r, independent = by_node[t["children"][s["node"]][s["child"]]]
        if not independent:
            continue
        assert r["status"] == 0 and r["bestNode"] == s["bestNode"], (s, r)
        assert abs(int(r["phase1"]) - s["phase1"]) <= max_count_diff, (s, r)
        assert [r["bLenTop"], r["bLenBottom"], r["bLenAppend"]] == [float(x) for x in s["blens"]], (s, r)
        assert r["bestScore"] == s["bestScore"] or abs(r["bestScore"] - s["bestScore"]) <= 1e-9, (s, r)
        checked += 1
    assert checked >= min_fraction * len(g["searches"]), (checked, len(g["searches"]))
    root = t["root"]
    if all(t["dist"][c] > 0 for c in t["children"][root]):  # no zero-length child of the root: nothing to fill, counts identical
        assert int(np.sum(rec["phase1"])) == g["phase1Total"]
    got = sorted((int(n), int(r["placement"])) for n, (r, _) in by_node.items() if r["placement"] >= 0)
    assert got == sorted((m[0], m[1]) for core in g["proposed"] for m in core)
