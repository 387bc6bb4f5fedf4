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
