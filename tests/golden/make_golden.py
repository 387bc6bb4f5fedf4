#!/usr/bin/env python3
"""Generate golden fixtures by RUNNING the unmodified reference (MAPLEv0.7.5.4.py).

This script only runs in the build container (it needs /root/reference); the fixtures it
writes under tests/golden/*.json.gz are committed and are what the tests read.  Nothing
here copies reference source: the reference script is executed with runpy as __main__,
with multiprocessing.Pool replaced by an in-process stand-in so that the frozen tree and
the model that the reference hands to startTopologyUpdatesParallel (MAPLEv0.7.5.4.py:12289)
can be snapshotted, and the module-level likelihood functions can be wrapped by recorders.

What is recorded per config (see CONFIGS):
  * env:    lRef, reference string, rootFreqs, flags, thresholds (module globals at the
            time of the first Pool.map, i.e. what forked workers would inherit)
  * model:  mutMatrixGlobal, errorRateGlobal, siteRates / errorRates when in use
  * tree:   up/children/dist/mutations/minor counts/dirty/replacements/coreNum and the four
            genome-list families, snapshotted BEFORE the searches run
  * calls:  sampled (inputs -> output) of appendProbNode, mergeVectors,
            estimateBranchLengthWithDerivative, areVectorsDifferent,
            passGenomeListThroughBranch, rootVector, findProbRoot, shorten
  * searches: every findBestParentTopology call of the first parallel round
            (args -> bestNode, bestScore, branch lengths, #phase-1 candidates)
  * proposed: the proposedMoves list returned by startTopologyUpdatesParallel per core
  * treeLK: calculateTreeLikelihood of the frozen tree, final LK of the run

Usage:  python tests/golden/make_golden.py [config ...]
"""
import gzip
import io
import json
import os
import runpy
import sys
import contextlib
import multiprocessing

REF = "/root/reference/MAPLEv0.7.5.4.py"
EX = "/root/reference/example_files/"
HERE = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    # name: (input file, max seqs (None = all), extra argv, per-function sample cap)
    "ex_unrest": (EX + "MAPLE_alignment_example.txt", None, ["--model", "UNREST"], 400),
    "ex_gtr": (EX + "MAPLE_alignment_example.txt", None, ["--model", "GTR"], 150),
    "ex_jc": (EX + "MAPLE_alignment_example.txt", None, ["--model", "JC"], 150),
    "ex_unrest_rv": (EX + "MAPLE_alignment_example.txt", None, ["--model", "UNREST", "--rateVariation"], 250),
    "ex_unrest_rv_sse": (EX + "MAPLE_alignment_example.txt", None,
                         ["--model", "UNREST", "--rateVariation", "--estimateSiteSpecificErrorRate"], 400),
    "ex_unrest_err": (EX + "MAPLE_alignment_example.txt", None, ["--model", "UNREST", "--estimateErrorRate"], 250),
    "ay_unrest_300": (EX + "sameRef_AY.4.2.2.maple.gz", 300, ["--model", "UNREST"], 300),
    "ay_unrest_1000": (EX + "sameRef_AY.4.2.2.maple.gz", 1000, ["--model", "UNREST", "--rateVariation"], 40),
    "ay_unrest_deep_200": (EX + "sameRef_AY.4.2.2.maple.gz", 200,
                           ["--model", "UNREST", "--deeperSearchForLongBranches"], 100),
    # alignments written by the bench's own generator (maple_b200/synthetic.py: "synthetic:<nSeq>[:rv][:err][:sse]"), so that the
    # reference pins the device on the kind of data bench.py measures, with and without the error model
    "syn_unrest_rv_2000": ("synthetic:2000:rv", None, ["--model", "UNREST", "--rateVariation"], 40),
    "syn_unrest_rv_sse_1500": ("synthetic:1500:rv:err:sse", None,
                               ["--model", "UNREST", "--rateVariation", "--estimateSiteSpecificErrorRate"], 40),
}

FUNCS = ["appendProbNode", "mergeVectors", "estimateBranchLengthWithDerivative", "areVectorsDifferent",
         "passGenomeListThroughBranch", "rootVector", "findProbRoot", "shorten"]


def truncate_input(path, max_seqs, out):
    if path.startswith("synthetic:"):  # the bench's generator, seed 1 (the tree it simulates is not used: the reference infers its own)
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        from maple_b200.synthetic import generate, write_maple_file
        opts = path.split(":")[1:]
        d = generate(int(opts[0]), rate_variation="rv" in opts, error_model="err" in opts, site_specific_errors="sse" in opts, seed=1,
                     ml_like_blens=True)
        write_maple_file(d, out)
        return out
    op = gzip.open if path.endswith(".gz") else open
    n = -1  # the reference record counts as the first '>'
    with op(path, "rt") as f, open(out, "w") as g:
        for line in f:
            if line.startswith(">"):
                n += 1
                if max_seqs is not None and n > max_seqs:
                    break
            g.write(line)
    return out


class ListTable:
    """Interns genome lists by content so that the fixture stores each distinct list once."""

    def __init__(self):
        self.index = {}
        self.lists = []

    @staticmethod
    def canon(gl):
        out = []
        for e in gl:
            ee = []
            for x in e:
                if isinstance(x, (list, tuple)):
                    ee.append(tuple(float(v) for v in x))
                elif isinstance(x, bool):
                    ee.append(bool(x))
                else:
                    ee.append(x)
            out.append(tuple(ee))
        return tuple(out)

    def add(self, gl):
        if gl is None:
            return None
        c = self.canon(gl)
        i = self.index.get(c)
        if i is None:
            i = len(self.lists)
            self.index[c] = i
            self.lists.append(c)
        return i


def jsonable_list(c):
    return [[list(x) if isinstance(x, tuple) else x for x in e] for e in c]


class Recorder:
    def __init__(self, G, cap):
        self.G = G
        self.cap = cap
        self.table = ListTable()
        self.calls = {f: [] for f in FUNCS}
        self.count = {f: 0 for f in FUNCS}
        self.searches = []
        self.phase1 = 0
        self.orig = {}
        self.nspecial = {}

    def keep(self, f, special):
        self.count[f] += 1
        n = self.count[f]
        if special:
            self.nspecial[f] = self.nspecial.get(f, 0) + 1
            if self.nspecial[f] <= self.cap // 4:
                return True
        # dense at the start, then geometric thinning
        if len(self.calls[f]) >= self.cap:
            return False
        return n <= self.cap // 2 or n % 37 == 0

    def install(self):
        G, T = self.G, self.table
        rec = self

        o_app = G["appendProbNode"]

        def appendProbNode(P, C, isTipC, bLen, **kw):
            r = o_app(P, C, isTipC, bLen, **kw)
            ln = sys._getframe(1).f_lineno
            if ln in (7011, 7223):
                rec.phase1 += 1
            if rec.keep("appendProbNode", r == float("-inf")):
                ip, ic = T.add(P), T.add(C)
                rec.calls["appendProbNode"].append({"P": ip, "C": ic, "isTipC": bool(isTipC), "bLen": bLen, "out": r, "line": ln})
            return r

        o_mer = G["mergeVectors"]

        def mergeVectors(v1, b1, t1, v2, b2, t2, returnLK=False, isUpDown=False, numMinor1=0, numMinor2=0, **kw):
            r = o_mer(v1, b1, t1, v2, b2, t2, returnLK=returnLK, isUpDown=isUpDown, numMinor1=numMinor1, numMinor2=numMinor2, **kw)
            if rec.keep("mergeVectors", r is None or returnLK):
                i1, i2 = T.add(v1), T.add(v2)
                if returnLK:
                    out, lk = T.add(r[0]), r[1]
                else:
                    out, lk = T.add(r), None
                rec.calls["mergeVectors"].append({"v1": i1, "b1": b1, "t1": bool(t1), "v2": i2, "b2": b2, "t2": bool(t2),
                                                  "returnLK": bool(returnLK), "isUpDown": bool(isUpDown),
                                                  "numMinor1": numMinor1, "numMinor2": numMinor2, "out": out, "lk": lk,
                                                  "line": sys._getframe(1).f_lineno})
            return r

        o_bl = G["estimateBranchLengthWithDerivative"]

        def estimateBranchLengthWithDerivative(P, C, fromTipC=False, **kw):
            r = o_bl(P, C, fromTipC=fromTipC, **kw)
            if rec.keep("estimateBranchLengthWithDerivative", r is False or r == 0.1):
                ip, ic = T.add(P), T.add(C)
                rec.calls["estimateBranchLengthWithDerivative"].append(
                    {"P": ip, "C": ic, "fromTipC": bool(fromTipC), "out": (None if r is False else r)})
            return r

        o_df = G["areVectorsDifferent"]

        def areVectorsDifferent(v1, v2):
            r = o_df(v1, v2)
            if rec.keep("areVectorsDifferent", False):
                i1, i2 = T.add(v1), T.add(v2)
                rec.calls["areVectorsDifferent"].append({"v1": i1, "v2": i2, "out": bool(r)})
            return r

        o_ps = G["passGenomeListThroughBranch"]

        def passGenomeListThroughBranch(v, mutations, dirIsUp=False):
            r = o_ps(v, mutations, dirIsUp=dirIsUp)
            if rec.keep("passGenomeListThroughBranch", False):
                i1 = T.add(v)
                rec.calls["passGenomeListThroughBranch"].append(
                    {"v": i1, "mutations": [list(m) for m in mutations], "dirIsUp": bool(dirIsUp), "out": T.add(r)})
            return r

        o_rv = G["rootVector"]

        def rootVector(v, bLen, isFromTip, tree, node, **kw):
            r = o_rv(v, bLen, isFromTip, tree, node, **kw)
            # only root-relative calls without MAT mutations on the path are self-contained
            n, clean = node, True
            while n is not None:
                if tree.mutations[n]:
                    clean = False
                n = tree.up[n]
            if clean and rec.keep("rootVector", False):
                i1 = T.add(v)
                rec.calls["rootVector"].append({"v": i1, "bLen": bLen, "isFromTip": bool(isFromTip), "out": T.add(r)})
            return r

        o_sh = G["shorten"]

        def shorten(v):
            before = T.canon(v)
            o_sh(v)
            if rec.keep("shorten", False):
                rec.calls["shorten"].append({"v": T.add(before), "out": T.add(v)})

        o_fb = G["findBestParentTopology"]

        def findBestParentTopology(tree, node, child, bestLKdiff, removedBLen, **kw):
            before = rec.phase1
            r = o_fb(tree, node, child, bestLKdiff, removedBLen, **kw)
            rec.searches.append({"node": node, "child": child, "bestLKdiff": bestLKdiff, "removedBLen": removedBLen,
                                 "strict": bool(kw.get("strictTopologyStopRules")), "fails": kw.get("allowedFailsTopology"),
                                 "thr": kw.get("thresholdLogLKtopology"),
                                 "bestNode": r[0], "bestScore": r[1], "blens": list(r[2]), "phase1": rec.phase1 - before})
            return r

        new = {"appendProbNode": appendProbNode, "mergeVectors": mergeVectors,
               "estimateBranchLengthWithDerivative": estimateBranchLengthWithDerivative,
               "areVectorsDifferent": areVectorsDifferent, "passGenomeListThroughBranch": passGenomeListThroughBranch,
               "rootVector": rootVector, "shorten": shorten, "findBestParentTopology": findBestParentTopology}
        for k, v in new.items():
            self.orig[k] = G[k]
            G[k] = v

    def uninstall(self):
        for k, v in self.orig.items():
            self.G[k] = v
        self.orig = {}


def snapshot_tree(tree, root, T):
    n = len(tree.up)
    return {
        "root": root,
        "up": list(tree.up),
        "children": [list(c) if c else [] for c in tree.children],
        "dist": [float(d) if d else 0.0 for d in tree.dist],
        "mutations": [[list(m) for m in (mm or [])] for mm in tree.mutations],
        "numMinor": [len(m) if m else 0 for m in tree.minorSequences],
        "dirty": [bool(d) for d in tree.dirty],
        "replacements": list(tree.replacements),
        "coreNum": list(getattr(tree, "coreNum", [None] * n)),
        "probVect": [T.add(v) for v in tree.probVect],
        "probVectUpRight": [T.add(v) for v in tree.probVectUpRight],
        "probVectUpLeft": [T.add(v) for v in tree.probVectUpLeft],
        "probVectTotUp": [T.add(v) for v in tree.probVectTotUp],
    }


TIPS_OUT = None


def harvest_placements(G, tree, root, T, n_base=14):
    """findBestParentForNewSample (:7912) on the FROZEN tree for new samples derived from the ones already placed: an exact
    copy, a copy without its last difference, a copy with one extra substitution, and a mix of two samples.  Every call
    runs on a deep copy of the tree (the reference appends absorbed samples to minorSequences and may update lists)."""
    import copy
    import random
    rng = random.Random(12345)
    lRef, ref = G["lRef"], G["ref"]
    _, data = G["readConciseAlignment"](G["inputFile"])  # the reference clears its own copy after the initial placement (:11816)
    data = {k: v for k, v in data.items() if v is not None}
    names = sorted(data)
    picks = [names[i] for i in sorted(rng.sample(range(len(names)), min(n_base, len(names))))]

    def covered(diffs):
        c = set()
        for d in diffs:
            ln = d[2] if len(d) > 2 else 1
            c.update(range(d[1], d[1] + ln))
        return c

    new_samples = []
    for i, nm in enumerate(picks):
        d = list(data[nm])
        new_samples.append(("copy:" + nm, d))
        if d:
            new_samples.append(("minus:" + nm, d[:-1]))
        cov = covered(d)
        for _ in range(50):
            pos = rng.randrange(1, lRef + 1)
            if pos not in cov:
                alt = rng.choice([b for b in "acgt" if b != ref[pos - 1].lower()])
                new_samples.append(("plus:" + nm, sorted(d + [(alt, pos)], key=lambda x: x[1])))
                break
        other = list(data[picks[(i + 1) % len(picks)]])
        half = lRef // 2
        mix = [x for x in d if x[1] + (x[2] if len(x) > 2 else 1) - 1 <= half] + [x for x in other if x[1] > half]
        new_samples.append(("mix:" + nm, mix))
    counter = {"n": 0}
    o_app = G["appendProbNode"]

    def counting_append(P, C, isTipC, bLen, **kw):
        if sys._getframe(1).f_lineno in (8033, 8050):
            counter["n"] += 1
        return o_app(P, C, isTipC, bLen, **kw)

    out = []
    G["appendProbNode"] = counting_append
    try:
        for label, diffs in new_samples:
            tcopy = copy.deepcopy(tree)
            partials = G["probVectTerminalNode"](diffs, None, None)
            before = T.canon(partials)
            counter["n"] = 0
            r = G["findBestParentForNewSample"](tcopy, root, partials, "new_" + label, False)
            minor = r[2] is None
            out.append({"label": label, "diffs": T.add(before), "bestNode": r[0], "bestScore": r[1], "minor": bool(minor),
                        "blens": None if minor else [float(x) if x else 0.0 for x in r[2]], "phase1": counter["n"]})
    finally:
        G["appendProbNode"] = o_app
    # inputs of the path: the reference's reader on a snippet of the input file, and probVectTerminalNode (:3882) of the
    # first samples under the flags that are live at this point of the run
    tips = []
    for nm in names[:80]:
        tips.append({"name": nm, "diffs": [list(x) for x in data[nm]], "list": T.add(T.canon(G["probVectTerminalNode"](data[nm], None, None)))})
    global TIPS_OUT
    TIPS_OUT = {"tips": tips, "onlyNambiguities": bool(G["onlyNambiguities"]), "usingErrorRate": bool(G["usingErrorRate"]),
                "errorRateSiteSpecific": bool(G["errorRateSiteSpecific"]), "errorRateGlobal": G.get("errorRateGlobal")}
    env = {"strictStopRules": bool(G["strictStopRules"]), "allowedFails": G["allowedFails"], "thresholdLogLK": G["thresholdLogLK"],
           "thresholdLogLKoptimization": G["thresholdLogLKoptimization"], "oneMutBLen": G["oneMutBLen"],
           "onlyFindIdentical": bool(G["errorRateSiteSpecificFile"] or G["errorRateFixed"] or G["estimateErrorRate"]
                                     or G["estimateSiteSpecificErrorRate"] or G["supportFor0Branches"] or G["HnZ"])}
    print("[golden] placements: %d new samples, %d absorbed as minor sequences, %d candidate branches scored" % (
        len(out), sum(o["minor"] for o in out), sum(o["phase1"] for o in out)), file=sys.stderr)
    return out, env


def _plain_tree(tree, root, T=None):
    """Topology / lengths / names of a reference Tree as plain lists (children None = node removed by the minor-sequence
    collapse of reCalculateAllGenomeLists(firstSetUp=True), :6108)."""
    out = {"root": root, "up": list(tree.up), "children": [None if c is None else list(c) for c in tree.children],
           "dist": [float(d) if d else 0.0 for d in tree.dist], "name": list(tree.name),
           "minorSequences": [list(m) for m in tree.minorSequences], "dirty": [bool(d) for d in tree.dirty],
           "replacements": list(getattr(tree, "replacements", [0] * len(tree.up))),
           "coreNum": list(getattr(tree, "coreNum", [None] * len(tree.up)))}
    if T is not None:
        for fam in ("probVect", "probVectUpRight", "probVectUpLeft", "probVectTotUp"):
            out[fam] = [T.add(v) for v in getattr(tree, fam)]
        out["mutations"] = [[list(m) for m in (mm or [])] for mm in tree.mutations]
    return out


def harvest_extras(G, tree, root, name):
    """Inputs/outputs of the callers either side of the search, all on deep copies of the frozen tree:
      * createNewick (:2816) of the frozen tree, binary and multifurcating; readNewick (:1812) + makeTreeBinary (:2117) of
        both strings; reCalculateAllGenomeLists(firstSetUp=True) (:6013) of the re-read binary tree from the alignment
        (without MAT local references) and its calculateTreeLikelihood;
      * traverseTreeToOptimizeBranchLengths (:8727) with fastPass=True (lists frozen: one batch of
        estimateBranchLengthWithDerivative + the root split) and in its default sequential mode."""
    import copy
    T = ListTable()
    names = G["namesInTree"]
    ex = {"namesInTree": list(names), "frozen": _plain_tree(tree, root)}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        nw = {"binary": G["createNewick"](tree, root, binary=True, namesInTree=names, estimateMAT=False, networkOutput=False, aBayesPlusOn=False),
              "multi": G["createNewick"](tree, root, binary=False, namesInTree=names, estimateMAT=False, networkOutput=False, aBayesPlusOn=False)}
    ex["newick"] = nw
    ex["read"] = {}
    with open(G["inputFile"]) as f:  # the (truncated) MAPLE-format input of this run, reference genome first
        ex["alignmentText"] = f.read()
    _, data = G["readConciseAlignment"](G["inputFile"])
    for kind, s in nw.items():
        path = "/tmp/golden_%s_%s.nwk" % (name, kind)
        with open(path, "w") as f:
            f.write(s + "\n")
        try:
            with contextlib.redirect_stdout(buf):
                trees, namesRead, namesDict = G["readNewick"](path, createDict=True)
        except ValueError as e:
            # the reference's own multifurcating output is unreadable by its reader when a zero-length tip carries minor
            # sequences ("name:0.0,minor:0.0:0.0", :2925-2946): record that instead
            ex["read"][kind] = {"error": repr(e)}
            continue
        t1, r1 = trees[0]
        rd = {"namesInTree": list(namesRead), "raw": _plain_tree(t1, r1)}
        G["makeTreeBinary"](t1, r1)
        rd["binary"] = _plain_tree(t1, r1)
        if kind == "binary":
            saved = (G["useLocalReference"], G["numMinorsRemoved"][0])
            G["useLocalReference"] = False
            try:
                with contextlib.redirect_stdout(buf):
                    G["reCalculateAllGenomeLists"](t1, r1, data=dict(data), names=namesRead, firstSetUp=True)
                    lk = G["calculateTreeLikelihood"](t1, r1)
            finally:
                G["useLocalReference"] = saved[0]
                G["numMinorsRemoved"][0] = saved[1]
            rd["loaded"] = _plain_tree(t1, r1, T)
            rd["loadedLK"] = lk
        ex["read"][kind] = rd
    sweeps = {}
    for mode, kw in (("fastPass", {"fastPass": True}), ("sequential", {})):
        tc = copy.deepcopy(tree)
        with contextlib.redirect_stdout(buf):
            upd = G["traverseTreeToOptimizeBranchLengths"](tc, root, **kw)
            lk = G["calculateTreeLikelihood"](tc, root) if mode == "sequential" else None
        sweeps[mode] = {"updates": upd, "dist": [float(d) if d else 0.0 for d in tc.dist], "dirty": [bool(d) for d in tc.dirty], "treeLK": lk}
    # the frozen tree has just been optimised (every estimate within 1 % of the stored length): also sweep a copy whose
    # positive lengths were scaled by 0.4 .. 2.2 and whose zero-length branches were partly opened, lists recalculated
    tp = copy.deepcopy(tree)
    for i in range(len(tp.dist)):
        if tp.up[i] is None or tp.children[i] is None:
            continue
        if tp.dist[i]:
            tp.dist[i] = tp.dist[i] * (0.4 + 0.3 * (i % 7))
        elif i % 3 == 0:
            tp.dist[i] = G["oneMutBLen"] * (1 + i % 4) / 2
        tp.dirty[i] = (i % 5 != 0)
    with contextlib.redirect_stdout(buf):
        G["reCalculateAllGenomeLists"](tp, root)
    ex["perturbed"] = _plain_tree(tp, root, T)
    for mode, kw in (("fastPass", {"fastPass": True}), ("sequential", {})):
        tc = copy.deepcopy(tp)
        with contextlib.redirect_stdout(buf):
            upd = G["traverseTreeToOptimizeBranchLengths"](tc, root, **kw)
        sweeps["perturbed_" + mode] = {"updates": upd, "dist": [float(d) if d else 0.0 for d in tc.dist], "dirty": [bool(d) for d in tc.dirty]}
    ex["sweeps"] = sweeps
    ex["lists"] = [jsonable_list(c) for c in T.lists]
    print("[golden] %s extras: newick %d/%d chars, loaded LK %r, sweep updates fast %d / sequential %d, perturbed %d / %d" % (
        name, len(nw["binary"]), len(nw["multi"]), ex["read"]["binary"]["loadedLK"], sweeps["fastPass"]["updates"],
        sweeps["sequential"]["updates"], sweeps["perturbed_fastPass"]["updates"], sweeps["perturbed_sequential"]["updates"]), file=sys.stderr)
    ex["_perturbed_tree"] = tp
    return ex


def harvest_more_placements(G, trees, root):
    """findBestParentForNewSample under stop rules the main fixture lacks (it has the reference's defaults: strict, 5 fails, 18 ln L):
    the non-strict form of the same rule, and a tight one (strict, 1 fail, 2 ln L), on the frozen tree and on the copy with perturbed
    branch lengths.  Same derived samples as the main fixture; own list table."""
    import math
    T = ListTable()
    out = {}
    L = math.log(G["lRef"])
    saved = (G["strictStopRules"], G["allowedFails"], G["thresholdLogLK"])
    global TIPS_OUT
    tips_saved = TIPS_OUT
    try:
        for tname, tr in trees.items():
            for rname, (strict, fails, thr) in (("default", saved), ("nonstrict", (False, saved[1], saved[2])), ("tight", (True, 1, 2.0 * L))):
                if tname == "frozen" and rname == "default":
                    continue  # the main fixture
                G["strictStopRules"], G["allowedFails"], G["thresholdLogLK"] = strict, fails, thr
                pl, env = harvest_placements(G, tr, root, T)
                out[tname + "_" + rname] = {"placements": pl, "placeEnv": env}
    finally:
        G["strictStopRules"], G["allowedFails"], G["thresholdLogLK"] = saved
        TIPS_OUT = tips_saved
    out["lists"] = [jsonable_list(c) for c in T.lists]
    return out


def harvest_rounds(G, func, inputs, trees, root):
    """startTopologyUpdatesParallel (:9580) of the reference under the stop rules the main fixture does not cover: the DEEP rules of the
    later rounds (module globals strictTopologyStopRules / allowedFailsTopology / thresholdLogLKtopology, :12155-12159) on the frozen
    tree, and both settings on the copy with perturbed branch lengths (recalculated lists, part of the nodes dirty).  Every search is
    recorded as in the main fixture; each run gets its own deep copy (the reference fills probVectTotUp lazily while it searches)."""
    import copy
    tail = tuple(inputs[0][7:])
    settings = {"deep": (G["strictTopologyStopRules"], G["allowedFailsTopology"], G["thresholdLogLKtopology"], G["thresholdTopologyPlacement"]),
                "fast": tuple(inputs[0][3:7])}
    out = {}
    buf = io.StringIO()
    for tname, tr in trees.items():
        for sname, (strict, fails, thr, thrPlace) in settings.items():
            if tname == "frozen" and sname == "fast":
                continue  # the main fixture
            tc = copy.deepcopy(tr)
            rec = Recorder(G, 0)
            rec.install()
            try:
                with contextlib.redirect_stdout(buf):
                    results = [func((tc, root, core, strict, fails, thr, thrPlace) + tail) for core in range(len(inputs))]
            finally:
                rec.uninstall()
            out[tname + "_" + sname] = {"params": {"strict": bool(strict), "fails": fails, "thr": thr, "thrPlace": thrPlace, "numCores": len(inputs)},
                                        "searches": rec.searches, "proposed": [[list(m) for m in r] for r in results],
                                        "phase1Total": rec.phase1}
            print("[golden] round %s/%s: %d searches, %d candidates, %d proposals" % (
                tname, sname, len(rec.searches), rec.phase1, sum(len(r) for r in results)), file=sys.stderr)
    return out


def harvest_fuzz(G, tree, root, seed, steps=2500):
    """A random chain of list operations (tests/fuzz_chain.py) run with the REFERENCE's own functions under the model of this run:
    digests of every output list and every scalar, for inputs far outside what the run itself produced."""
    import copy
    sys.path.insert(0, os.path.dirname(HERE))
    import fuzz_chain

    class Ref:
        def merge(self, a, b1, t1, b, b2, t2, lk, updown, nm1, nm2):
            try:
                return G["mergeVectors"](a, b1, t1, b, b2, t2, returnLK=lk, isUpDown=updown, numMinor1=nm1, numMinor2=nm2)
            except Exception as e:  # with returnLK an impossible merge raises Exception("exit") instead of returning None (:4753-4758)
                if e.args != ("exit",):
                    raise
                return None

        def append(self, a, b, tip, bl):
            return G["appendProbNode"](a, b, tip, bl)

        def blen(self, a, b, tip):
            r = G["estimateBranchLengthWithDerivative"](a, b, fromTipC=tip)
            return None if r is False else r

        def differ(self, a, b):
            return G["areVectorsDifferent"](a, b)

        def root_vector(self, a, bl, tip):
            return G["rootVector"](a, bl, tip, flat, 0)  # a one-node tree without MAT mutations: the list is taken as it is

        def prob_root(self, a):
            return G["findProbRoot"](a)

        def shorten(self, a):
            v = copy.deepcopy(list(a))
            G["shorten"](v)
            return v

    import types
    flat = types.SimpleNamespace(up=[None], mutations=[[]])
    plain = {"up": tree.up, "mutations": [m or [] for m in tree.mutations]}
    lower, upper = fuzz_chain.initial_pools(plain, lambda fam, i: getattr(tree, fam)[i])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = fuzz_chain.run_chain(Ref(), lower, upper, G["lRef"], seed, steps)
    return {"seed": seed, "steps": steps, "results": res}


ENV_KEYS = ["lRef", "rootFreqs", "usingErrorRate", "errorRateSiteSpecific", "useRateVariation",
            "thresholdLogLKoptimizationTopology", "thresholdLogLKconsecutivePlacement", "deeperSearchForLongBranches",
            "BLenThresholdDeeperSearch", "effectivelyNon0BLen", "minBLenSensitivity", "thresholdProb",
            "thresholdDiffForUpdate", "thresholdFoldChangeUpdate", "maxReplacements", "doNotImproveTopology",
            "defaultBLen", "oneMutBLen", "HnZ", "aBayesPlus", "model", "minimumCarryOver", "globalTotRate",
            "thresholdTopologyPlacement", "supportFor0Branches", "doTimeTree"]


class Harvest:
    def __init__(self, name, cap):
        self.name = name
        self.cap = cap
        self.rounds = []
        self.first = None

    def pool_map(self, func, inputs):
        G = func.__globals__
        inputs = list(inputs)
        tree, root = inputs[0][0], inputs[0][1]
        if self.first is None:
            rec = Recorder(G, self.cap)
            env = {k: G[k] for k in ENV_KEYS}
            env["ref"] = G["ref"]
            (_, _, _, strict, fails, thr, thrPlace, mutRate, errorRateGlobal, mutMatrixGlobal, errorRates, mutMatrices,
             cumulativeRate, cumulativeErrorRate) = inputs[0]
            siteRates = G.get("siteRates") if G["useRateVariation"] else None
            # the device side recomputes mutMatrices / cumulative tables from these; check that this is lossless
            if mutMatrices is not None and G["useRateVariation"]:
                for p in range(0, G["lRef"], 997):
                    for i in range(4):
                        for j in range(4):
                            assert mutMatrices[p][i][j] == mutMatrixGlobal[i][j] * siteRates[p]
            cr = [0.0]
            refIdx = G["refIndeces"]
            for i in range(G["lRef"]):
                cr.append(cr[-1] + mutMatrixGlobal[refIdx[i]][refIdx[i]] * (siteRates[i] if siteRates else 1.0))
            if siteRates:
                assert cr == list(cumulativeRate), "cumulativeRate is not recomputable bit-exactly"
            else:
                cr2 = [0.0]
                for i in range(G["lRef"]):
                    cr2.append(cr2[-1] + mutMatrixGlobal[refIdx[i]][refIdx[i]])
                assert cr2 == list(cumulativeRate), "cumulativeRate (no rate variation) is not recomputable"
            if errorRates is not None and G["usingErrorRate"] and G["errorRateSiteSpecific"]:
                ce = [0.0]
                for i in range(G["lRef"]):
                    ce.append(ce[-1] + errorRates[i])
                assert ce == list(cumulativeErrorRate)
            model = {"mutMatrixGlobal": [list(r) for r in mutMatrixGlobal], "errorRateGlobal": errorRateGlobal,
                     "siteRates": list(siteRates) if siteRates else None,
                     "errorRates": list(errorRates) if (errorRates is not None and G["errorRateSiteSpecific"]) else None,
                     "totError": G.get("totError") if G["usingErrorRate"] else None}
            params = {"strict": bool(strict), "fails": fails, "thr": thr, "thrPlace": thrPlace, "numCores": len(inputs)}
            snap = snapshot_tree(tree, root, rec.table)
            treeLK = G["calculateTreeLikelihood"](tree, root)
            # explicit vectors for the functions the searches do not exercise: the per-node merges of
            # calculateTreeLikelihood (returnLK=True, :9756), findProbRoot (:4865) and rootVector (:4916)
            T = rec.table
            nodes, stack = [], [root]
            while stack:
                nd = stack.pop()
                nodes.append(nd)
                stack.extend(tree.children[nd])
            for nd in nodes:
                ch = tree.children[nd]
                if not ch or len(rec.calls["mergeVectors"]) >= 160:
                    continue
                if tree.mutations[ch[0]] or tree.mutations[ch[1]]:
                    continue
                tip = [len(tree.children[c]) == 0 and len(tree.minorSequences[c]) == 0 for c in ch]
                nm = [len(tree.minorSequences[c]) for c in ch]
                out, lk = G["mergeVectors"](tree.probVect[ch[0]], tree.dist[ch[0]], tip[0], tree.probVect[ch[1]],
                                            tree.dist[ch[1]], tip[1], returnLK=True, numMinor1=nm[0], numMinor2=nm[1])
                rec.calls["mergeVectors"].append({"v1": T.add(tree.probVect[ch[0]]), "b1": tree.dist[ch[0]], "t1": bool(tip[0]),
                                                  "v2": T.add(tree.probVect[ch[1]]), "b2": tree.dist[ch[1]], "t2": bool(tip[1]),
                                                  "returnLK": True, "isUpDown": False, "numMinor1": nm[0], "numMinor2": nm[1],
                                                  "out": T.add(out), "lk": lk, "line": 9756})
            for nd in nodes[:60]:
                v = tree.probVect[nd]
                rec.calls["findProbRoot"].append({"v": T.add(v), "out": G["findProbRoot"](v)})
                tip = len(tree.children[nd]) == 0 and len(tree.minorSequences[nd]) == 0
                for bl in (0.0, tree.dist[nd] if tree.dist[nd] else 1e-5):
                    r = G["rootVector"](v, bl, tip, tree, root)
                    rec.calls["rootVector"].append({"v": T.add(v), "bLen": bl, "isFromTip": bool(tip), "out": T.add(r)})
            placements, place_env = harvest_placements(G, tree, root, T)
            self.extras = harvest_extras(G, tree, root, self.name)
            tp = self.extras.pop("_perturbed_tree")
            self.extras["rounds"] = harvest_rounds(G, func, inputs, {"frozen": tree, "perturbed": tp}, root)
            self.extras["placements_more"] = harvest_more_placements(G, {"frozen": tree, "perturbed": tp}, root)
            self.extras["fuzz"] = harvest_fuzz(G, tree, root, seed=sum(map(ord, self.name)))
            print("[golden] %s fuzz chain: %d steps" % (self.name, len(self.extras["fuzz"]["results"])), file=sys.stderr)
            rec.install()
            try:
                results = [func(x) for x in inputs]
            finally:
                rec.uninstall()
            self.first = {"env": env, "model": model, "params": params, "tree": snap, "treeLK": treeLK,
                          "calls": rec.calls, "callCounts": rec.count, "searches": rec.searches,
                          "proposed": [[list(m) for m in r] for r in results], "phase1Total": rec.phase1,
                          "placements": placements, "placeEnv": place_env, "tipInputs": TIPS_OUT,
                          "lists": [jsonable_list(c) for c in rec.table.lists]}
            print("[golden] %s: harvested %d lists, %d searches, %d phase-1 candidates, counts %s" % (
                self.name, len(rec.table.lists), len(rec.searches), rec.phase1, rec.count), file=sys.stderr)
            return [list(r) for r in results]
        return [func(x) for x in inputs]


def run_config(name):
    path, max_seqs, extra, cap = CONFIGS[name]
    tmp_in = "/tmp/golden_%s_input.txt" % name
    truncate_input(path, max_seqs, tmp_in)
    out_prefix = "/tmp/golden_%s_out" % name
    hv = Harvest(name, cap)

    class FakePool:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def map(self, func, inputs):
            return hv.pool_map(func, inputs)

    real_pool = multiprocessing.Pool
    multiprocessing.Pool = FakePool
    argv = sys.argv
    sys.argv = [REF, "--input", tmp_in, "--output", out_prefix, "--overwrite", "--numCores", "2"] + extra
    crashed = None
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(REF, run_name="__main__")
    except SystemExit:  # the reference ends with exit()
        pass
    except Exception as e:  # --model JC dies in the EM after round 1 (reference behaviour)
        crashed = repr(e)
        if hv.first is None:
            import traceback
            traceback.print_exc()
    finally:
        sys.argv = argv
        multiprocessing.Pool = real_pool
    assert hv.first is not None, "reference never reached the parallel SPR round: " + buf.getvalue()[-2000:]
    finalLK = None
    if crashed is None and os.path.isfile(out_prefix + "_LK.txt"):
        with open(out_prefix + "_LK.txt") as f:
            finalLK = float(f.read().split()[0])
    fx = hv.first
    fx["config"] = {"name": name, "input": os.path.basename(path), "max_seqs": max_seqs, "argv": extra,
                    "reference": "MAPLEv0.7.5.4.py", "interpreter": "CPython %d.%d" % sys.version_info[:2],
                    "crashedAfterHarvest": crashed}
    fx["finalLK"] = finalLK
    out = os.path.join(HERE, name + ".json.gz")
    text = json.dumps(fx, separators=(",", ":"))
    same = False
    if os.path.isfile(out):
        with gzip.open(out, "rt") as f:
            same = f.read() == text
    if not same:  # the reference run is deterministic: leave an unchanged fixture (and its gzip timestamp) alone
        with gzip.open(out, "wt", compresslevel=9) as f:
            f.write(text)
    hv.extras["config"] = fx["config"]
    with gzip.open(os.path.join(HERE, "extras", name + ".json.gz"), "wt", compresslevel=9) as f:
        json.dump(hv.extras, f, separators=(",", ":"))
    print("[golden] wrote %s (%.1f kB), treeLK=%r finalLK=%r crashed=%r" % (
        out, os.path.getsize(out) / 1e3, fx["treeLK"], finalLK, crashed), file=sys.stderr)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        run_config(n)
