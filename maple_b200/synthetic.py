"""Synthetic SARS-CoV-2-like inputs: the reference ships no simulator.

reference genome : `lRef` random bases with composition A .299 / C .183 / G .196 / T .321 (seed)
tree             : random binary tree by joining random pairs of lineages (Kingman topology)
substitutions    : dropped on branches with an UNREST matrix like the one MAPLE estimates
                   (C->T and G->T dominated); branch lengths ~ Exp, scaled so that the mean tip is
                   `mean_diffs` substitutions from the root; Gamma(alpha=0.5) site rates optional
tips             : N-runs at both ends (U[20,300]), with prob .2 an internal N-run (U[50,500]),
                   with prob .1 one IUPAC two-state ambiguity at a random site
Outputs: the tree (numpy arrays, node = int index like the reference's Tree, :331-376), the tip
genome lists as reference-style tuples (what probVectTerminalNode builds, :3882-3962), and
optionally a MAPLE-format alignment file (reader :3498-3553) so the reference can run on it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .model import MapleModel, ref_indices

BASES = "acgt"
# UNREST matrix of the kind MAPLE estimates on SARS-CoV-2 data (rows: from, cols: to), before normalisation
_UNREST_RAW = np.array([[0.0, 0.039, 0.310, 0.123],
                        [0.140, 0.0, 0.022, 3.028],
                        [0.747, 0.113, 0.0, 2.953],
                        [0.056, 0.261, 0.036, 0.0]])
_AMBIG = {frozenset((1, 3)): "y", frozenset((0, 2)): "r", frozenset((0, 3)): "w", frozenset((1, 2)): "s",
          frozenset((2, 3)): "k", frozenset((0, 1)): "m"}


@dataclass
class SyntheticData:
    ref: str
    model: MapleModel
    up: np.ndarray        # int32 [nNodes], -1 at the root
    child0: np.ndarray    # int32 [nNodes], -1 for tips
    child1: np.ndarray
    dist: np.ndarray      # float64 [nNodes]
    root: int
    tip_nodes: np.ndarray  # int32 [nSeq]
    tip_lists: List[list]  # reference-style genome lists, one per tip (aligned with tip_nodes)
    tip_diffs: List[list]  # MAPLE-format diffs per tip: [(char, pos[, len])]


def normalised_unrest(pi: np.ndarray) -> np.ndarray:
    Q = _UNREST_RAW.copy()
    for i in range(4):
        Q[i, i] = -Q[i].sum()
    tot = -float((pi * np.diag(Q)).sum())
    return Q / tot


def make_reference(lRef: int, seed: int = 1) -> str:
    rng = np.random.default_rng(seed)
    idx = rng.choice(4, size=lRef, p=[0.299, 0.183, 0.196, 0.322])
    return "".join(BASES[i] for i in idx)


def make_tree(nSeq: int, seed: int = 2):
    """Random binary tree; tips are nodes 0..nSeq-1, internal nodes follow, root is the last node."""
    rng = np.random.default_rng(seed)
    n = 2 * nSeq - 1
    up = np.full(n, -1, np.int32)
    c0 = np.full(n, -1, np.int32)
    c1 = np.full(n, -1, np.int32)
    active = list(range(nSeq))
    nxt = nSeq
    # random pair joins; swap-remove keeps this O(n)
    r = rng.random(2 * nSeq)
    for step in range(nSeq - 1):
        m = len(active)
        i = int(r[2 * step] * m)
        a = active[i]
        active[i] = active[-1]
        active.pop()
        j = int(r[2 * step + 1] * (m - 1))
        b = active[j]
        active[j] = nxt
        c0[nxt], c1[nxt] = a, b
        up[a] = up[b] = nxt
        nxt += 1
    return up, c0, c1, n - 1


def _depths(up: np.ndarray, root: int, c0, c1) -> np.ndarray:
    depth = np.zeros(len(up), np.int32)
    order = [root]
    for nd in order:
        for c in (c0[nd], c1[nd]):
            if c >= 0:
                depth[c] = depth[nd] + 1
                order.append(c)
    return depth, np.array(order, np.int32)


def generate(nSeq: int, lRef: int = 29903, mean_diffs: float = 10.0, rate_variation: bool = False, error_model: bool = False,
             site_specific_errors: bool = False, seed: int = 1, ml_like_blens: bool = False) -> SyntheticData:
    """ml_like_blens: replace the simulated branch lengths by (substitutions on the branch)/lRef, i.e. zero for
    branches without a substitution -- the shape of the trees MAPLE itself estimates (large multifurcations of
    zero-length branches), instead of the simulator's strictly positive lengths."""
    ref = make_reference(lRef, seed)
    ridx = ref_indices(ref).astype(np.int64)
    pi = np.bincount(ridx, minlength=4) / float(lRef)
    Q = normalised_unrest(pi)
    rng = np.random.default_rng(seed + 2)
    siteRates = None
    if rate_variation:
        siteRates = rng.gamma(0.5, 2.0, lRef)
        siteRates /= siteRates.mean()
        siteRates = np.maximum(siteRates, 1e-3)
    errorRates = None
    eps = 1e-5
    if error_model and site_specific_errors:
        errorRates = np.full(lRef, eps) * rng.uniform(0.5, 1.5, lRef)
    up, c0, c1, root = make_tree(nSeq, seed + 1)
    depth, order = _depths(up, root, c0, c1)
    mean_tip_depth = float(depth[:nSeq].mean())
    mu = mean_diffs / (mean_tip_depth * lRef)  # mean branch length in substitutions per site
    dist = rng.exponential(mu, len(up))
    dist[root] = 0.0
    # substitutions: Poisson number per branch, sites by rate, target base by the UNREST row of the current base
    site_p = None if siteRates is None else siteRates / siteRates.sum()
    nmut = rng.poisson(dist * lRef)
    to_p = [np.where(np.arange(4) == i, 0.0, Q[i]) / -Q[i, i] for i in range(4)]
    genomes: List[Optional[dict]] = [None] * len(up)  # node -> {pos0: base}
    genomes[root] = {}
    for nd in order[1:]:
        g = dict(genomes[up[nd]])
        k = int(nmut[nd])
        if k:
            sites = rng.choice(lRef, size=k, p=site_p)
            for s in sites:
                cur = g.get(int(s), int(ridx[s]))
                new = int(rng.choice(4, p=to_p[cur]))
                if new == ridx[s]:
                    g.pop(int(s), None)
                else:
                    g[int(s)] = new
        genomes[nd] = g
    if ml_like_blens:
        dist = nmut.astype(np.float64) / float(lRef)
        dist[root] = 0.0
    for nd in order:  # internal genomes are no longer needed once their children exist
        if c0[nd] >= 0:
            genomes[nd] = None
    tip_lists, tip_diffs = [], []
    U = bool(error_model)
    for t in range(nSeq):
        g = genomes[t]
        if U:  # sequencing errors at eps per site
            for s in rng.integers(0, lRef, rng.poisson(eps * lRef)):
                g[int(s)] = int((g.get(int(s), int(ridx[s])) + rng.integers(1, 4)) % 4)
                if g[int(s)] == ridx[s]:
                    g.pop(int(s))
        nstart = int(rng.integers(20, 301))
        nend = int(rng.integers(20, 301))
        nruns = [(1, nstart), (lRef - nend + 1, nend)]  # (first position 1-based, length)
        if rng.random() < 0.2:
            ln = int(rng.integers(50, 501))
            st = int(rng.integers(nstart + 2, lRef - nend - ln - 1))
            nruns.append((st, ln))
        amb = None
        if rng.random() < 0.1:
            amb = int(rng.integers(nstart + 1, lRef - nend))  # 1-based position
        events = []  # (pos1, kind, payload)
        for st, ln in nruns:
            events.append((st, "n", ln))
        masked = lambda p1: any(st <= p1 < st + ln for st, ln in nruns)  # noqa: E731
        for s, b in g.items():
            if not masked(s + 1) and s + 1 != amb:
                events.append((s + 1, "b", b))
        if amb is not None and not masked(amb):
            true = g.get(amb - 1, int(ridx[amb - 1]))
            other = int((true + rng.integers(1, 4)) % 4)
            code = _AMBIG[frozenset((true, other))]
            events.append((amb, "o", code))
        events.sort()
        gl, diffs, pos = [], [], 1
        for p1, kind, payload in events:
            if p1 > pos:
                gl.append((4, p1 - 1))
                pos = p1
            if kind == "n":
                gl.append((5, p1 + payload - 1))
                diffs.append(("n", p1, payload))
                pos = p1 + payload
            elif kind == "b":
                gl.append((payload, int(ridx[p1 - 1])))
                diffs.append((BASES[payload], p1))
                pos = p1 + 1
            else:
                vec = [1.0 if BASES[i] in {"y": "ct", "r": "ag", "w": "at", "s": "cg", "k": "gt", "m": "ac"}[payload] else 0.0
                       for i in range(4)]
                if U:  # probVectTerminalNode under the error model (:3922-3931)
                    e = float(errorRates[p1 - 1]) if errorRates is not None else eps
                    vec = [e * 0.33333 if v == 0 else v - e * 0.33333 for v in vec]
                gl.append((6, int(ridx[p1 - 1]), vec))
                diffs.append((payload, p1))
                pos = p1 + 1
        if pos <= lRef:
            gl.append((4, lRef))
        tip_lists.append(gl)
        tip_diffs.append(diffs)
        genomes[t] = None
    model = MapleModel(lRef=lRef, refIdx=ridx.astype(np.int8), rootFreqs=pi, Q=Q, usingErrorRate=U,
                       errorRateSiteSpecific=bool(U and site_specific_errors), useRateVariation=bool(rate_variation),
                       errorRate=eps if U else 1.0 / lRef, siteRates=siteRates, errorRates=errorRates)
    return SyntheticData(ref=ref, model=model, up=up, child0=c0, child1=c1, dist=dist, root=root,
                         tip_nodes=np.arange(nSeq, dtype=np.int32), tip_lists=tip_lists, tip_diffs=tip_diffs)


def write_maple_file(data: SyntheticData, path: str, names: Optional[List[str]] = None):
    """MAPLE-format alignment: '>reference', the reference, then per sample '>name' and 'char<TAB>pos[<TAB>len]'."""
    with open(path, "w") as f:
        f.write(">reference\n%s\n" % data.ref)
        for i, diffs in enumerate(data.tip_diffs):
            f.write(">%s\n" % (names[i] if names else "S%d" % i))
            for d in diffs:
                f.write("\t".join(str(x) for x in d) + "\n")


def newick(data: SyntheticData, names: Optional[List[str]] = None) -> str:
    out = {}
    order = []
    stack = [data.root]
    while stack:
        nd = stack.pop()
        order.append(nd)
        if data.child0[nd] >= 0:
            stack.extend((int(data.child0[nd]), int(data.child1[nd])))
    for nd in reversed(order):
        if data.child0[nd] < 0:
            s = names[nd] if names else "S%d" % nd
        else:
            s = "(%s,%s)" % (out.pop(int(data.child0[nd])), out.pop(int(data.child1[nd])))
        out[nd] = s + (":%r" % float(data.dist[nd]) if nd != data.root else "")
    return out[data.root] + ";"
