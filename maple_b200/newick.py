"""Newick in and out, and the set-up of an input tree from an alignment: the callers on the file side of the search.

Mirrors, for the reference's default feature set (no MAT / support / lineage annotations):
  * readNewick (MAPLEv0.7.5.4.py:1812-1957)   -> read_newick: same node numbering (a node is created at every "(" and ","),
    same naming modes (keepNames / namesInTree list + dict / inputDictNames / onlyTerminalNodeName), same branch-length
    conventions (missing length = defaultBLen, negative length made positive only for a node closed by ","), "[...]" skipped;
  * makeTreeBinary (:2117-2133)                -> make_tree_binary: polytomies resolved from the last two children, new nodes
    appended with length 0;
  * createNewick (:2816-2956)                  -> create_newick: binary (polytomies as zero-length branches, minor sequences as
    a ladder of zero-length cherries named <tip>_MinorSeqsClade) or multifurcating output, lengths printed with repr();
  * the first pass of reCalculateAllGenomeLists(firstSetUp=True) (:6039-6146) -> set_up_input_tree: tip genome lists from the
    alignment, and the collapse of zero-length sibling tips that are minor sequences of one another (isMinorSequence, :5919);
  * the _subs.txt / _LK.txt writers (:12480-12519).

Host code: strings, dictionaries and O(n) list walks, as in the reference.  The likelihood work that follows (all four
genome-list families, the tree likelihood) runs on the device: HostTree.arrays() feeds DeviceTree.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Sequence

import numpy as np

from .alignment import tip_genome_list


class NewickError(ValueError):
    """The reference prints a message and raises Exception("exit")."""


class HostTree:
    """The reference's Tree (:331-376) reduced to what the file side needs; node = index into every list."""

    def __init__(self):
        self.up: List[Optional[int]] = []
        self.children: List[Optional[List[int]]] = []
        self.dist: List[float] = []
        self.name: list = []
        self.minorSequences: List[list] = []
        self.dirty: List[bool] = []
        self.probVect: List[Optional[list]] = []

    def add_node(self, dirtiness: bool = True) -> int:
        self.up.append(None)
        self.children.append([])
        self.dist.append(0.0)
        self.name.append("")
        self.minorSequences.append([])
        self.dirty.append(dirtiness)
        self.probVect.append(None)
        return len(self.up) - 1

    def __len__(self):
        return len(self.up)

    def arrays(self):
        """int32 / float64 arrays in DeviceTree's convention (-1 = none); nodes removed by the minor-sequence collapse
        (children None) stay in the numbering, unreachable from the root, as in the reference."""
        n = len(self)
        up = np.array([-1 if u is None else u for u in self.up], np.int32)
        child0 = np.array([c[0] if c else -1 for c in self.children], np.int32)
        child1 = np.array([c[1] if c else -1 for c in self.children], np.int32)
        dist = np.array([float(d) if d else 0.0 for d in self.dist], np.float64)
        numMinor = np.array([len(m) for m in self.minorSequences], np.int32)
        isTip = np.array([(not self.children[i]) and numMinor[i] == 0 for i in range(n)], np.uint8)
        return {"up": up, "child0": child0, "child1": child1, "dist": dist, "numMinor": numMinor, "isTip": isTip}

    def reachable(self, root: int) -> List[int]:
        out, stack = [], [root]
        while stack:
            nd = stack.pop()
            out.append(nd)
            stack.extend(self.children[nd] or [])
        return out


_TOKEN = re.compile(r"\[[^\]]*\]|[(),;]|:[^,);]*|[^(),;:\[]+")


def read_newick(source: str, default_blen: float, multiple_trees: bool = False, dirtiness: bool = True, create_dict: bool = False,
                input_dict_names: Optional[Dict[str, int]] = None, keep_names: bool = False, only_terminal_node_name: bool = False,
                normalize_input_blen: float = 1.0, is_text: bool = False):
    """readNewick.  `source` is a file name (or the Newick text itself with is_text=True).  Returns what the reference returns:
    keep_names -> [(tree, root)...]; create_dict -> (trees, namesInTree, namesInTreeDict); default -> (trees, namesInTree);
    input_dict_names given -> trees."""
    if is_text:
        lines = source.split("\n")
    else:
        with open(source) as f:
            lines = f.read().split("\n")
    trees = []
    names_in_tree: List[str] = []
    names_dict: Dict[str, int] = {}

    for line in lines:
        if line == "":
            if trees and not multiple_trees:
                break
            continue
        tree = HostTree()
        node = tree.add_node(dirtiness)
        pending_name, pending_dist, internal, finished = "", "", False, False

        def close(nd, after_comma):
            nonlocal pending_name, pending_dist
            if pending_name != "":
                if keep_names:
                    tree.name[nd] = pending_name
                elif not (only_terminal_node_name and internal):
                    if input_dict_names is None:
                        tree.name[nd] = len(names_in_tree)
                        if create_dict:
                            names_dict[pending_name] = len(names_in_tree)
                        names_in_tree.append(pending_name)
                    else:
                        key = pending_name.replace("?", "_").replace("&", "_")
                        if key not in input_dict_names:
                            raise NewickError("sample %s not found in the original tree" % key)
                        tree.name[nd] = input_dict_names[key]
                pending_name = ""
            if pending_dist != "":
                d = float(pending_dist) * normalize_input_blen
                if after_comma and d < 0.0:  # :1882-1884 (a node closed by ")" keeps its sign, :1918)
                    d = abs(d)
                tree.dist[nd] = d
                pending_dist = ""
            else:
                tree.dist[nd] = default_blen

        for tok in _TOKEN.findall(line):
            c = tok[0]
            if c == "(":
                child = tree.add_node(dirtiness)
                tree.children[node].append(child)
                tree.up[child] = node
                node, internal = child, False
            elif c == ",":
                close(node, True)
                parent = tree.up[node]
                child = tree.add_node(dirtiness)
                tree.children[parent].append(child)
                tree.up[child] = parent
                node, internal = child, False
            elif c == ")":
                close(node, False)
                node, internal = tree.up[node], True
            elif c == ":":
                pending_dist += tok[1:]
            elif c == "[":
                pass
            elif c == ";":
                trees.append((tree, node))
                finished = True
                break
            else:
                pending_name += tok
        if not finished:
            raise NewickError("final character ; not found in newick string" + ("" if is_text else " in file " + source))
        if not multiple_trees:
            break
    if keep_names:
        return trees
    if create_dict:
        return trees, names_in_tree, names_dict
    if input_dict_names is None:
        return trees, names_in_tree
    return trees


def make_tree_binary(tree: HostTree, root: int) -> None:
    """makeTreeBinary (:2117): a node with k > 2 children keeps its first child and gets a ladder of k-2 new zero-length nodes."""
    stack = [root]
    while stack:
        node = stack.pop()
        ch = tree.children[node]
        if not ch:
            continue
        while len(ch) > 2:
            c2, c1 = ch.pop(), ch.pop()
            new = tree.add_node()
            tree.up[c1] = tree.up[c2] = new
            tree.children[new] = [c1, c2]
            tree.up[new] = node
            ch.append(new)
        stack.extend(ch[:2])


def create_newick(tree, root: int, binary: bool = True, names_in_tree: Optional[Sequence[str]] = None,
                  include_minor_seqs: bool = True) -> str:
    """createNewick without annotations.  `tree` needs up / children / dist / name / minorSequences (HostTree or anything alike)."""
    up, children, dist, name, minors = tree.up, tree.children, tree.dist, tree.name, tree.minorSequences

    def label(x) -> str:
        if names_in_tree is None:
            return str(x)
        return "" if x == "" else names_in_tree[x]

    def blen(nd) -> str:
        return ":" + repr(float(dist[nd])) if dist[nd] else ":0.0"

    out: List[str] = []
    # explicit stack of (node, stage): stage 0 = entering, 1 = between the two children, 2 = leaving
    stack = [(root, 0)]
    while stack:
        nd, stage = stack.pop()
        ch = children[nd]
        if ch:
            wrapped = bool(dist[nd]) or binary or up[nd] is None
            if stage == 0:
                if wrapped:
                    out.append("(")
                stack.append((nd, 1))
                stack.append((ch[0], 0))
            elif stage == 1:
                out.append(",")
                stack.append((nd, 2))
                stack.append((ch[1], 0))
            elif wrapped:
                out.append(")" + label(name[nd]) + blen(nd))
            continue
        ms = minors[nd]
        if ms and include_minor_seqs:
            me = label(name[nd])
            if binary:  # ((tip:0.0,m1:0.0):0.0,m2:0.0)tip_MinorSeqsClade
                out.append("(" * len(ms) + me + ":")
                for m in ms[:-1]:
                    out.append("0.0," + label(m) + ":0.0):")
                out.append("0.0," + label(ms[-1]) + ":0.0)" + me + "_MinorSeqsClade")
            else:
                wrapped = bool(dist[nd]) or up[nd] is None
                out.append(("(" if wrapped else "") + me + ":0.0" + "".join("," + label(m) + ":0.0" for m in ms))
                if wrapped:
                    out.append(")" + me + "_MinorSeqsClade")
        else:
            out.append(label(name[nd]))
        out.append(blen(nd))
    out.append(";")
    return "".join(out)


# ---------------------------------------------------------------------------------------------------------------------------
def shorten(vec: list, threshold_prob: float) -> None:
    """shorten (:3721-3745), in place: neighbouring R entries with the same details become one."""
    i = 0
    while i < len(vec) - 1:
        a, b = vec[i], vec[i + 1]
        if a[0] == 4 and b[0] == 4 and len(a) == len(b):
            if len(b) == 2:
                vec.pop(i)
            elif abs(b[2] - a[2]) > threshold_prob:
                i += 1
            elif len(b) == 3:
                vec.pop(i)
            elif abs(b[3] - a[3]) > threshold_prob:
                i += 1
            elif len(b) == 4 or b[4] == a[4]:
                vec.pop(i)
            else:
                i += 1
        else:
            i += 1


def is_minor_sequence(v1: list, v2: list, lRef: int, only_find_identical: bool = False) -> int:
    """isMinorSequence (:5919-6003) for two tip lists: 1 = the first is at least as informative (or they are identical),
    2 = the second is strictly more informative, 0 = not comparable."""
    i1 = i2 = pos = 0
    e1, e2 = v1[0], v2[0]
    big1 = big2 = False
    while True:
        t1, t2 = e1[0], e2[0]
        if t1 != t2:
            if only_find_identical:
                return 0
            if t1 == 5 or t2 == 5:
                other = t2 if t1 == 5 else t1
                pos = min(e1[1], e2[1]) if other == 4 else pos + 1
                if t1 == 5:
                    big2 = True
                else:
                    big1 = True
            elif t1 == 6:
                if e1[-1][e1[1] if t2 == 4 else t2] > 0.1:
                    big2 = True
                else:
                    return 0
                pos += 1
            elif t2 == 6:
                if e2[-1][e2[1] if t1 == 4 else t1] > 0.1:
                    big1 = True
                else:
                    return 0
                pos += 1
            else:
                return 0
        elif t1 == 6:
            for j in range(4):
                a, b = e1[-1][j], e2[-1][j]
                if only_find_identical:
                    if a != b:
                        return 0
                elif b > 0.1 and a < 0.1:
                    big1 = True
                elif a > 0.1 and b < 0.1:
                    big2 = True
            pos += 1
        elif t1 < 4:
            pos += 1
        else:
            pos = min(e1[1], e2[1])
        if big1 and big2:
            return 0
        if pos == lRef:
            break
        if t1 < 4 or t1 == 6 or pos == e1[1]:
            i1 += 1
            e1 = v1[i1]
        if t2 < 4 or t2 == 6 or pos == e2[1]:
            i2 += 1
            e2 = v2[i2]
    if big2 and not big1:
        return 2
    return 1


def set_up_input_tree(tree: HostTree, root: int, data: Dict[str, list], names: Sequence[str], model,
                      only_find_identical: Optional[bool] = None, only_n_ambiguities: bool = False) -> int:
    """First pass of reCalculateAllGenomeLists(firstSetUp=True) as far as the tips go (:6039-6132): every leaf gets
    probVectTerminalNode of its alignment record (shortened); a zero-length leaf that is the second child of its parent, next
    to a zero-length leaf, is merged with it when one is a minor sequence of the other (the loser's name goes to the winner's
    minorSequences, the parent node is taken out of the tree); under an error model the ambiguity vectors of tips are then
    rewritten for their final minor-sequence count (updateProbVectTerminalNode, :3966).  The tree is modified in place;
    returns the new root (the reference keeps its root variable: a root is never removed because it has no parent, but the
    parent of a collapsed pair may be the root -- then the surviving tip becomes the root).  `model` is a MapleModel.

    only_find_identical: the reference's rule (:6083) -- identical sequences only when any error / support / HnZ option is on;
    defaults to model.usingErrorRate."""
    lRef, U = model.lRef, bool(model.usingErrorRate)
    if only_find_identical is None:
        only_find_identical = U
    refIdx = model.refIdx
    err_rates = model.errorRates if (U and model.errorRateSiteSpecific) else None
    up, children, dist, minors, probVect = tree.up, tree.children, tree.dist, tree.minorSequences, tree.probVect
    lookup = dict(data)
    converted = False
    removed = 0

    def tip_list(nd):
        nonlocal converted
        nm = names[tree.name[nd]] if tree.name[nd] != "" else ""
        if nm not in lookup and not converted:  # :6048-6056: try again with ? and & replaced in the alignment's names
            for k in list(lookup):
                k2 = k.replace("?", "_").replace("&", "_")
                if k2 != k:
                    lookup[k2] = lookup[k]
            converted = True
        if nm not in lookup:
            raise NewickError("sample name %s not found in the input sequence data - all samples in the input tree need a sequence entry" % nm)
        v = tip_genome_list(lookup[nm], refIdx, lRef, usingErrorRate=U, errorRate=model.errorRate or 0.0, errorRates=err_rates,
                            onlyNambiguities=only_n_ambiguities, numMinSeqs=len(minors[nd]))
        shorten(v, model.thresholdProb)
        return v

    # post-order with the reference's visiting order (child 0 before child 1); leaves are handled when reached
    order, stack = [], [root]
    while stack:
        nd = stack.pop()
        order.append(nd)
        stack.extend(reversed(children[nd]))  # pre-order child 0 first; leaves then come in the reference's order
    for nd in order:
        if children[nd] is None or children[nd]:
            continue
        probVect[nd] = tip_list(nd)
        node = nd
        while up[node] is not None and children[up[node]][1] == node and not dist[node]:
            sib = children[up[node]][0]
            if dist[sib] or children[sib]:
                break
            cmp = is_minor_sequence(probVect[node], probVect[sib], lRef, only_find_identical)
            if cmp == 1:
                major, minor = node, sib
            elif cmp == 2:
                major, minor = sib, node
            else:
                break
            removed += 1
            minors[major].append(tree.name[minor])
            minors[major].extend(minors[minor])
            probVect[minor] = None
            parent = up[major]
            up[major] = up[parent]
            dist[major] = dist[parent]
            if up[major] is not None:
                gp = children[up[major]]
                gp[0 if gp[0] == parent else 1] = major
            elif parent == root:
                root = major
            children[parent] = None
            node = major
        if U and not only_n_ambiguities and minors[node]:
            probVect[node] = tip_list(node)  # numMinSeqs > 0 changes the ambiguity vectors
    tree.numMinorsRemoved = removed
    return root


def load_input_tree(newick: str, alignment: str, model, default_blen: Optional[float] = None, only_terminal_node_name: bool = False, only_find_identical: Optional[bool] = None,
                    only_n_ambiguities: bool = False, normalize_input_blen: float = 1.0):
    """--inputTree + --input as the reference sets them up (:3644-3651, :6431-6441): read the Newick file, make it binary,
    read the MAPLE alignment, build the tip lists and collapse minor sequences.  Returns (tree, root, namesInTree, tip_nodes,
    tip_lists): everything DeviceTree needs to build the four list families on the device."""
    from .alignment import read_maple_alignment
    if default_blen is None:
        default_blen = 1.0 / model.lRef  # the reference's defaultBLen = oneMutBLen unless changed on the command line
    trees, names, _ = read_newick(newick, default_blen, create_dict=True, only_terminal_node_name=only_terminal_node_name,
                                  normalize_input_blen=normalize_input_blen)
    tree, root = trees[0]
    make_tree_binary(tree, root)
    _, data = read_maple_alignment(alignment)
    root = set_up_input_tree(tree, root, data, names, model, only_find_identical, only_n_ambiguities)
    tip_nodes = [i for i in tree.reachable(root) if not tree.children[i]]
    return tree, root, names, tip_nodes, [tree.probVect[i] for i in tip_nodes]


# ---------------------------------------------------------------------------------------------------------------------------
def write_subs(path: str, mutMatrix, siteRates=None, errorRates=None, errorRate=None) -> None:
    """<output>_subs.txt (:12487-12503): the 4x4 rate matrix, then site rates / site error rates / the global error rate when
    those options are on.  Numbers are printed with repr() like the reference's str()."""
    with open(path, "w") as f:
        for i in range(4):
            for j in range(4):
                f.write(repr(float(mutMatrix[i][j])) + "\t")
            f.write("\n")
        if siteRates is not None:
            f.write("\n\nSite rates:\n")
            for i, r in enumerate(siteRates):
                f.write(str(i + 1) + "\t" + repr(float(r)) + "\n")
        if errorRates is not None:
            f.write("\n\nSite error rates:\n")
            for i, r in enumerate(errorRates):
                f.write(str(i + 1) + "\t" + repr(float(r)) + "\n")
        elif errorRate is not None:
            f.write("\n\nError rate: " + repr(float(errorRate)) + "\n")


def write_lk(path: str, total_lk: float) -> None:
    """<output>_LK.txt (:12513-12517)."""
    with open(path, "w") as f:
        f.write(repr(float(total_lk)) + "\n")
