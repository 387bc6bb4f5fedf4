// Device-resident SPR search: the per-node body of startTopologyUpdatesParallel (MAPLEv0.7.5.4.py:9615-9711)
// and findBestParentTopology (:6817-7724) with evaluatePlacement (:6790-6806), one search per thread.
//
// Default feature set of the reference (no time tree, no HnZ, no SPRTA).  Parallelism comes from the 10^5-10^6
// independent searches of a round, not from inside one search: every search keeps the reference's exact visiting
// order (LIFO stack, child 0 pushed first), its running best score and per-path failure counters, so the visited
// set and every comparison are the reference's.  Each thread owns a scratch arena for the lists it derives while
// the passed partials have not yet converged to the stored ones (needsUpdating), managed as a stack that unwinds
// with the DFS.
//
// Phase 2 (:7460-7639) is folded into the walk: an entry is appended to bestNodes only if its score is within
// thresholdLogLKoptimizationTopology of the running best, which never drops below originalLK, so every entry passes
// the phase-2 filter (:7463) and its evaluation depends only on the lists it carries.  Evaluating it on the spot and
// keeping the arg-max in discovery order (ties: later wins, :7635) gives the same result without keeping lists alive.
//
// probVectTotUp of a zero-length child of the root is filled lazily by the reference during the round (:7198-7200),
// which makes its proposals depend on the order of searches inside a worker; here the host fills those (at most two)
// lists before the round (DeviceTree.prepare_search), i.e. the reference's behaviour once they exist.
#pragma once
#include "likelihood.cuh"

namespace maple {

struct DevTree {
    int nNodes, root;
    const int32_t *up, *child0, *child1;  // -1 = none
    const double* dist;
    const uint8_t* isTip;     // no children and no minor sequences
    const int32_t* mutStart;  // [nNodes+1] CSR into mut or nullptr
    const int32_t* mut;       // triples (pos1, upNuc, downNuc)
    const uint32_t* key;      // list id = family*nNodes + node: 0 lower, 1 upRight, 2 upLeft, 3 totUp
    const double* pay;
    const int64_t *keyStart, *payStart;
    const int32_t* nkeys;
    const int32_t* npay;      // payload doubles per list, or nullptr (then lists are read where they lie)
    // derived by maple_tree_bind for the warp-cooperative subtree scans (search_fsm.cuh); order == nullptr disables them
    const int32_t* order;     // [nNodes] nodes in the search's own pre-order: a node, then the subtree of child 1, then of child 0
    const int32_t* pre;       // position of a node in `order` (-1: not reachable from the root)
    const int32_t* size;      // nodes in the subtree of a node (itself included)
    const int32_t* depth;     // edges between the root and a node
    const uint8_t* mutBelow;  // some node strictly below carries MAT mutations
    const struct ScanNode* scan;  // [nNodes] per pre-order position, refreshed by k_scan_prepare before every search launch
    // second form of the subtree scans (scan2.cuh): records and scan-format copies of the probVectTotUp lists, rebuilt by
    // k_scan_build before every search launch; scan2 == nullptr disables it
    const struct ScanRec* scan2;
    const uint4* scanArena;
    const uint32_t* scanOff;      // [nNodes] per pre-order position: offset of the list's copy in scanArena (16-byte units), ~0u = none
};

// Everything a subtree scan needs to know about the node at one pre-order position, in one 32-byte record (two 16-byte loads,
// consecutive lanes read consecutive records).
struct ScanNode {
    int32_t node;
    int32_t parentPos;  // pre-order position of the parent, -1 for the root
    int32_t size;       // nodes in the subtree
    int32_t depth;
    uint32_t keyOff;    // probVectTotUp list: key offset / 4 and payload offset / 2 in the arena (valid with SN_STAGE)
    uint32_t payOff;
    uint32_t cnt;       // 16-byte units: keys | payload << 16
    uint32_t flags;
};
constexpr uint32_t SN_ELIG = 1;    // has a parent and (dist > effectivelyNon0BLen or the parent is the root): gets a score (:6978, :7184)
constexpr uint32_t SN_TOT = 2;     // probVectTotUp exists
constexpr uint32_t SN_PUSHED = 4;  // the parent's upper list towards this node exists, so the walk pushes this node (:7120, :7157)
constexpr uint32_t SN_INNER = 8;   // has children
constexpr uint32_t SN_STAGE = 16;  // list is 16-byte aligned and small enough for the offsets above
constexpr uint32_t SN_LONG = 32;   // dist > effectivelyNon0BLen: the branch a new sample is scored against (:8013), whatever its parent

struct SearchParams {
    int strictTopologyStopRules, allowedFailsTopology, deeperSearchForLongBranches, reserved;
    double thresholdLogLKtopology, thresholdTopologyPlacement, thresholdLogLKoptimizationTopology;
    double thresholdLogLKconsecutivePlacement, effectivelyNon0BLen, BLenThresholdDeeperSearch, defaultBLen;
};

struct SearchResult {
    int placement;  // proposed re-attachment node or -1
    int bestNode;   // findBestParentTopology's bestNode, -1 when the search did not run
    int status;     // 0 ok, 1 search not needed, 2 aborted (reference: try/except -> no proposal), 3 scratch overflow
    int phase1;     // candidate placements scored by the phase-1 appendProbNode calls (:7011 / :7223)
    double improvement, bestCurrentLK, bestScore, bLenTop, bLenBottom, bLenAppend;
};

struct LRef {
    const uint32_t* k;
    const double* p;
    int nk;
};

struct ScratchD {
    uint32_t* key;
    double* pay;
    double* ais;
    unsigned capK, capP, capA, topK, topP;
    int err;  // 0 ok, 3 scratch overflow, 2 abort (the reference would raise: a None list reached a likelihood call)
};

struct StackE {
    LRef passed, removed;
    double distance, lastLK;
    int t1, failedPasses;
    unsigned markK, markP;
    signed char direction, needsUpdating;
};

__device__ __forceinline__ LRef lnull() { return LRef{nullptr, nullptr, 0}; }

__device__ MAPLE_HELPER_INLINE LRef tree_list(const DevTree& t, int fam, int node) {
    const int64_t id = (int64_t)fam * t.nNodes + node;
    const int64_t ks = t.keyStart[id];
    if (ks < 0) return lnull();
    return LRef{t.key + ks, t.pay + t.payStart[id], t.nkeys[id]};
}

__device__ __forceinline__ int n_mut(const DevTree& t, int node) { return t.mutStart ? t.mutStart[node + 1] - t.mutStart[node] : 0; }

__device__ __forceinline__ bool sc_reserve(ScratchD& s, unsigned nk) {
    const unsigned k = (nk + 3u) & ~3u;
    if (s.topK + k > s.capK || s.topP + 6u * k > s.capP) { s.err = 3; return false; }
    return true;
}

__device__ __forceinline__ LRef sc_commit(ScratchD& s, int nk, int np) {
    LRef r{s.key + s.topK, s.pay + s.topP, nk};
    s.topK += (unsigned(nk) + 3u) & ~3u;
    s.topP += (unsigned(np) + 1u) & ~1u;
    return r;
}

// The ScanNode record of pre-order position i for the bound tree and arena (k_scan_prepare: one thread per position).
__device__ __forceinline__ ScanNode make_scan_node(const DevTree& T, double eff, int i) {
    ScanNode r;
    const int node = T.order[i];
    r.node = node;
    r.keyOff = r.payOff = r.cnt = r.flags = 0;
    if (node < 0 || T.pre[node] != i) {  // positions past the reachable nodes
        r.node = -1; r.parentPos = -1; r.size = 1; r.depth = 0;
        return r;
    }
    const int up = T.up[node];
    const int64_t nN = T.nNodes;
    r.parentPos = up >= 0 ? T.pre[up] : -1;
    r.size = T.size[node];
    r.depth = T.depth[node];
    uint32_t fl = 0;
    if (up >= 0 && (T.dist[node] > eff || T.up[up] < 0)) fl |= SN_ELIG;
    if (up >= 0 && T.dist[node] > eff) fl |= SN_LONG;
    if (up >= 0 && T.keyStart[(T.child0[up] == node ? 1 : 2) * nN + up] >= 0) fl |= SN_PUSHED;
    if (T.child0[node] >= 0) fl |= SN_INNER;
    const int64_t id = 3 * nN + node, ks = T.keyStart[id];
    if (ks >= 0) {
        fl |= SN_TOT;
        const int64_t ps = T.payStart[id];
        const uintptr_t ak = reinterpret_cast<uintptr_t>(T.key + ks), ap = reinterpret_cast<uintptr_t>(T.pay + ps);
        if (T.npay && ((ak | ap) & 15) == 0 && (ks >> 2) < (int64_t(1) << 32) && (ps >> 1) < (int64_t(1) << 32)) {
            const int nk4 = (T.nkeys[id] + 3) >> 2, np2 = (T.npay[id] + 1) >> 1;
            if (nk4 < 65536 && np2 < 65536) {
                fl |= SN_STAGE;
                r.keyOff = uint32_t(ks >> 2);
                r.payOff = uint32_t(ps >> 1);
                r.cnt = uint32_t(nk4) | (uint32_t(np2) << 16);
            }
        }
    }
    r.flags = fl;
    return r;
}

// passGenomeListThroughBranch (:3749-3877)
__device__ __noinline__ void dev_pass_branch(int lRef, const uint32_t* k, const double* p, const int32_t* mut, int nMut, bool dirIsUp,
                                             Writer& o) {
    Cursor<false> c;
    c.init(k, p);
    int iM = 0, lastPos = 0;
    for (;;) {
        const double l0 = c.l0(), l1 = c.l1();
        if (c.type == T_N) {
            o.put0(T_N, 0, c.end);
            lastPos = c.end;
            while (iM < nMut && mut[3 * iM] <= lastPos) iM++;
        } else if (c.type < 4) {
            lastPos += 1;
            if (iM < nMut && mut[3 * iM] <= lastPos) {
                const int target = dirIsUp ? mut[3 * iM + 1] : mut[3 * iM + 2];
                iM++;
                if (c.type == target) o.put(T_R, c.nl, c.flag, 0, lastPos, l0, l1, nullptr);
                else o.put(c.type, c.nl, c.flag, target, lastPos, l0, l1, nullptr);
            } else o.put(c.type, c.nl, c.flag, c.nuc, lastPos, l0, l1, nullptr);
        } else if (c.type == T_R) {
            while (iM < nMut && mut[3 * iM] <= c.end) {
                if (mut[3 * iM] > lastPos + 1) {
                    lastPos = mut[3 * iM] - 1;
                    o.put(T_R, c.nl, c.flag, 0, lastPos, l0, l1, nullptr);
                }
                lastPos += 1;
                const int nucToPass = dirIsUp ? mut[3 * iM + 2] : mut[3 * iM + 1];
                const int newEl = dirIsUp ? mut[3 * iM + 1] : mut[3 * iM + 2];
                iM++;
                o.put(nucToPass, c.nl, c.flag, newEl, lastPos, l0, l1, nullptr);
            }
            if (lastPos < c.end) {
                lastPos = c.end;
                o.put(T_R, c.nl, c.flag, 0, lastPos, l0, l1, nullptr);
            }
        } else {
            double v[4];
            c.vec(v);
            lastPos += 1;
            int nuc = c.nuc;
            if (iM < nMut && mut[3 * iM] <= lastPos) {
                nuc = dirIsUp ? mut[3 * iM + 1] : mut[3 * iM + 2];
                iM++;
            }
            o.put(T_O, c.nl, 0, nuc, lastPos, l0, 0.0, v);
        }
        if (lastPos == lRef) break;
        c.next();
    }
}

// out-of-line instances of the primitives: the search calls each of them from many places
__device__ __noinline__ double f_append(const DevModel& m, LRef P, LRef C, bool isTipC, double bLen) {
    return dev_append<false>(m, P.k, P.p, C.k, C.p, isTipC, bLen);
}
__device__ __noinline__ int f_merge(const DevModel& m, LRef a, double b1, bool t1, LRef b, double b2, bool t2, int flags, Writer& w) {
    return dev_merge<false>(m, a.k, a.p, b1, t1, b.k, b.p, b2, t2, flags, 0, 0, w, nullptr);
}
__device__ __noinline__ double f_blen(const DevModel& m, LRef P, LRef C, bool fromTipC, double* ais) {
    double out = 0.0;
    dev_blen<false>(m, P.k, P.p, C.k, C.p, fromTipC, ais, &out);  // python False and 0.0 both mean "zero length" to the callers
    return out;
}
__device__ __noinline__ bool f_differ(const DevModel& m, LRef a, LRef b) { return dev_differ<false>(m, a.k, a.p, b.k, b.p); }
__device__ __noinline__ void f_shorten_inplace(const DevModel& m, LRef& v) {
    Writer w;
    w.init(const_cast<uint32_t*>(v.k), const_cast<double*>(v.p));
    dev_shorten<false>(m, v.k, v.p, w);
    v.nk = w.nk;
}

__device__ LRef s_pass(const DevModel& m, const DevTree& t, ScratchD& s, LRef v, int node, bool dirIsUp) {
    const int nm = n_mut(t, node);
    if (!v.k || !sc_reserve(s, unsigned(v.nk) + 2u * unsigned(nm) + 2u)) return lnull();
    Writer w;
    w.init(s.key + s.topK, s.pay + s.topP);
    dev_pass_branch(m.lRef, v.k, v.p, t.mut + 3 * (size_t)t.mutStart[node], nm, dirIsUp, w);
    return sc_commit(s, w.nk, w.np);
}

__device__ LRef s_merge(const DevModel& m, ScratchD& s, LRef a, double b1, bool t1, LRef b, double b2, bool t2, bool upDown) {
    if (!a.k || !b.k) { if (!s.err) s.err = 2; return lnull(); }
    if (!sc_reserve(s, unsigned(a.nk) + unsigned(b.nk))) return lnull();
    Writer w;
    w.init(s.key + s.topK, s.pay + s.topP);
    if (f_merge(m, a, b1, t1, b, b2, t2, upDown ? 1 : 0, w) != 0) return lnull();
    return sc_commit(s, w.nk, w.np);
}

// rootVector(probVect, bLen, isFromTip, tree, node) for node == root (:4916-4996), shortened (:4994)
__device__ LRef s_root_vector(const DevModel& m, const DevTree& t, ScratchD& s, LRef v, double bLen, bool isFromTip) {
    const int root = t.root;
    if (n_mut(t, root)) v = s_pass(m, t, s, v, root, true);
    if (!v.k || !sc_reserve(s, unsigned(v.nk))) return lnull();
    Writer w;
    w.init(s.key + s.topK, s.pay + s.topP);
    dev_root_vector<false>(m, v.k, v.p, bLen, isFromTip, w);
    LRef r = sc_commit(s, w.nk, w.np);
    if (n_mut(t, root)) r = s_pass(m, t, s, r, root, false);
    if (r.k) f_shorten_inplace(m, r);
    return r;
}

__device__ LRef s_copy(ScratchD& s, LRef v) {
    if (!v.k || !sc_reserve(s, unsigned(v.nk))) return lnull();
    uint32_t* dk = s.key + s.topK;
    double* dp = s.pay + s.topP;
    int np = 0;
    for (int i = 0; i < v.nk; i++) {
        const uint32_t k = v.k[i];
        dk[i] = k;
        np += int((k >> 3) & 3u) + (((k & 7u) == 6u) ? 4 : 0);
    }
    for (int i = 0; i < np; i++) dp[i] = v.p[i];
    return sc_commit(s, v.nk, np);
}

__device__ double s_blen(const DevModel& m, ScratchD& s, LRef P, LRef C, bool fromTipC) {
    if (!P.k || !C.k) { if (!s.err) s.err = 2; return 0.0; }
    if (unsigned(P.nk) + unsigned(C.nk) + 1u > s.capA) { s.err = 3; return 0.0; }
    return f_blen(m, P, C, fromTipC, s.ais);
}

// evaluatePlacement (:6790-6806); true when the reference would raise (a None list reaches the next call) or on overflow
__device__ bool eval_placement(const DevModel& m, const SearchParams& sp, ScratchD& s, LRef midTot, LRef downVect, LRef upVect,
                               double distance, LRef removed, bool isRemovedTip, bool fromTip1, double& cost, double& bBottom,
                               double& bTop, double& bAppend) {
    if (!midTot.k || !downVect.k || !upVect.k) return true;
    const unsigned mk = s.topK, mp = s.topP;
    const double bestAppending = s_blen(m, s, midTot, removed, isRemovedTip);
    LRef midLower = s_merge(m, s, downVect, distance / 2, fromTip1, removed, bestAppending, isRemovedTip, false);
    if (!midLower.k) return true;
    double bestTop = s_blen(m, s, upVect, midLower, false);
    LRef midTop = s_merge(m, s, upVect, bestTop, false, removed, bestAppending, isRemovedTip, true);
    if (!midTop.k) {
        if (s.err) return true;
        bestTop = sp.defaultBLen * 0.1;
        midTop = s_merge(m, s, upVect, bestTop, false, removed, bestAppending, isRemovedTip, true);
        if (!midTop.k) return true;
    }
    const double bestBottom = s_blen(m, s, midTop, downVect, fromTip1);
    LRef newMid = s_merge(m, s, upVect, bestTop, false, downVect, bestBottom, fromTip1, true);
    if (!newMid.k) return true;
    cost = f_append(m, newMid, removed, isRemovedTip, bestAppending);
    bBottom = bestBottom;
    bTop = bestTop;
    bAppend = bestAppending;
    s.topK = mk;
    s.topP = mp;
    return s.err != 0;
}

struct Phase2 {
    double bestScore, bTop, bBottom, bAppend;
    int bestNode;
};

// one bestNodes entry, evaluated on the spot (:7460-7639 without HnZ / time / SPRTA)
__device__ bool phase2_entry(const DevModel& m, const SearchParams& sp, ScratchD& s, int t1, LRef midTot, LRef downVect, LRef upVect,
                             double distance, LRef removed, bool isRemovedTip, bool fromTip1, Phase2& ph) {
    double cost, bB, bT, bA;
    if (eval_placement(m, sp, s, midTot, downVect, upVect, distance, removed, isRemovedTip, fromTip1, cost, bB, bT, bA)) return true;
    const double initialCost = f_append(m, upVect, downVect, fromTip1, distance);
    const double newPartialCost = f_append(m, upVect, downVect, fromTip1, bB + bT);
    const double optimizedScore = cost + newPartialCost - initialCost;
    if (optimizedScore >= ph.bestScore) {
        ph.bestNode = t1;
        ph.bestScore = optimizedScore;
        ph.bTop = bT;
        ph.bBottom = bB;
        ph.bAppend = bA;
    }
    return false;
}

__device__ __forceinline__ LRef up_list_for(const DevModel& m, const DevTree& t, ScratchD& s, int t1) {
    // the upper list seen by t1 from its parent, in t1's local reference (:7001-7006, :7467-7472)
    LRef v = (t1 == t.child0[t.up[t1]]) ? tree_list(t, 1, t.up[t1]) : tree_list(t, 2, t.up[t1]);
    if (n_mut(t, t1)) v = s_pass(m, t, s, v, t1, false);
    return v;
}

__device__ int find_best_parent_topology(const DevModel& m, const DevTree& t, const SearchParams& sp, ScratchD& s, StackE* stack, int stackCap,
                                         int node, int child, double bestLKdiff, double removedBLen, Phase2& ph, int& phase1) {
    const int32_t* up = t.up;
    const double* dist = t.dist;
    const double eff = sp.effectivelyNon0BLen;
#define CH(n, i) ((i) == 0 ? t.child0[n] : t.child1[n])
#define PUSH(T1, DIR, NU, PASSED, DISTANCE, LASTLK, FAILS, REMOVED)                                   \
    do {                                                                                               \
        if (s.err) return s.err;                                                                        \
        if (spN >= stackCap) return 3;                                                             \
        StackE& e_ = stack[spN++];                                                                     \
        e_.t1 = (T1); e_.direction = (signed char)(DIR); e_.needsUpdating = (signed char)(NU);          \
        e_.passed = (PASSED); e_.distance = (DISTANCE); e_.lastLK = (LASTLK); e_.failedPasses = (FAILS); \
        e_.removed = (REMOVED); e_.markK = s.topK; e_.markP = s.topP;                                  \
    } while (0)
    int spN = 0;
    const int pruned = CH(node, child), sibling = CH(node, 1 - child);
    // the removed list is copied to scratch: the reference may shorten it in place (:7087)
    LRef removedRel = s_copy(s, tree_list(t, 0, pruned));
    if (n_mut(t, pruned)) removedRel = s_pass(m, t, s, removedRel, pruned, true);
    LRef bestRemoved = removedRel;
    if (n_mut(t, sibling)) bestRemoved = s_pass(m, t, s, bestRemoved, sibling, false);
    if (!removedRel.k) return s.err ? s.err : 2;
    const bool isRemovedTip = t.isTip[pruned] != 0;
    ph.bestNode = sibling;
    ph.bestScore = bestLKdiff;  // originalLK
    if (up[node] >= 0) {
        int childUp;
        LRef vectUpUp;
        if (t.child0[up[node]] == node) { childUp = 1; vectUpUp = tree_list(t, 1, up[node]); }
        else { childUp = 2; vectUpUp = tree_list(t, 2, up[node]); }
        LRef probVect1 = tree_list(t, 0, sibling);
        if (n_mut(t, sibling)) probVect1 = s_pass(m, t, s, probVect1, sibling, true);
        LRef removedRel1 = removedRel;
        if (n_mut(t, node)) {
            probVect1 = s_pass(m, t, s, probVect1, node, true);
            removedRel1 = s_pass(m, t, s, removedRel, node, true);
        }
        PUSH(up[node], childUp, 1, probVect1, dist[sibling] + dist[node], bestLKdiff, 0, removedRel1);
        if (n_mut(t, node)) vectUpUp = s_pass(m, t, s, vectUpUp, node, false);
        removedRel1 = removedRel;
        if (n_mut(t, sibling)) {
            vectUpUp = s_pass(m, t, s, vectUpUp, sibling, false);
            removedRel1 = s_pass(m, t, s, removedRel, sibling, false);
        }
        PUSH(sibling, 0, 1, vectUpUp, dist[sibling] + dist[node], bestLKdiff, 0, removedRel1);
        ph.bTop = dist[node]; ph.bBottom = dist[sibling]; ph.bAppend = removedBLen;
    } else {
        if (t.child0[sibling] >= 0) {
            const int c1 = t.child0[sibling], c2 = t.child1[sibling];
            for (int which = 0; which < 2; which++) {
                const int target = which == 0 ? c1 : c2, other = which == 0 ? c2 : c1;
                LRef vectUp1 = tree_list(t, 0, other);
                if (n_mut(t, other)) vectUp1 = s_pass(m, t, s, vectUp1, other, true);
                vectUp1 = s_root_vector(m, t, s, vectUp1, dist[other], t.isTip[other] != 0);
                LRef removedRel1 = bestRemoved;
                if (n_mut(t, target)) {
                    removedRel1 = s_pass(m, t, s, bestRemoved, target, false);
                    vectUp1 = s_pass(m, t, s, vectUp1, target, false);
                }
                PUSH(target, 0, 1, vectUp1, dist[target], bestLKdiff, 0, removedRel1);
            }
        }
        ph.bTop = 0.0; ph.bBottom = dist[sibling]; ph.bAppend = removedBLen;
    }
    if (s.err) return s.err;

    while (spN > 0) {
        const StackE E = stack[--spN];
        s.topK = E.markK;
        s.topP = E.markP;
        const int t1 = E.t1, direction = E.direction;
        bool needsUpdating = E.needsUpdating != 0;
        int failedPasses = E.failedPasses;
        const LRef passed = E.passed;
        LRef removed = E.removed;
        double distance = E.distance, midProb;
        const double lastLK = E.lastLK;
        if (needsUpdating && !passed.k) return s.err ? s.err : 2;
        if (direction == 0) {
            if (!(up[t1] == node || up[t1] < 0) && (dist[t1] > eff || up[up[t1]] < 0)) {
                LRef midTot;
                if (needsUpdating) {
                    midTot = s_merge(m, s, passed, distance / 2, false, tree_list(t, 0, t1), distance / 2, t.isTip[t1] != 0, true);
                    if (s.err) return s.err;
                    if (!midTot.k) continue;
                    if (!f_differ(m, midTot, tree_list(t, 3, t1))) needsUpdating = false;
                } else {
                    midTot = tree_list(t, 3, t1);
                    distance = dist[t1];
                }
                if (!midTot.k) continue;
                if (sp.deeperSearchForLongBranches && distance > sp.BLenThresholdDeeperSearch) {
                    LRef vectUp = up_list_for(m, t, s, t1);
                    double bB, bT, bA;
                    if (eval_placement(m, sp, s, midTot, tree_list(t, 0, t1), vectUp, distance, removed, isRemovedTip, t.isTip[t1] != 0, midProb,
                                       bB, bT, bA))
                        return s.err ? s.err : 2;
                } else {
                    midProb = f_append(m, midTot, removed, isRemovedTip, removedBLen);
                    phase1++;
                }
                if (midProb > bestLKdiff - sp.thresholdLogLKoptimizationTopology) {  // :7071
                    LRef upV, mt;
                    double dd;
                    if (needsUpdating) { upV = passed; dd = distance; mt = midTot; }
                    else { upV = up_list_for(m, t, s, t1); dd = dist[t1]; mt = tree_list(t, 3, t1); }
                    if (phase2_entry(m, sp, s, t1, mt, tree_list(t, 0, t1), upV, dd, removed, isRemovedTip, t.isTip[t1] != 0, ph))
                        return s.err ? s.err : 2;
                }
                if (midProb > bestLKdiff) {
                    bestLKdiff = midProb;
                    failedPasses = 0;
                    f_shorten_inplace(m, removed);  // :7087
                } else if (midProb < (lastLK - sp.thresholdLogLKconsecutivePlacement)) failedPasses++;
            } else midProb = lastLK;

            bool traverse = false;
            if (sp.strictTopologyStopRules) {
                if (failedPasses <= sp.allowedFailsTopology && midProb > (bestLKdiff - sp.thresholdLogLKtopology) && t.child0[t1] >= 0) traverse = true;
            } else if (failedPasses <= sp.allowedFailsTopology || midProb > (bestLKdiff - sp.thresholdLogLKtopology)) {
                if (t.child0[t1] >= 0) traverse = true;
            }
            if (traverse) {
                for (int which = 0; which < 2; which++) {  // child 0 is pushed first, so child 1 is explored first
                    const int c1 = CH(t1, which), otherChild = CH(t1, 1 - which);
                    LRef vUp;
                    if (needsUpdating) {
                        LRef otherPV = tree_list(t, 0, otherChild);
                        if (n_mut(t, otherChild)) otherPV = s_pass(m, t, s, otherPV, otherChild, true);
                        vUp = s_merge(m, s, passed, distance, false, otherPV, dist[otherChild], t.isTip[otherChild] != 0, true);
                        if (s.err) return s.err;
                    } else vUp = which == 0 ? tree_list(t, 1, t1) : tree_list(t, 2, t1);
                    if (vUp.k) {
                        LRef removed1 = removed;
                        if (n_mut(t, c1)) removed1 = s_pass(m, t, s, removed, c1, false);
                        if (needsUpdating && n_mut(t, c1)) vUp = s_pass(m, t, s, vUp, c1, false);
                        PUSH(c1, 0, needsUpdating ? 1 : 0, needsUpdating ? vUp : lnull(), dist[c1], midProb, failedPasses, removed1);
                    }
                }
            }
        } else {  // crawling up from child to parent (:7179-7429)
            const int otherChild = CH(t1, 2 - direction);
            LRef midBottom = lnull(), vectUp = lnull();
            if (up[t1] >= 0 && (dist[t1] > eff || up[up[t1]] < 0)) {
                LRef midTot;
                if (needsUpdating) {
                    LRef otherPV = tree_list(t, 0, otherChild);
                    if (n_mut(t, otherChild)) otherPV = s_pass(m, t, s, otherPV, otherChild, true);
                    midBottom = s_merge(m, s, passed, distance, false, otherPV, dist[otherChild], t.isTip[otherChild] != 0, false);
                    if (s.err) return s.err;
                    if (!midBottom.k) continue;
                    vectUp = up_list_for(m, t, s, t1);
                    midTot = s_merge(m, s, vectUp, dist[t1] / 2, false, midBottom, dist[t1] / 2, false, true);
                    if (s.err) return s.err;
                    if (!midTot.k) continue;
                    if (!f_differ(m, midTot, tree_list(t, 3, t1))) needsUpdating = false;
                } else midTot = tree_list(t, 3, t1);
                if (!midTot.k) continue;
                if (sp.deeperSearchForLongBranches && dist[t1] > sp.BLenThresholdDeeperSearch) {
                    if (!needsUpdating) {
                        midBottom = tree_list(t, 0, t1);
                        vectUp = up_list_for(m, t, s, t1);
                    }
                    double bB, bT, bA;
                    if (eval_placement(m, sp, s, midTot, midBottom, vectUp, dist[t1], removed, isRemovedTip, false, midProb, bB, bT, bA))
                        return s.err ? s.err : 2;
                } else {
                    midProb = f_append(m, midTot, removed, isRemovedTip, removedBLen);
                    phase1++;
                }
                if (midProb >= (bestLKdiff - sp.thresholdLogLKoptimizationTopology)) {  // :7293
                    LRef upV, downV, mt;
                    if (needsUpdating) { upV = vectUp; downV = midBottom; mt = midTot; }
                    else { upV = up_list_for(m, t, s, t1); downV = tree_list(t, 0, t1); mt = tree_list(t, 3, t1); }
                    if (phase2_entry(m, sp, s, t1, mt, downV, upV, dist[t1], removed, isRemovedTip, t.isTip[t1] != 0, ph)) return s.err ? s.err : 2;
                }
                if (midProb > bestLKdiff) { bestLKdiff = midProb; failedPasses = 0; }
                else if (midProb < (lastLK - sp.thresholdLogLKconsecutivePlacement)) failedPasses++;
            } else midProb = lastLK;

            bool keep = false;
            if (sp.strictTopologyStopRules) {
                if (failedPasses <= sp.allowedFailsTopology && midProb > (bestLKdiff - sp.thresholdLogLKtopology)) keep = true;
            } else if (failedPasses <= sp.allowedFailsTopology || midProb > (bestLKdiff - sp.thresholdLogLKtopology)) keep = true;
            if (keep) {
                if (up[t1] >= 0) {
                    const int upChild = (t1 == t.child0[up[t1]]) ? 0 : 1;
                    LRef vUp;
                    if (needsUpdating) {
                        LRef vectUpUp = up_list_for(m, t, s, t1);
                        vUp = s_merge(m, s, vectUpUp, dist[t1], false, passed, distance, false, true);
                        if (s.err) return s.err;
                    } else vUp = direction == 1 ? tree_list(t, 2, t1) : tree_list(t, 1, t1);
                    if (!vUp.k) continue;
                    {
                        LRef removed1 = removed;
                        if (n_mut(t, otherChild)) removed1 = s_pass(m, t, s, removed, otherChild, false);
                        if (needsUpdating && n_mut(t, otherChild)) vUp = s_pass(m, t, s, vUp, otherChild, false);
                        PUSH(otherChild, 0, needsUpdating ? 1 : 0, needsUpdating ? vUp : lnull(), dist[otherChild], midProb, failedPasses, removed1);
                    }
                    if (needsUpdating && !midBottom.k) {
                        LRef otherPV = tree_list(t, 0, otherChild);
                        if (n_mut(t, otherChild)) otherPV = s_pass(m, t, s, otherPV, otherChild, true);
                        midBottom = s_merge(m, s, passed, distance, false, otherPV, dist[otherChild], t.isTip[otherChild] != 0, false);
                        if (s.err) return s.err;
                        if (!midBottom.k) continue;
                    }
                    {
                        LRef removed1 = removed;
                        if (n_mut(t, t1)) removed1 = s_pass(m, t, s, removed, t1, true);
                        if (needsUpdating && n_mut(t, t1)) midBottom = s_pass(m, t, s, midBottom, t1, true);
                        PUSH(up[t1], upChild + 1, needsUpdating ? 1 : 0, needsUpdating ? midBottom : lnull(), dist[t1], midProb, failedPasses, removed1);
                    }
                } else {  // t1 is the root (:7406-7429)
                    LRef vUp = lnull();
                    if (needsUpdating) {
                        vUp = s_root_vector(m, t, s, passed, distance, false);
                        if (n_mut(t, otherChild)) vUp = s_pass(m, t, s, vUp, otherChild, false);
                    }
                    LRef removed1 = removed;
                    if (n_mut(t, otherChild)) removed1 = s_pass(m, t, s, removed, otherChild, false);
                    PUSH(otherChild, 0, needsUpdating ? 1 : 0, vUp, dist[otherChild], midProb, failedPasses, removed1);
                }
            }
        }
    }
    return 0;
#undef CH
#undef PUSH
}

// the per-node body of startTopologyUpdatesParallel (:9619-9711)
__device__ void search_node(const DevModel& m, const DevTree& t, const SearchParams& sp, int node, ScratchD& s, StackE* stack, int stackCap,
                            SearchResult& r) {
    r.placement = -1;
    r.bestNode = -1;
    r.status = 1;
    r.phase1 = 0;
    r.improvement = r.bestCurrentLK = r.bestScore = r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
    if (t.up[node] < 0) return;
    s.topK = s.topP = 0;
    s.err = 0;
    const int parent = t.up[node];
    const int child = (t.child0[parent] == node) ? 0 : 1;
    LRef vectUp = child == 0 ? tree_list(t, 1, parent) : tree_list(t, 2, parent);
    if (n_mut(t, node)) vectUp = s_pass(m, t, s, vectUp, node, false);
    const LRef own = tree_list(t, 0, node);
    if (!vectUp.k || !own.k) { r.status = s.err ? s.err : 2; return; }
    const double bestCurrenBLen = t.dist[node];
    const double bestCurrentLK = f_append(m, vectUp, own, t.isTip[node] != 0, bestCurrenBLen);
    r.bestCurrentLK = bestCurrentLK;
    if (!(bestCurrentLK < sp.thresholdTopologyPlacement || t.dist[node] != 0.0)) return;
    Phase2 ph;
    ph.bestNode = -1;
    ph.bestScore = ph.bTop = ph.bBottom = ph.bAppend = 0.0;
    int phase1 = 0;
    s.topK = s.topP = 0;
    const int rc = find_best_parent_topology(m, t, sp, s, stack, stackCap, parent, child, bestCurrentLK, bestCurrenBLen, ph, phase1);
    r.phase1 = phase1;
    r.status = rc;
    if (rc != 0) return;
    r.bestNode = ph.bestNode;
    r.bestScore = ph.bestScore;
    r.bLenTop = ph.bTop;
    r.bLenBottom = ph.bBottom;
    r.bLenAppend = ph.bAppend;
    if (ph.bestScore + sp.thresholdTopologyPlacement > bestCurrentLK) {  // :9681-9702
        bool updated = true;
        int topNode = t.up[node];
        if (ph.bestNode == topNode) updated = false;
        while (t.dist[topNode] == 0.0 && t.up[topNode] >= 0) topNode = t.up[topNode];
        if (ph.bestNode == topNode && ph.bBottom == 0.0) updated = false;
        const int sibling = child == 0 ? t.child1[parent] : t.child0[parent];
        if (ph.bestNode == sibling) updated = false;
        if (t.up[ph.bestNode] == sibling && ph.bTop == 0.0) updated = false;
        if (updated) {
            r.improvement = ph.bestScore - bestCurrentLK;
            r.placement = ph.bestNode;
        }
    }
}

}  // namespace maple
