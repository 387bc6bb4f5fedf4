// Placement of new samples on a frozen tree: findBestParentForNewSample (MAPLEv0.7.5.4.py:7912-8292) with isMinorSequence
// (:5919-6004), default feature set (computePlacementSupportOnly=False, no HnZ, no time tree).  One sample per thread; samples
// are independent on a frozen tree (the reference's own batch seam for this is process_chunk / joblib, :11190-11287).
//
// Phase 1 is a pre-order walk from the root that scores appendProbNode(probVectTotUp[t1], sample, True, oneMutBLen) at every
// visited branch (:8050) with the same kind of stop rule as the SPR search; phase 2 re-optimises the three branch lengths at
// every branch within thresholdLogLKoptimization of the best (:8109-8187).  This first version is the straight-line form
// (it shares every likelihood primitive with the SPR search); the walk only ever reads STORED lists, so it is the next
// candidate for the warp-cooperative subtree scans of search_fsm.cuh.
#pragma once
#include "search.cuh"

namespace maple {

struct PlaceParams {
    int strictStopRules, allowedFails, deeperSearchForLongBranches, onlyFindIdentical;
    double thresholdLogLK, thresholdLogLKoptimization, thresholdLogLKconsecutivePlacement;
    double effectivelyNon0BLen, BLenThresholdDeeperSearch, oneMutBLen;
};

struct PlaceResult {
    int bestNode;
    int status;        // 0 placed; 1 absorbed as a minor sequence of leaf bestNode (:7949, :8002); 2 aborted; 3 scratch exhausted
    int phase1;        // candidate branches scored in the walk (:8033 / :8050)
    int missedMinors;  // leaves strictly less informative than the sample (:7960, :8004)
    double bestScore, bLenTop, bLenBottom, bLenAppend;  // python False in bestBranchLengths is 0.0 here
};

struct PlaceStackE {
    LRef diffs;
    double parentLK;
    int t1, failedPasses;
};

struct PlaceBest {
    LRef diffs;
    double score;
    int t1, pad;
};

// isMinorSequence(probVect1, probVect2, onlyFindIdentical) (:5919-6004)
__device__ __noinline__ int dev_is_minor(int lRef, LRef a, LRef b, bool onlyFindIdentical) {
    Cursor<false> e1, e2;
    e1.init(a.k, a.p);
    e2.init(b.k, b.p);
    int pos = 0;
    bool found1bigger = false, found2bigger = false;
    for (;;) {
        if (e1.type != e2.type) {
            if (onlyFindIdentical) return 0;
            else if (e1.type == T_N) {
                if (e2.type == T_R) pos = min(e1.end, e2.end);
                else pos += 1;
                found2bigger = true;
            } else if (e2.type == T_N) {
                if (e1.type == T_R) pos = min(e1.end, e2.end);
                else pos += 1;
                found1bigger = true;
            } else if (e1.type == T_O) {
                const int i2 = (e2.type == T_R) ? e1.nuc : e2.type;
                if (e1.v(i2) > 0.1) found2bigger = true;
                else return 0;
                pos += 1;
            } else if (e2.type == T_O) {
                const int i1 = (e1.type == T_R) ? e2.nuc : e1.type;
                if (e2.v(i1) > 0.1) found1bigger = true;
                else return 0;
                pos += 1;
            } else return 0;
        } else if (e1.type == T_O) {
            for (int j = 0; j < 4; j++) {
                const double x1 = e1.v(j), x2 = e2.v(j);
                if (onlyFindIdentical) {
                    if (x2 != x1) return 0;
                } else if (x2 > 0.1 && x1 < 0.1) found1bigger = true;
                else if (x1 > 0.1 && x2 < 0.1) found2bigger = true;
            }
            pos += 1;
        } else {
            if (e1.type < 4) pos += 1;
            else pos = min(e1.end, e2.end);
        }
        if (found1bigger && found2bigger) return 0;
        if (pos == lRef) break;
        if (e1.type < 4 || e1.type == T_O || pos == e1.end) e1.next();
        if (e2.type < 4 || e2.type == T_O || pos == e2.end) e2.next();
    }
    if (found1bigger) return found2bigger ? 0 : 1;
    return found2bigger ? 2 : 1;
}

// One entry of the refinement loop (:8109-8187): the three branch lengths of a placement on the branch above `node` are
// re-optimised and the placement is scored again.  Entries are independent of one another (the loop only keeps the last
// best one).  Returns 0, or the status the placement ends with (2: the reference would raise, 3: scratch exhausted).
struct PlaceEval {
    double score, top, bottom, append;
};

__device__ int place_refine_entry(const DevModel& m, const DevTree& t, ScratchD& s, int node, LRef d, PlaceEval& e) {
    const unsigned mk = s.topK, mp = s.topP;
    const LRef upVect = up_list_for(m, t, s, node);
    const bool isTip = t.isTip[node] != 0;
    const LRef pv = tree_list(t, 0, node), tot = tree_list(t, 3, node);
    if (!upVect.k || !pv.k || !tot.k) return s.err ? s.err : 2;
    const double bestAppendingLength = s_blen(m, s, tot, d, true);
    const LRef midLower = s_merge(m, s, pv, t.dist[node] / 2, isTip, d, bestAppendingLength, true, false);
    if (!midLower.k) return s.err ? s.err : 2;
    const double bestTopLength = s_blen(m, s, upVect, midLower, false);
    const LRef midTop = s_merge(m, s, upVect, bestTopLength, false, d, bestAppendingLength, true, true);
    if (!midTop.k) return s.err ? s.err : 2;
    const double bestBottomLength = s_blen(m, s, midTop, pv, isTip);
    const LRef newMid = s_merge(m, s, upVect, bestTopLength, false, pv, bestBottomLength, isTip, true);
    if (!newMid.k) return s.err ? s.err : 2;
    const double appendingCost = f_append(m, newMid, d, true, bestAppendingLength);
    const double initialCost = f_append(m, upVect, pv, isTip, t.dist[node]);
    const double newPartialCost = f_append(m, upVect, pv, isTip, bestBottomLength + bestTopLength);
    e.score = appendingCost + newPartialCost - initialCost;
    e.top = bestTopLength; e.bottom = bestBottomLength; e.append = bestAppendingLength;
    s.topK = mk; s.topP = mp;
    return s.err;
}

__device__ void place_sample(const DevModel& m, const DevTree& t, const PlaceParams& pp, LRef in, ScratchD& s, PlaceStackE* stack, int stackCap,
                             PlaceBest* best, int bestCap, PlaceResult& r) {
    r.bestNode = -1; r.status = 0; r.phase1 = 0; r.missedMinors = 0;
    r.bestScore = r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
    s.topK = s.topP = 0;
    s.err = 0;
    const int root = t.root;
    const double eff = pp.effectivelyNon0BLen, one = pp.oneMutBLen;
    LRef diffs = s_copy(s, in);  // shorten() works in place on it (:8065)
    if (!diffs.k) { r.status = 3; return; }
    if (n_mut(t, root)) diffs = s_pass(m, t, s, diffs, root, false);
    if (!diffs.k) { r.status = 3; return; }
    int bestNode = root, nBest = 0, sp = 0;
    double bTop = 0.0, bBottom = 0.0, bAppend = one;  // (False, False, oneMutBLen)
#define PLACE_FAIL(code) do { r.status = (code); return; } while (0)
    if (t.child0[root] < 0) {
        const int cmp = dev_is_minor(m.lRef, tree_list(t, 0, root), diffs, pp.onlyFindIdentical != 0);
        if (cmp == 1) { r.bestNode = root; r.bestScore = 1.0; PLACE_FAIL(1); }
        else if (cmp == 2) r.missedMinors++;
    }
    const LRef rootVect = s_root_vector(m, t, s, tree_list(t, 0, root), 0.0, false);
    if (!rootVect.k) PLACE_FAIL(s.err ? s.err : 2);
    double bestLKdiff = f_append(m, rootVect, diffs, true, one);
    const double originalLKdiff = bestLKdiff;
    if (t.child0[root] >= 0) {
        for (int i = 0; i < 2; i++) {
            const int c = i == 0 ? t.child0[root] : t.child1[root];
            LRef dc = diffs;
            if (n_mut(t, c)) dc = s_pass(m, t, s, diffs, c, false);
            if (!dc.k || sp >= stackCap) PLACE_FAIL(3);
            stack[sp].t1 = c; stack[sp].parentLK = bestLKdiff; stack[sp].failedPasses = 0; stack[sp].diffs = dc; sp++;
        }
    }
    while (sp > 0) {
        const PlaceStackE E = stack[--sp];
        const int t1 = E.t1;
        int failedPasses = E.failedPasses;
        LRef d = E.diffs;
        double LKdiff;
        if (t.child0[t1] < 0) {  // a leaf: is the new sample identical to / contained in it? (:7975-8005)
            const int cmp = dev_is_minor(m.lRef, tree_list(t, 0, t1), d, pp.onlyFindIdentical != 0);
            if (cmp == 1) { r.bestNode = t1; r.bestScore = 1.0; PLACE_FAIL(1); }
            else if (cmp == 2) r.missedMinors++;
        }
        if (t.dist[t1] > eff && t.up[t1] >= 0) {
            double bestTopLength, bestBottomLength, bestAppendingLength;
            if (pp.deeperSearchForLongBranches && t.dist[t1] > pp.BLenThresholdDeeperSearch) {
                const bool isTip = t.isTip[t1] != 0;
                const LRef pv = tree_list(t, 0, t1);
                const unsigned mk = s.topK, mp = s.topP;
                const LRef upVect = up_list_for(m, t, s, t1);
                bestAppendingLength = one;
                const LRef midLower = s_merge(m, s, pv, t.dist[t1] / 2, isTip, d, bestAppendingLength, true, false);
                if (!midLower.k) PLACE_FAIL(s.err ? s.err : 2);
                bestTopLength = s_blen(m, s, upVect, midLower, false);
                const LRef midTop = s_merge(m, s, upVect, bestTopLength, false, d, bestAppendingLength, true, true);
                if (!midTop.k) PLACE_FAIL(s.err ? s.err : 2);
                bestBottomLength = s_blen(m, s, midTop, pv, isTip);
                const LRef newMid = s_merge(m, s, upVect, bestTopLength, false, pv, bestBottomLength, isTip, true);
                if (!newMid.k) PLACE_FAIL(s.err ? s.err : 2);
                LKdiff = f_append(m, newMid, d, true, bestAppendingLength);
                s.topK = mk; s.topP = mp;
            } else {
                const LRef tot = tree_list(t, 3, t1);
                if (!tot.k) PLACE_FAIL(2);
                LKdiff = f_append(m, tot, d, true, one);
                bestBottomLength = t.dist[t1] / 2;
                bestTopLength = t.dist[t1] / 2;
                bestAppendingLength = one;
            }
            r.phase1++;
            if (LKdiff >= bestLKdiff) {
                f_shorten_inplace(m, d);  // :8065
                bestLKdiff = LKdiff;
                bestNode = t1;
                failedPasses = 0;
                if (nBest >= bestCap) PLACE_FAIL(3);
                best[nBest].t1 = t1; best[nBest].score = LKdiff; best[nBest].diffs = d; nBest++;
                bTop = bestTopLength; bBottom = bestBottomLength / 2; bAppend = bestAppendingLength;
            } else if (LKdiff > bestLKdiff - pp.thresholdLogLKoptimization) {
                if (nBest >= bestCap) PLACE_FAIL(3);
                best[nBest].t1 = t1; best[nBest].score = LKdiff; best[nBest].diffs = d; nBest++;
            }
            if (LKdiff < (E.parentLK - pp.thresholdLogLKconsecutivePlacement)) failedPasses++;
        } else LKdiff = E.parentLK;
        bool go;
        if (pp.strictStopRules) go = failedPasses <= pp.allowedFails && LKdiff > (bestLKdiff - pp.thresholdLogLK);
        else go = failedPasses <= pp.allowedFails || LKdiff > (bestLKdiff - pp.thresholdLogLK);
        if (go && t.child0[t1] >= 0) {
            for (int i = 0; i < 2; i++) {
                const int c = i == 0 ? t.child0[t1] : t.child1[t1];
                LRef dc = d;
                if (n_mut(t, c)) dc = s_pass(m, t, s, d, c, false);
                if (!dc.k || sp >= stackCap) PLACE_FAIL(3);
                stack[sp].t1 = c; stack[sp].parentLK = LKdiff; stack[sp].failedPasses = failedPasses; stack[sp].diffs = dc; sp++;
            }
        }
    }
    // refinement of every branch within thresholdLogLKoptimization of the best (:8109-8187)
    double bestScore = bestLKdiff;
    for (int i = 0; i < nBest; i++) {
        if (!(best[i].score >= bestLKdiff - pp.thresholdLogLKoptimization)) continue;
        PlaceEval e;
        const int rc = place_refine_entry(m, t, s, best[i].t1, best[i].diffs, e);
        if (rc) PLACE_FAIL(rc);
        if (e.score >= bestScore) {
            bestNode = best[i].t1;
            bestScore = e.score;
            bTop = e.top; bBottom = e.bottom; bAppend = e.append;
        }
    }
#undef PLACE_FAIL
    if (bestScore == -INFINITY) bestScore = originalLKdiff;
    r.bestNode = bestNode;
    r.bestScore = bestScore;
    r.bLenTop = bTop; r.bLenBottom = bBottom; r.bLenAppend = bAppend;
    r.status = s.err ? s.err : 0;
}

}  // namespace maple
