// Warp-converged form of the device SPR search (same algorithm and results as search.cuh).
//
// search.cuh runs one search per thread as straight-line code; on the GPU the 32 lanes of a warp are then almost
// always inside DIFFERENT functions (one in appendProbNode, one in mergeVectors, one popping its stack...), so a warp
// issues for one lane at a time (measured: ~1.3e6 SM cycles per candidate placement).  Here every lane still owns
// one search, but the search is a resumable state machine: a lane runs its (cheap, divergent) control code until it
// needs one of the four heavy co-walks, parks the request, and the warp then executes each kind of co-walk ONCE for
// all lanes that asked for it -- the lanes of a warp are converged inside appendProbNode / mergeVectors /
// estimateBranchLength / areVectorsDifferent, which is where the time goes.
//
// The control flow is search.cuh's find_best_parent_topology written as a coroutine: `pc` is the resume point, all
// state that lives across a co-walk sits in the Fsm struct, and YIELD_* are the suspension points.
#pragma once
#include "search.cuh"
#include "scan2.cuh"

namespace maple {

constexpr int kNumSearchStats = 40;

enum FsmOp { OP_NONE = 0, OP_APPEND = 1, OP_MERGE = 2, OP_BLEN = 3, OP_DIFFER = 4, OP_DONE = 5, OP_SCAN = 6, OP_EVALQ = 7 };

struct Fsm {
    int pc, op;
    // request / reply registers of the co-walks
    LRef a1, a2;
    double ab1, ab2, resD;
    LRef resL;
    int at1, at2, aflags, resB;
    // search level
    int parent, child, pruned, sibling, isRemovedTip, phase1, spN, rc;
    double removedBLen, bestLKdiff;
    Phase2 ph;
    // current stack entry
    int t1, direction, needsUpdating, failedPasses, otherChild, which;
    LRef passed, removed, midTot, midBottom, vectUp, vUp;
    double distance, lastLK, midProb;
    // evaluatePlacement / phase-2 sub-machine
    LRef eMidTot, eDown, eUp, eRemoved, midLower, midTop, newMid;
    double eDist, bestAppending, bestTop, bestBottom, cost, initialCost;
    int eFromTip1, eT1, evalRet;
    unsigned mk, mp;
    // subtree scan (warp_scan_job): phase-2 entries it found, waiting to be evaluated
    int qN, qi, scanNewBest;
    unsigned qRes;
    int evalqDone;  // the queued entries were evaluated by the whole warp (warp_eval_queue)
};

struct PathE {  // per-depth state of a subtree scan: what a node hands to its children (lastLK, failedPasses)
    double lk;
    int failed, pad;
};

#define FSM_FAIL(code)          \
    do {                        \
        f.rc = (code);          \
        f.op = OP_DONE;         \
        return;                 \
    } while (0)

#define FSM_CHECK_ERR()                  \
    do {                                 \
        if (s.err) FSM_FAIL(s.err);      \
    } while (0)

// co-walk requests.  LABEL must be a unique positive integer literal.
#define YIELD_APPEND(LABEL, P_, C_, TIP_, BL_)                                   \
    do {                                                                         \
        f.a1 = (P_); f.a2 = (C_); f.at1 = (TIP_) ? 1 : 0; f.ab1 = (BL_);         \
        if (!f.a1.k || !f.a2.k) FSM_FAIL(2);                                     \
        f.op = OP_APPEND; f.pc = LABEL; return;                                  \
        case LABEL:;                                                             \
    } while (0)

#define YIELD_MERGE(LABEL, A_, B1_, T1_, B_, B2_, T2_, UPDOWN_)                                              \
    do {                                                                                                     \
        f.a1 = (A_); f.ab1 = (B1_); f.at1 = (T1_) ? 1 : 0; f.a2 = (B_); f.ab2 = (B2_); f.at2 = (T2_) ? 1 : 0; \
        f.aflags = (UPDOWN_) ? 1 : 0;                                                                        \
        if (!f.a1.k || !f.a2.k) { if (!s.err) s.err = 2; FSM_FAIL(s.err); }                                  \
        if (!sc_reserve(s, unsigned(f.a1.nk) + unsigned(f.a2.nk))) FSM_FAIL(3);                              \
        f.op = OP_MERGE; f.pc = LABEL; return;                                                               \
        case LABEL:;                                                                                         \
    } while (0)

#define YIELD_BLEN(LABEL, P_, C_, TIP_)                                                      \
    do {                                                                                     \
        f.a1 = (P_); f.a2 = (C_); f.at1 = (TIP_) ? 1 : 0;                                    \
        if (!f.a1.k || !f.a2.k) FSM_FAIL(2);                                                 \
        if (unsigned(f.a1.nk) + unsigned(f.a2.nk) + 1u > s.capA) FSM_FAIL(3);                \
        f.op = OP_BLEN; f.pc = LABEL; return;                                                \
        case LABEL:;                                                                         \
    } while (0)

#define YIELD_DIFFER(LABEL, A_, B_)                         \
    do {                                                    \
        f.a1 = (A_); f.a2 = (B_);                           \
        if (!f.a1.k) FSM_FAIL(2);                           \
        f.op = OP_DIFFER; f.pc = LABEL; return;             \
        case LABEL:;                                        \
    } while (0)

#define FSM_PUSH(T1, DIR, NU, PASSED, DISTANCE, LASTLK, FAILS, REMOVED)                                 \
    do {                                                                                               \
        FSM_CHECK_ERR();                                                                               \
        if (f.spN >= stackCap) FSM_FAIL(3);                                                            \
        StackE& e_ = stack[f.spN++];                                                                   \
        e_.t1 = (T1); e_.direction = (signed char)(DIR); e_.needsUpdating = (signed char)(NU);          \
        e_.passed = (PASSED); e_.distance = (DISTANCE); e_.lastLK = (LASTLK); e_.failedPasses = (FAILS); \
        e_.removed = (REMOVED); e_.markK = s.topK; e_.markP = s.topP;                                  \
    } while (0)

// Runs the search of f until it needs a co-walk (f.op says which) or finishes (f.op == OP_DONE, f.rc = status).
__device__ void fsm_step(Fsm& f, const DevModel& m, const DevTree& t, const SearchParams& sp, ScratchD& s, StackE* stack, int stackCap,
                         int scanMinSize) {
    const int32_t* up = t.up;
    const double* dist = t.dist;
    const double eff = sp.effectivelyNon0BLen;
    f.op = OP_NONE;
#define CH(n, i) ((i) == 0 ? t.child0[n] : t.child1[n])
    switch (f.pc) {
        case 0: {
            f.spN = 0;
            f.pruned = CH(f.parent, f.child);
            f.sibling = CH(f.parent, 1 - f.child);
            f.removed = s_copy(s, tree_list(t, 0, f.pruned));  // search-level "removedRel"; copied: may be shortened in place (:7087)
            if (n_mut(t, f.pruned)) f.removed = s_pass(m, t, s, f.removed, f.pruned, true);
            f.vUp = f.removed;  // bestRemoved
            if (n_mut(t, f.sibling)) f.vUp = s_pass(m, t, s, f.vUp, f.sibling, false);
            if (!f.removed.k) FSM_FAIL(s.err ? s.err : 2);
            f.isRemovedTip = t.isTip[f.pruned] != 0;
            f.ph.bestNode = f.sibling;
            f.ph.bestScore = f.bestLKdiff;  // originalLK
            if (up[f.parent] >= 0) {
                const int node = f.parent, sib = f.sibling;
                int childUp;
                LRef vectUpUp;
                if (t.child0[up[node]] == node) { childUp = 1; vectUpUp = tree_list(t, 1, up[node]); }
                else { childUp = 2; vectUpUp = tree_list(t, 2, up[node]); }
                LRef probVect1 = tree_list(t, 0, sib);
                if (n_mut(t, sib)) probVect1 = s_pass(m, t, s, probVect1, sib, true);
                LRef removedRel1 = f.removed;
                if (n_mut(t, node)) {
                    probVect1 = s_pass(m, t, s, probVect1, node, true);
                    removedRel1 = s_pass(m, t, s, f.removed, node, true);
                }
                FSM_PUSH(up[node], childUp, 1, probVect1, dist[sib] + dist[node], f.bestLKdiff, 0, removedRel1);
                if (n_mut(t, node)) vectUpUp = s_pass(m, t, s, vectUpUp, node, false);
                removedRel1 = f.removed;
                if (n_mut(t, sib)) {
                    vectUpUp = s_pass(m, t, s, vectUpUp, sib, false);
                    removedRel1 = s_pass(m, t, s, f.removed, sib, false);
                }
                FSM_PUSH(sib, 0, 1, vectUpUp, dist[sib] + dist[node], f.bestLKdiff, 0, removedRel1);
                f.ph.bTop = dist[node]; f.ph.bBottom = dist[sib]; f.ph.bAppend = f.removedBLen;
            } else {
                const int sib = f.sibling;
                if (t.child0[sib] >= 0) {
                    const int c1 = t.child0[sib], c2 = t.child1[sib];
                    for (int which = 0; which < 2; which++) {
                        const int target = which == 0 ? c1 : c2, other = which == 0 ? c2 : c1;
                        LRef vectUp1 = tree_list(t, 0, other);
                        if (n_mut(t, other)) vectUp1 = s_pass(m, t, s, vectUp1, other, true);
                        vectUp1 = s_root_vector(m, t, s, vectUp1, dist[other], t.isTip[other] != 0);
                        LRef removedRel1 = f.vUp;
                        if (n_mut(t, target)) {
                            removedRel1 = s_pass(m, t, s, f.vUp, target, false);
                            vectUp1 = s_pass(m, t, s, vectUp1, target, false);
                        }
                        FSM_PUSH(target, 0, 1, vectUp1, dist[target], f.bestLKdiff, 0, removedRel1);
                    }
                }
                f.ph.bTop = 0.0; f.ph.bBottom = dist[sib]; f.ph.bAppend = f.removedBLen;
            }
            FSM_CHECK_ERR();
        }
        // fall through into the main loop
        while (f.spN > 0) {
            {
                const StackE& E = stack[--f.spN];
                s.topK = E.markK;
                s.topP = E.markP;
                f.t1 = E.t1; f.direction = E.direction; f.needsUpdating = E.needsUpdating; f.failedPasses = E.failedPasses;
                f.passed = E.passed; f.removed = E.removed; f.distance = E.distance; f.lastLK = E.lastLK;
            }
            if (f.needsUpdating && !f.passed.k) FSM_FAIL(s.err ? s.err : 2);
            if (f.direction == 0 && !f.needsUpdating && scanMinSize > 0 && t.size[f.t1] >= scanMinSize && !t.mutBelow[f.t1] &&
                !sp.deeperSearchForLongBranches) {
                // The passed partials have converged to the stored ones: everything below t1 is scored against stored
                // lists only, so the whole warp walks this subtree together (warp_scan_job) and hands back the new
                // running best, the number of candidates scored and the phase-2 entries it met, in discovery order.
                f.op = OP_SCAN;
                f.pc = 30;
                return;
                case 30:;
                FSM_CHECK_ERR();
                if (f.scanNewBest) f_shorten_inplace(m, f.removed);  // :7087
                f.qRes = (unsigned(f.qN) + 3u) & ~3u;
                s.capK -= f.qRes;  // the queue sits at the top end of the key scratch
                // The entries are independent of one another (stored lists + the removed list in, a score and three lengths out) and
                // only folded into the best placement in order: the whole warp takes them one per lane (warp_eval_queue); what it
                // cannot do (no per-lane scratch in this launch, or an entry that does not fit its slice) is done here, in turn.
                f.evalqDone = 0;
                if (f.qN >= 2) {
                    f.op = OP_EVALQ;
                    f.pc = 31;
                    return;
                    case 31:;
                }
                for (f.qi = f.evalqDone ? f.qN : 0; f.qi < f.qN; f.qi++) {
                    f.t1 = int(ld_cg(s.key + (s.capK + f.qRes - 1u - unsigned(f.qi))));  // may have been written by a serving warp on another SM
                    f.eUp = up_list_for(m, t, s, f.t1); f.eDist = dist[f.t1]; f.eMidTot = tree_list(t, 3, f.t1);
                    f.eDown = tree_list(t, 0, f.t1);
                    f.eFromTip1 = t.isTip[f.t1] != 0; f.evalRet = 5;
                    goto L_EVAL;
                L_EVAL_RET5:;
                }
                s.capK += f.qRes;
                continue;
            }
            if (f.direction == 0) {
                if (!(up[f.t1] == f.parent || up[f.t1] < 0) && (dist[f.t1] > eff || up[up[f.t1]] < 0)) {
                    if (f.needsUpdating) {
                        YIELD_MERGE(1, f.passed, f.distance / 2, false, tree_list(t, 0, f.t1), f.distance / 2, t.isTip[f.t1] != 0, true);
                        f.midTot = f.resL;
                        if (!f.midTot.k) continue;
                        f.a2 = tree_list(t, 3, f.t1);
                        if (f.a2.k) {
                            YIELD_DIFFER(2, f.midTot, f.a2);
                            if (!f.resB) f.needsUpdating = 0;
                        }
                    } else {
                        f.midTot = tree_list(t, 3, f.t1);
                        f.distance = dist[f.t1];
                    }
                    if (!f.midTot.k) continue;
                    if (sp.deeperSearchForLongBranches && f.distance > sp.BLenThresholdDeeperSearch) {
                        f.eMidTot = f.midTot; f.eDown = tree_list(t, 0, f.t1); f.eUp = up_list_for(m, t, s, f.t1); f.eDist = f.distance;
                        f.eFromTip1 = t.isTip[f.t1] != 0; f.evalRet = 1;
                        goto L_EVAL;
                    L_EVAL_RET1:
                        f.midProb = f.cost;
                    } else {
                        YIELD_APPEND(3, f.midTot, f.removed, f.isRemovedTip, f.removedBLen);
                        f.midProb = f.resD;
                        f.phase1++;
                    }
                    if (f.midProb > f.bestLKdiff - sp.thresholdLogLKoptimizationTopology) {  // :7071
                        if (f.needsUpdating) { f.eUp = f.passed; f.eDist = f.distance; f.eMidTot = f.midTot; }
                        else { f.eUp = up_list_for(m, t, s, f.t1); f.eDist = dist[f.t1]; f.eMidTot = tree_list(t, 3, f.t1); }
                        f.eDown = tree_list(t, 0, f.t1);
                        f.eFromTip1 = t.isTip[f.t1] != 0; f.evalRet = 2;
                        goto L_EVAL;
                    L_EVAL_RET2:;
                    }
                    if (f.midProb > f.bestLKdiff) {
                        f.bestLKdiff = f.midProb;
                        f.failedPasses = 0;
                        f_shorten_inplace(m, f.removed);  // :7087
                    } else if (f.midProb < (f.lastLK - sp.thresholdLogLKconsecutivePlacement)) f.failedPasses++;
                } else f.midProb = f.lastLK;

                {
                    bool traverse = false;
                    if (sp.strictTopologyStopRules) {
                        if (f.failedPasses <= sp.allowedFailsTopology && f.midProb > (f.bestLKdiff - sp.thresholdLogLKtopology) && t.child0[f.t1] >= 0)
                            traverse = true;
                    } else if (f.failedPasses <= sp.allowedFailsTopology || f.midProb > (f.bestLKdiff - sp.thresholdLogLKtopology)) {
                        if (t.child0[f.t1] >= 0) traverse = true;
                    }
                    if (!traverse) continue;
                }
                for (f.which = 0; f.which < 2; f.which++) {  // child 0 is pushed first, so child 1 is explored first
                    f.otherChild = CH(f.t1, 1 - f.which);
                    if (f.needsUpdating) {
                        f.a2 = tree_list(t, 0, f.otherChild);
                        if (n_mut(t, f.otherChild)) f.a2 = s_pass(m, t, s, f.a2, f.otherChild, true);
                        YIELD_MERGE(4, f.passed, f.distance, false, f.a2, dist[f.otherChild], t.isTip[f.otherChild] != 0, true);
                        f.vUp = f.resL;
                    } else f.vUp = f.which == 0 ? tree_list(t, 1, f.t1) : tree_list(t, 2, f.t1);
                    if (f.vUp.k) {
                        const int c1 = CH(f.t1, f.which);
                        LRef removed1 = f.removed;
                        if (n_mut(t, c1)) removed1 = s_pass(m, t, s, f.removed, c1, false);
                        if (f.needsUpdating && n_mut(t, c1)) f.vUp = s_pass(m, t, s, f.vUp, c1, false);
                        FSM_PUSH(c1, 0, f.needsUpdating, f.needsUpdating ? f.vUp : lnull(), dist[c1], f.midProb, f.failedPasses, removed1);
                    }
                }
            } else {  // crawling up from child to parent (:7179-7429)
                f.otherChild = CH(f.t1, 2 - f.direction);
                f.midBottom = lnull();
                f.vectUp = lnull();
                if (up[f.t1] >= 0 && (dist[f.t1] > eff || up[up[f.t1]] < 0)) {
                    if (f.needsUpdating) {
                        f.a2 = tree_list(t, 0, f.otherChild);
                        if (n_mut(t, f.otherChild)) f.a2 = s_pass(m, t, s, f.a2, f.otherChild, true);
                        YIELD_MERGE(5, f.passed, f.distance, false, f.a2, dist[f.otherChild], t.isTip[f.otherChild] != 0, false);
                        f.midBottom = f.resL;
                        if (!f.midBottom.k) continue;
                        f.vectUp = up_list_for(m, t, s, f.t1);
                        YIELD_MERGE(6, f.vectUp, dist[f.t1] / 2, false, f.midBottom, dist[f.t1] / 2, false, true);
                        f.midTot = f.resL;
                        if (!f.midTot.k) continue;
                        f.a2 = tree_list(t, 3, f.t1);
                        if (f.a2.k) {
                            YIELD_DIFFER(7, f.midTot, f.a2);
                            if (!f.resB) f.needsUpdating = 0;
                        }
                    } else f.midTot = tree_list(t, 3, f.t1);
                    if (!f.midTot.k) continue;
                    if (sp.deeperSearchForLongBranches && dist[f.t1] > sp.BLenThresholdDeeperSearch) {
                        if (!f.needsUpdating) {
                            f.midBottom = tree_list(t, 0, f.t1);
                            f.vectUp = up_list_for(m, t, s, f.t1);
                        }
                        f.eMidTot = f.midTot; f.eDown = f.midBottom; f.eUp = f.vectUp; f.eDist = dist[f.t1]; f.eFromTip1 = 0; f.evalRet = 3;
                        goto L_EVAL;
                    L_EVAL_RET3:
                        f.midProb = f.cost;
                    } else {
                        YIELD_APPEND(8, f.midTot, f.removed, f.isRemovedTip, f.removedBLen);
                        f.midProb = f.resD;
                        f.phase1++;
                    }
                    if (f.midProb >= (f.bestLKdiff - sp.thresholdLogLKoptimizationTopology)) {  // :7293
                        if (f.needsUpdating) { f.eUp = f.vectUp; f.eDown = f.midBottom; f.eMidTot = f.midTot; }
                        else { f.eUp = up_list_for(m, t, s, f.t1); f.eDown = tree_list(t, 0, f.t1); f.eMidTot = tree_list(t, 3, f.t1); }
                        f.eDist = dist[f.t1];
                        f.eFromTip1 = t.isTip[f.t1] != 0; f.evalRet = 4;
                        goto L_EVAL;
                    L_EVAL_RET4:;
                    }
                    if (f.midProb > f.bestLKdiff) { f.bestLKdiff = f.midProb; f.failedPasses = 0; }
                    else if (f.midProb < (f.lastLK - sp.thresholdLogLKconsecutivePlacement)) f.failedPasses++;
                } else f.midProb = f.lastLK;

                {
                    bool keep = false;
                    if (sp.strictTopologyStopRules) {
                        if (f.failedPasses <= sp.allowedFailsTopology && f.midProb > (f.bestLKdiff - sp.thresholdLogLKtopology)) keep = true;
                    } else if (f.failedPasses <= sp.allowedFailsTopology || f.midProb > (f.bestLKdiff - sp.thresholdLogLKtopology)) keep = true;
                    if (!keep) continue;
                }
                if (up[f.t1] >= 0) {
                    if (f.needsUpdating) {
                        f.a1 = up_list_for(m, t, s, f.t1);
                        YIELD_MERGE(9, f.a1, dist[f.t1], false, f.passed, f.distance, false, true);
                        f.vUp = f.resL;
                    } else f.vUp = f.direction == 1 ? tree_list(t, 2, f.t1) : tree_list(t, 1, f.t1);
                    if (!f.vUp.k) continue;
                    {
                        LRef removed1 = f.removed;
                        if (n_mut(t, f.otherChild)) removed1 = s_pass(m, t, s, f.removed, f.otherChild, false);
                        if (f.needsUpdating && n_mut(t, f.otherChild)) f.vUp = s_pass(m, t, s, f.vUp, f.otherChild, false);
                        FSM_PUSH(f.otherChild, 0, f.needsUpdating, f.needsUpdating ? f.vUp : lnull(), dist[f.otherChild], f.midProb, f.failedPasses,
                                 removed1);
                    }
                    if (f.needsUpdating && !f.midBottom.k) {
                        f.a2 = tree_list(t, 0, f.otherChild);
                        if (n_mut(t, f.otherChild)) f.a2 = s_pass(m, t, s, f.a2, f.otherChild, true);
                        YIELD_MERGE(10, f.passed, f.distance, false, f.a2, dist[f.otherChild], t.isTip[f.otherChild] != 0, false);
                        f.midBottom = f.resL;
                        if (!f.midBottom.k) continue;
                    }
                    {
                        const int upChild = (f.t1 == t.child0[up[f.t1]]) ? 0 : 1;
                        LRef removed1 = f.removed;
                        if (n_mut(t, f.t1)) removed1 = s_pass(m, t, s, f.removed, f.t1, true);
                        if (f.needsUpdating && n_mut(t, f.t1)) f.midBottom = s_pass(m, t, s, f.midBottom, f.t1, true);
                        FSM_PUSH(up[f.t1], upChild + 1, f.needsUpdating, f.needsUpdating ? f.midBottom : lnull(), dist[f.t1], f.midProb,
                                 f.failedPasses, removed1);
                    }
                } else {  // t1 is the root (:7406-7429)
                    f.vUp = lnull();
                    if (f.needsUpdating) {
                        f.vUp = s_root_vector(m, t, s, f.passed, f.distance, false);
                        if (n_mut(t, f.otherChild)) f.vUp = s_pass(m, t, s, f.vUp, f.otherChild, false);
                    }
                    LRef removed1 = f.removed;
                    if (n_mut(t, f.otherChild)) removed1 = s_pass(m, t, s, f.removed, f.otherChild, false);
                    FSM_PUSH(f.otherChild, 0, f.needsUpdating, f.vUp, dist[f.otherChild], f.midProb, f.failedPasses, removed1);
                }
            }
            continue;

            // ---- evaluatePlacement (:6790-6806) [+ the rest of a phase-2 entry (:7510-7512, :7635-7639) for evalRet 2 and 4]
        L_EVAL:
            if (!f.eMidTot.k || !f.eDown.k || !f.eUp.k) FSM_FAIL(s.err ? s.err : 2);
            f.mk = s.topK;
            f.mp = s.topP;
            YIELD_BLEN(20, f.eMidTot, f.removed, f.isRemovedTip);
            f.bestAppending = f.resD;
            YIELD_MERGE(21, f.eDown, f.eDist / 2, f.eFromTip1, f.removed, f.bestAppending, f.isRemovedTip, false);
            f.midLower = f.resL;
            if (!f.midLower.k) FSM_FAIL(s.err ? s.err : 2);
            YIELD_BLEN(22, f.eUp, f.midLower, false);
            f.bestTop = f.resD;
            YIELD_MERGE(23, f.eUp, f.bestTop, false, f.removed, f.bestAppending, f.isRemovedTip, true);
            f.midTop = f.resL;
            if (!f.midTop.k) {
                FSM_CHECK_ERR();
                f.bestTop = sp.defaultBLen * 0.1;
                YIELD_MERGE(24, f.eUp, f.bestTop, false, f.removed, f.bestAppending, f.isRemovedTip, true);
                f.midTop = f.resL;
                if (!f.midTop.k) FSM_FAIL(s.err ? s.err : 2);
            }
            YIELD_BLEN(25, f.midTop, f.eDown, f.eFromTip1);
            f.bestBottom = f.resD;
            YIELD_MERGE(26, f.eUp, f.bestTop, false, f.eDown, f.bestBottom, f.eFromTip1, true);
            f.newMid = f.resL;
            if (!f.newMid.k) FSM_FAIL(s.err ? s.err : 2);
            YIELD_APPEND(27, f.newMid, f.removed, f.isRemovedTip, f.bestAppending);
            f.cost = f.resD;
            s.topK = f.mk;
            s.topP = f.mp;
            FSM_CHECK_ERR();
            if (f.evalRet == 1) goto L_EVAL_RET1;
            if (f.evalRet == 3) goto L_EVAL_RET3;
            YIELD_APPEND(28, f.eUp, f.eDown, f.eFromTip1, f.eDist);
            f.initialCost = f.resD;
            YIELD_APPEND(29, f.eUp, f.eDown, f.eFromTip1, f.bestBottom + f.bestTop);
            {
                const double optimizedScore = f.cost + f.resD - f.initialCost;
                if (optimizedScore >= f.ph.bestScore) {
                    f.ph.bestNode = f.t1;
                    f.ph.bestScore = optimizedScore;
                    f.ph.bTop = f.bestTop;
                    f.ph.bBottom = f.bestBottom;
                    f.ph.bAppend = f.bestAppending;
                }
            }
            if (f.evalRet == 2) goto L_EVAL_RET2;
            if (f.evalRet == 5) goto L_EVAL_RET5;
            goto L_EVAL_RET4;
        }
            f.rc = 0;
            f.op = OP_DONE;
            return;
        default:
            f.rc = 2;
            f.op = OP_DONE;
            return;
    }
#undef CH
}


// ---- warp-cooperative subtree scan -----------------------------------------------------------------------------
// Lane `src` popped a stack entry (t1, direction 0, needsUpdating False) whose subtree carries no MAT mutations.  From
// there on the reference's walk (:6975-7170) only reads STORED lists: at every node below, the candidate score is
// appendProbNode(probVectTotUp[node], removedPartials, ...) and whether a node is scored at all was decided when its
// parent was processed.  In the search's pre-order (DevTree::order) the subtree is a contiguous range, a pruned
// subtree is a contiguous sub-range, and what a node inherits from its parent (lastLK, failedPasses) is a per-depth
// value.  So the warp takes a window of the range holding up to 32 nodes that need a score, copies their mid-branch
// lists (and once per job the removed list) into shared memory, scores them one per lane -- all lanes inside
// appendProbNode at once, walking shared memory -- then replays the reference's sequential bookkeeping over the window
// in order: running best, failure counters, stop rule, jumping over the ranges the stop rule prunes.  Scores of nodes
// that turn out to be pruned are discarded; nothing the reference would not have scored is counted or kept.
constexpr int kWinMax = 96;      // nodes per window
constexpr int kPathSm = 40;      // per-depth states kept in shared memory (deeper ones go to global scratch)
constexpr int kCKeys = 64;       // removed list staged in shared memory when it has at most this many entries ...
constexpr int kCPay = 96;        // ... and this many payload doubles
struct ScanSmem {
    double winScore[kWinMax];  // candidate scores of the nodes that need one
    double cPay[kCPay];
    PathE path[kPathSm];
    int winInfo[kWinMax], winSize[kWinMax], winNode[kWinMax], winParent[kWinMax];
    int wp[kWinMax], wv[kWinMax], wd[kWinMax];  // pointer-jumping work arrays of the parallel replay
    uint32_t cKey[kCKeys];
    uint4 pool[1];  // mid-branch lists of one batch: poolBytes of them, the rest of the warp's share of the dynamic shared memory
};
// winInfo: bit0 eligible, bit1 has probVectTotUp, bit2 was pushed, bit3 has children, bits 8.. depth relative to the job's root

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += x;
    }
    return v;
}

__device__ __noinline__ double f_append_sitewise(const DevModel& m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC,
                                                 bool isTipC, double bLen) {
    return dev_append_sitewise<false>(m, kP, pP, kC, pC, isTipC, bLen);
}
__device__ __noinline__ double f_append_q4(const DevModel& m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC,
                                           bool isTipC, double bLen) {
    return dev_append_q4<false>(m, kP, pP, kC, pC, isTipC, bLen);
}

__device__ void warp_scan_job(int src, Fsm& f, const DevModel& m, const DevTree& t, const SearchParams& sp, ScratchD& s, StackE* stack,
                              int stackCap, ScanSmem& W, int poolBytes, int scanFlags /* 1: queued-site append, 2: node-by-node replay */,
                              unsigned long long* st) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int R = __shfl_sync(FULL, f.t1, src);
    const int prunedParent = __shfl_sync(FULL, f.parent, src);
    double best = __shfl_sync(FULL, f.bestLKdiff, src);
    const double lastLK0 = __shfl_sync(FULL, f.lastLK, src);
    const int failed0 = __shfl_sync(FULL, f.failedPasses, src);
    const double removedBLen = __shfl_sync(FULL, f.removedBLen, src);
    const bool isRemovedTip = __shfl_sync(FULL, f.isRemovedTip, src) != 0;
    const uint32_t* remK = reinterpret_cast<const uint32_t*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(f.removed.k), src));
    const double* remP = reinterpret_cast<const double*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(f.removed.p), src));
    // deep per-depth state lives in the unused part of lane src's DFS stack; the phase-2 queue grows down from the top of its key scratch
    PathE* gpath = reinterpret_cast<PathE*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(stack + f.spN), src));
    const int pathCap = int((size_t)(stackCap - __shfl_sync(FULL, f.spN, src)) * sizeof(StackE) / sizeof(PathE));
    uint32_t* qTop = reinterpret_cast<uint32_t*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(s.key + s.capK), src));
    const int qCap = int(__shfl_sync(FULL, s.capK - s.topK, src)) - 8;
    const int64_t nN = t.nNodes;
    int phase1 = 0, qN = 0, newBest = 0, err = 0;
    int pos = t.pre[R];
    const int end = pos + t.size[R], d0 = t.depth[R];
    long long tk = st ? clock64() : 0;
    if (st && lane == 0) { st[17] += 1; st[18] += (unsigned long long)(end - pos); }
    __syncwarp();  // the previous job's shared-memory reads are over
    // ---- the removed list: the same for every candidate of the job
    {
        int nk = 0;  // entries up to the one that ends at lRef
        for (int base = 0; base < kCKeys && nk == 0; base += 32) {
            const uint32_t k = remK[base + lane];  // reading past the end stays inside the owner's scratch / the arena slack
            const unsigned hit = __ballot_sync(FULL, int(k >> 8) == m.lRef);
            if (hit) nk = base + __ffs(hit);
        }
        int np = kCPay + 1;
        if (nk) {
            int mine = 0;
            for (int i = lane; i < nk; i += 32) {
                const uint32_t k = remK[i];
                mine += int((k >> 3) & 3u) + (((k & 7u) == 6u) ? 4 : 0);
            }
            np = __shfl_sync(FULL, warp_incl_scan(mine, lane), 31);
        }
        if (nk && np <= kCPay) {
            for (int i = lane; i < nk; i += 32) W.cKey[i] = remK[i];
            for (int i = lane; i < np; i += 32) W.cPay[i] = remP[i];
            remK = W.cKey;
            remP = W.cPay;
        }
    }
    if (pathCap < 2) err = 3;
    else if (lane == 0) W.path[0] = PathE{lastLK0, failed0, 0};
    __syncwarp();
    const int prunedParentPos = t.pre[prunedParent];
    while (pos < end && !err) {
        // ---- window: nodes pos .. pos+nWin-1, at most 32 of them need a score
        int nWin = 0, nScore = 0;
        int myW = -1;  // window position this lane scores, and where its list lies
        uint32_t myKeyOff = 0, myPayOff = 0, myCnt = 0;
        bool myStage = false;
        for (int sweep = 0; sweep < kWinMax / 32 && pos + nWin < end && nScore < 32; sweep++) {
            const int w = nWin + lane, idx = pos + w;
            int info = 0;
            ScanNode rec;
            rec.keyOff = rec.payOff = rec.cnt = rec.flags = 0;
            if (idx < end) {
                const uint4* src4 = reinterpret_cast<const uint4*>(t.scan + idx);
                const uint4 a4 = __ldg(src4), b4 = __ldg(src4 + 1);
                rec.node = int(a4.x); rec.parentPos = int(a4.y); rec.size = int(a4.z); rec.depth = int(a4.w);
                rec.keyOff = b4.x; rec.payOff = b4.y; rec.cnt = b4.z; rec.flags = b4.w;
                const bool elig = (rec.flags & SN_ELIG) && rec.parentPos != prunedParentPos;
                const bool pushed = (rec.flags & SN_PUSHED) || rec.node == R;
                info = (elig ? 1 : 0) | ((rec.flags & SN_TOT) ? 2 : 0) | (pushed ? 4 : 0) | ((rec.flags & SN_INNER) ? 8 : 0) | ((rec.depth - d0) << 8);
                W.winInfo[w] = info;
                W.winSize[w] = rec.size;
                W.winNode[w] = rec.node;
                W.winParent[w] = rec.parentPos - pos;
            }
            const unsigned need = __ballot_sync(FULL, (info & 7) == 7);
            const int room = 32 - nScore;
            int take = min(32, end - pos - nWin);  // nodes of this sweep that join the window
            if (__popc(need) > room) take = __fns(need, 0, room) + 1;  // cut right after the node that fills the last lane
            const unsigned mine = need & (take >= 32 ? FULL : ((1u << take) - 1u));
            // the k-th node that needs a score goes to lane nScore + k
            const int k = lane - nScore;
            const bool gets = k >= 0 && k < __popc(mine);
            const int from = gets ? int(__fns(mine, 0, k + 1)) : lane;
            const uint32_t ko = __shfl_sync(FULL, rec.keyOff, from), po = __shfl_sync(FULL, rec.payOff, from), cn = __shfl_sync(FULL, rec.cnt, from),
                           fl = __shfl_sync(FULL, rec.flags, from);
            if (gets) { myW = nWin + from; myKeyOff = ko; myPayOff = po; myCnt = cn; myStage = (fl & SN_STAGE) != 0; }
            nScore += __popc(mine);
            nWin += take;
        }
        __syncwarp();
        // ---- copy the lists to shared memory (16-byte loads, all in flight together) and score, one candidate per lane
        {
            const uint32_t* kP = nullptr;
            const double* pP = nullptr;
            int bytes = 0;
            const int nk4 = int(myCnt & 0xffffu), np2 = int(myCnt >> 16);
            if (myW >= 0) {
                if (myStage) {
                    kP = t.key + 4 * (size_t)myKeyOff;
                    pP = t.pay + 2 * (size_t)myPayOff;
                    bytes = (nk4 + np2) * 16;
                } else {
                    const int64_t id = 3 * nN + W.winNode[myW];
                    kP = t.key + t.keyStart[id];
                    pP = t.pay + t.payStart[id];
                }
            }
            const int endOff = warp_incl_scan(bytes, lane);
            if (bytes && endOff <= poolBytes) {
                uint4* dk = W.pool + ((endOff - bytes) >> 4);
                uint4* dp = dk + nk4;
                const uint4* gk = reinterpret_cast<const uint4*>(kP);
                const uint4* gp = reinterpret_cast<const uint4*>(pP);
#pragma unroll 4
                for (int i = 0; i < nk4; i++) dk[i] = __ldg(gk + i);
#pragma unroll 4
                for (int i = 0; i < np2; i++) dp[i] = __ldg(gp + i);
                kP = reinterpret_cast<const uint32_t*>(dk);
                pP = reinterpret_cast<const double*>(dp);
            }
            if (st) {
                const long long now = clock64();
                if (lane == 0) { st[23] += (unsigned long long)(now - tk); st[19] += 1; st[20] += nScore; st[24] += nWin; }
                tk = now;
            }
            if (myW >= 0)
                W.winScore[myW] = (scanFlags & 1) ? f_append_q4(m, kP, pP, remK, remP, isRemovedTip, removedBLen)
                                               : f_append_sitewise(m, kP, pP, remK, remP, isRemovedTip, removedBLen);
        }
        __syncwarp();
        if (st) {
            const long long now = clock64();
            if (lane == 0) st[6] += (unsigned long long)(now - tk);
            tk = now;
        }
        // ---- replay of the reference's bookkeeping over the window (running best, failure counters, stop rule, jumps).
        // Parallel form.  What a node hands to its children -- midProb (its own score, or the inherited one when it is not
        // scored) and failedPasses (reset on a new best, +1 on a consecutive worsening, else inherited) -- are chains over
        // ancestors, and "reached" is the AND of the ancestors' stop-rule outcomes: three pointer-jumping passes over the
        // parent links of the window, log2(depth span) rounds each.  The running best before a node is taken as the prefix
        // maximum of the window's scores, which is right unless a node that the stop rule prunes holds a new best; that is
        // checked, and such a window (rare) is replayed node by node below instead.
        int j = 0;
        bool replayed = false;
        if (!(scanFlags & 2)) {
            constexpr int NC = kWinMax / 32;
            constexpr int FL_NB = 1, FL_DESC = 2, FL_OK = 4;
            int infc[NC], parc[NC], ptr[NC], fval[NC], flg[NC];
            double bb[NC];  // running best before node c*32+lane
            auto scoreOf = [&](int c, int w) { return (infc[c] & 7) == 7 ? W.winScore[w] : -INFINITY; };  // scored nodes keep their score throughout
            {
                double carry = best;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const int w = c * 32 + lane;
                    infc[c] = w < nWin ? W.winInfo[w] : 0;
                    parc[c] = w < nWin ? W.winParent[w] : -1;
                    double v = scoreOf(c, w);
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double x = __shfl_up_sync(FULL, v, o);
                        if (lane >= o) v = fmax(v, x);
                    }
                    double ex = __shfl_up_sync(FULL, v, 1);
                    if (lane == 0) ex = -INFINITY;
                    bb[c] = fmax(carry, ex);
                    carry = fmax(carry, __shfl_sync(FULL, v, 31));
                }
            }
            auto incoming = [&](int c) {  // what a node whose parent lies before the window inherits
                const int rel = infc[c] >> 8;
                return rel < kPathSm ? W.path[rel] : gpath[rel];
            };
            // pass 1: midProb handed down = score of the nearest scored ancestor-or-self, else what came in from above the window
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int w = c * 32 + lane;
                ptr[c] = -1;
                if (w < nWin) {
                    if ((infc[c] & 7) != 7) {
                        if (parc[c] < 0) W.winScore[w] = incoming(c).lk;
                        else ptr[c] = parc[c];
                    }
                    W.wp[w] = ptr[c];
                }
            }
            __syncwarp();
            for (int round = 0; round < 8; round++) {
                if (!__any_sync(FULL, ptr[0] >= 0 || ptr[1] >= 0 || ptr[2] >= 0)) break;
                int pp[NC];
                double pv[NC];
#pragma unroll
                for (int c = 0; c < NC; c++)
                    if (ptr[c] >= 0) { pp[c] = W.wp[ptr[c]]; pv[c] = W.winScore[ptr[c]]; }
                __syncwarp();
#pragma unroll
                for (int c = 0; c < NC; c++)
                    if (ptr[c] >= 0) {
                        const int w = c * 32 + lane;
                        if (pp[c] < 0) { W.winScore[w] = pv[c]; ptr[c] = -1; }
                        else ptr[c] = pp[c];
                        W.wp[w] = ptr[c];
                    }
                __syncwarp();
            }
            // pass 2: failedPasses handed down.  Own effect: SET 0 on a new best, ADD 1 on a consecutive worsening, else ADD 0.
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int w = c * 32 + lane;
                fval[c] = 0; flg[c] = 0; ptr[c] = -1;
                if (w < nWin) {
                    const bool scored = (infc[c] & 7) == 7;
                    const double sc = scoreOf(c, w);
                    double lkIn;
                    int failedIn = 0;
                    if (parc[c] >= 0) lkIn = W.winScore[parc[c]];
                    else { const PathE pe = incoming(c); lkIn = pe.lk; failedIn = pe.failed; }
                    const bool nb = scored && sc > bb[c];
                    const int inc = (scored && !nb && sc < (lkIn - sp.thresholdLogLKconsecutivePlacement)) ? 1 : 0;
                    if (nb) flg[c] = FL_NB;
                    else if (parc[c] < 0) fval[c] = failedIn + inc;
                    else { fval[c] = inc; ptr[c] = parc[c]; }
                    W.wp[w] = ptr[c];
                    W.wv[w] = fval[c];
                }
            }
            __syncwarp();
            for (int round = 0; round < 8; round++) {
                if (!__any_sync(FULL, ptr[0] >= 0 || ptr[1] >= 0 || ptr[2] >= 0)) break;
                int pp[NC], pf[NC];
#pragma unroll
                for (int c = 0; c < NC; c++)
                    if (ptr[c] >= 0) { pp[c] = W.wp[ptr[c]]; pf[c] = W.wv[ptr[c]]; }
                __syncwarp();
#pragma unroll
                for (int c = 0; c < NC; c++)
                    if (ptr[c] >= 0) {
                        const int w = c * 32 + lane;
                        fval[c] += pf[c];
                        ptr[c] = pp[c];
                        W.wp[w] = ptr[c];
                        W.wv[w] = fval[c];
                    }
                __syncwarp();
            }
            // stop rule of every node as if it were reached, then pass 3: reached = every ancestor in the window descends
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int w = c * 32 + lane;
                if (w < nWin) {
                    const int inf = infc[c];
                    const double mid = W.winScore[w];
                    const double bestAfter = fmax(bb[c], scoreOf(c, w));
                    const bool within = mid > (bestAfter - sp.thresholdLogLKtopology);
                    const bool rule = sp.strictTopologyStopRules ? (fval[c] <= sp.allowedFailsTopology && within)
                                                                 : (fval[c] <= sp.allowedFailsTopology || within);
                    const bool d = (inf & 4) && (inf & 3) != 1 && (inf & 8) && rule;
                    if (d) flg[c] |= FL_DESC;
                    W.wd[w] = d ? 1 : 0;
                }
            }
            __syncwarp();
            // ok[w] starts as "my parent descends" (true when the parent lies before the window: this window would not have been
            // reached otherwise) and absorbs the parent's ok, the grandparent's, ... by pointer jumping
            int ok[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int w = c * 32 + lane;
                ok[c] = 1; ptr[c] = -1;
                if (w < nWin) {
                    if (parc[c] >= 0) { ok[c] = W.wd[parc[c]]; ptr[c] = parc[c]; }
                    W.wp[w] = ptr[c];
                }
            }
            __syncwarp();  // wd has been read: wv (failedPasses, final) stays, wd is reused for ok
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int w = c * 32 + lane;
                if (w < nWin) W.wd[w] = ok[c];
            }
            __syncwarp();
            for (int round = 0; round < 8; round++) {
                if (!__any_sync(FULL, ptr[0] >= 0 || ptr[1] >= 0 || ptr[2] >= 0)) break;
                int pp[NC], po[NC];
#pragma unroll
                for (int c = 0; c < NC; c++)
                    if (ptr[c] >= 0) { pp[c] = W.wp[ptr[c]]; po[c] = W.wd[ptr[c]]; }
                __syncwarp();
#pragma unroll
                for (int c = 0; c < NC; c++)
                    if (ptr[c] >= 0) {
                        const int w = c * 32 + lane;
                        ok[c] &= po[c];
                        ptr[c] = pp[c];
                        W.wp[w] = ptr[c];
                        W.wd[w] = ok[c];
                    }
                __syncwarp();
            }
            // check the prefix-maximum assumption, then commit
            bool bad = false;
            int maxTarget = 0, nCounted = 0, anyNb = 0;
            double newBestVal = best;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int w = c * 32 + lane;
                if (w < nWin) {
                    const bool scored = (infc[c] & 7) == 7;
                    if (ok[c]) {
                        flg[c] |= FL_OK;
                        maxTarget = max(maxTarget, (flg[c] & FL_DESC) ? w + 1 : w + W.winSize[w]);
                        if (scored) { nCounted++; newBestVal = fmax(newBestVal, W.winScore[w]); anyNb |= flg[c] & FL_NB; }
                    } else if (scored && W.winScore[w] > bb[c]) bad = true;
                }
            }
            if (!__any_sync(FULL, bad)) {
                replayed = true;
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    maxTarget = max(maxTarget, __shfl_xor_sync(FULL, maxTarget, o));
                    nCounted += __shfl_xor_sync(FULL, nCounted, o);
                    newBestVal = fmax(newBestVal, __shfl_xor_sync(FULL, newBestVal, o));
                }
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const int w = c * 32 + lane;
                    const bool live = (flg[c] & FL_OK) != 0;
                    const bool queued = live && (infc[c] & 7) == 7 && W.winScore[w] > bb[c] - sp.thresholdLogLKoptimizationTopology;  // :7071
                    const unsigned qm = __ballot_sync(FULL, queued);
                    if (queued) {
                        const int at = qN + __popc(qm & ((1u << lane) - 1u));
                        if (at < qCap) qTop[-1 - at] = uint32_t(W.winNode[w]);
                    }
                    qN += __popc(qm);
                    // the last node of each depth that descends leaves its state for the windows that follow
                    const bool dsc = live && (flg[c] & FL_DESC);
                    const int rel1 = (infc[c] >> 8) + 1;
                    const unsigned peers = __match_any_sync(FULL, dsc ? rel1 : -1 - lane);
                    if (dsc && lane == 31 - __clz(peers)) {
                        if (rel1 >= pathCap) err = 3;
                        else if (rel1 < kPathSm) W.path[rel1] = PathE{W.winScore[w], fval[c], 0};
                        else gpath[rel1] = PathE{W.winScore[w], fval[c], 0};
                    }
                    __syncwarp();
                }
                if (qN > qCap) err = 3;
                err = __any_sync(FULL, err == 3) ? 3 : err;
                if (__any_sync(FULL, anyNb)) newBest = 1;
                best = newBestVal;
                phase1 += nCounted;
                j = maxTarget;
            } else if (st && lane == 0) st[25] += 1;  // the node-by-node replay only reads the scores of scored nodes: untouched above
        }
        if (!replayed) {
            int inf = W.winInfo[0], sz = W.winSize[0];
            double sc = W.winScore[0];
            while (j < nWin) {
                const int rel = inf >> 8;
                // both possible successors are fetched before the decision is known
                const int jA = j + 1, jB = j + sz;
                int infA = 0, szA = 1, infB = 0, szB = 1;
                double scA = 0.0, scB = 0.0;
                if (jA < nWin) { infA = W.winInfo[jA]; szA = W.winSize[jA]; scA = W.winScore[jA]; }
                if (jB < nWin) { infB = W.winInfo[jB]; szB = W.winSize[jB]; scB = W.winScore[jB]; }
                bool descend = false;
                if (rel + 1 < kPathSm) {
                    // straight-line form of the bookkeeping below (the usual case: the per-depth states are in shared memory)
                    const PathE pe = W.path[rel];
                    const bool scored = (inf & 7) == 7, dead = (inf & 3) == 1;  // dead: eligible but no probVectTotUp
                    const bool nb = scored && sc > best;
                    const bool queued = scored && sc > best - sp.thresholdLogLKoptimizationTopology;  // :7071
                    const double mid = scored ? sc : pe.lk;
                    const int failed = nb ? 0 : pe.failed + ((scored && sc < (pe.lk - sp.thresholdLogLKconsecutivePlacement)) ? 1 : 0);
                    if (queued) {
                        if (qN >= qCap) err = 3;
                        else qTop[-1 - qN] = uint32_t(W.winNode[j]);
                        qN++;
                    }
                    best = nb ? sc : best;
                    newBest |= nb ? 1 : 0;
                    phase1 += scored ? 1 : 0;
                    const bool within = mid > (best - sp.thresholdLogLKtopology);
                    const bool rule = sp.strictTopologyStopRules ? (failed <= sp.allowedFailsTopology && within)
                                                                 : (failed <= sp.allowedFailsTopology || within);
                    descend = (inf & 4) && !dead && (inf & 8) && rule;
                    W.path[rel + 1] = PathE{mid, failed, 0};  // every lane stores the same value; unused unless this node descends
                } else if (inf & 4) {
                    PathE pe;
                    if (rel < kPathSm) pe = W.path[rel];
                    else pe = gpath[rel];
                    double midProb = pe.lk;
                    int failed = pe.failed;
                    bool alive = true;
                    if (inf & 1) {
                        if (!(inf & 2)) alive = false;  // no probVectTotUp: the reference moves on without visiting the children
                        else {
                            midProb = sc;
                            phase1++;
                            if (midProb > best - sp.thresholdLogLKoptimizationTopology) {  // :7071
                                if (qN >= qCap) err = 3;
                                else qTop[-1 - qN] = uint32_t(W.winNode[j]);
                                qN++;
                            }
                            if (midProb > best) { best = midProb; failed = 0; newBest = 1; }
                            else if (midProb < (pe.lk - sp.thresholdLogLKconsecutivePlacement)) failed++;
                        }
                    }
                    if (alive && (inf & 8)) {
                        if (sp.strictTopologyStopRules) descend = failed <= sp.allowedFailsTopology && midProb > (best - sp.thresholdLogLKtopology);
                        else descend = failed <= sp.allowedFailsTopology || midProb > (best - sp.thresholdLogLKtopology);
                        if (descend) {
                            if (rel + 1 >= pathCap) { err = 3; descend = false; }
                            else if (rel + 1 < kPathSm) W.path[rel + 1] = PathE{midProb, failed, 0};
                            else gpath[rel + 1] = PathE{midProb, failed, 0};
                        }
                    }
                }
                if (descend) { j = jA; inf = infA; sz = szA; sc = scA; }
                else { j = jB; inf = infB; sz = szB; sc = scB; }
                if (err) break;
            }
        }
        pos += j;
        __syncwarp();
        if (st) {
            const long long now = clock64();
            if (lane == 0) st[7] += (unsigned long long)(now - tk);
            tk = now;
        }
    }
    __syncwarp();
    if (st && lane == 0) { st[21] += (unsigned long long)phase1; st[22] += (unsigned long long)qN; }
    if (lane == src) {
        f.bestLKdiff = best;
        f.phase1 += phase1;
        f.qN = qN;
        f.scanNewBest = newBest;
        if (err) s.err = err;
    }
}

// Per-lane scratch slices for warp_eval_queue: one ScratchD worth of room per lane of every warp of the launch.
struct EvalScratch {
    uint32_t* key;
    double* pay;
    double* ais;
    unsigned capK, capP, capA;  // per lane; key == nullptr: none (the owning lane evaluates its queue itself)
};

// The phase-2 entries a subtree scan queued for lane src's search (:7460-7639 for the entries found below a converged node),
// one entry per lane: evaluatePlacement (:6790) + the two appendProbNode calls of :7510-7511, with the lanes converged inside each
// co-walk; then folded into the search's best placement as the reference's loop does in order -- an entry replaces the best when
// its score is >= it (:7635), so the final best is the LAST entry that attains the maximum, if that maximum reaches the old best.
// On any failure (a slice too small, a None list) nothing is folded and the owning lane goes through the queue itself.
__device__ __noinline__ void warp_eval_queue(int src, Fsm& f, const DevModel& m, const DevTree& t, const SearchParams& sp, ScratchD& s, const EvalScratch& es,
                                size_t laneSlot, unsigned long long* st) {
    const unsigned FULL = 0xffffffffu;
    const int lane = int(threadIdx.x & 31);
    const int qN = __shfl_sync(FULL, f.qN, src);
    const uint32_t* qTop = reinterpret_cast<const uint32_t*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(s.key + s.capK + f.qRes), src));
    LRef removed;
    removed.k = reinterpret_cast<const uint32_t*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(f.removed.k), src));
    removed.p = reinterpret_cast<const double*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(f.removed.p), src));
    removed.nk = __shfl_sync(FULL, f.removed.nk, src);
    const bool isRemovedTip = __shfl_sync(FULL, f.isRemovedTip, src) != 0;
    double best = __shfl_sync(FULL, f.ph.bestScore, src);
    int bestNode = -1;
    double bT = 0.0, bB = 0.0, bA = 0.0;
    bool ok = es.key != nullptr;
    ScratchD ls;
    ls.key = es.key + laneSlot * es.capK;
    ls.pay = es.pay + laneSlot * es.capP;
    ls.ais = es.ais + laneSlot * es.capA;
    ls.capK = es.capK; ls.capP = es.capP; ls.capA = es.capA; ls.err = 0;
    __syncwarp();
    for (int base = 0; base < qN && ok; base += 32) {
        const int e = base + lane;
        const bool have = e < qN;
        double score = -INFINITY, cT = 0.0, cB = 0.0, cA = 0.0;
        int t1 = -1;
        bool bad = false;
        if (have) {
            t1 = int(ld_cg(qTop - 1 - e));
            ls.topK = ls.topP = 0;
            ls.err = 0;
            const LRef eUp = up_list_for(m, t, ls, t1), eDown = tree_list(t, 0, t1), eMidTot = tree_list(t, 3, t1);
            const bool fromTip1 = t.isTip[t1] != 0;
            const double eDist = t.dist[t1];
            double cost = 0.0;
            if (eval_placement(m, sp, ls, eMidTot, eDown, eUp, eDist, removed, isRemovedTip, fromTip1, cost, cB, cT, cA)) bad = true;
            else {
                const double initialCost = f_append(m, eUp, eDown, fromTip1, eDist);
                const double newPartialCost = f_append(m, eUp, eDown, fromTip1, cB + cT);
                score = cost + newPartialCost - initialCost;
            }
        }
        if (__any_sync(FULL, bad)) { ok = false; break; }
        // fold this batch: the maximum, and the last lane that holds it
        double mx = score;  // (a NaN score never replaces anything in the reference's `>=` either: fmax drops it)
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
        const unsigned at = __ballot_sync(FULL, have && score == mx);
        if (at && mx >= best) {
            const int w = 31 - __clz(at);
            best = mx;
            bestNode = __shfl_sync(FULL, t1, w);
            bT = __shfl_sync(FULL, cT, w);
            bB = __shfl_sync(FULL, cB, w);
            bA = __shfl_sync(FULL, cA, w);
        }
    }
    if (st && lane == 0) { st[ok ? 32 : 33] += (unsigned long long)qN; }
    if (lane == src) {
        f.evalqDone = ok ? 1 : 0;
        if (ok && bestNode >= 0) {
            f.ph.bestNode = bestNode;
            f.ph.bestScore = best;
            f.ph.bTop = bT;
            f.ph.bBottom = bB;
            f.ph.bAppend = bA;
        }
    }
    __syncwarp();
}

// After a search finished: startTopologyUpdatesParallel's acceptance logic (:9681-9702)
__device__ void fsm_finish(const Fsm& f, const DevTree& t, const SearchParams& sp, int node, double bestCurrentLK, SearchResult& r) {
    r.phase1 = f.phase1;
    r.status = f.rc;
    if (f.rc != 0) return;
    r.bestNode = f.ph.bestNode;
    r.bestScore = f.ph.bestScore;
    r.bLenTop = f.ph.bTop;
    r.bLenBottom = f.ph.bBottom;
    r.bLenAppend = f.ph.bAppend;
    if (f.ph.bestScore + sp.thresholdTopologyPlacement > bestCurrentLK) {
        bool updated = true;
        int topNode = t.up[node];
        if (f.ph.bestNode == topNode) updated = false;
        while (t.dist[topNode] == 0.0 && t.up[topNode] >= 0) topNode = t.up[topNode];
        if (f.ph.bestNode == topNode && f.ph.bBottom == 0.0) updated = false;
        const int sibling = f.child == 0 ? t.child1[f.parent] : t.child0[f.parent];
        if (f.ph.bestNode == sibling) updated = false;
        if (t.up[f.ph.bestNode] == sibling && f.ph.bTop == 0.0) updated = false;
        if (updated) {
            r.improvement = f.ph.bestScore - bestCurrentLK;
            r.placement = f.ph.bestNode;
        }
    }
}

// The warp's main loop of k_spr_search_fsm: every iteration each lane advances its control code to the next co-walk request,
// then the warp runs each kind of co-walk once for all lanes that requested it, then the pending subtree scans one after the
// other with the whole warp.  Lanes pull searches from the global counter.  SCAN2 selects the second form of the scans
// (scan2.cuh).  A function of its own so that tests/hostsim can run the very same loop with the lanes emulated.
// A few large scratch slots shared by the whole launch: a search that exhausts its lane's scratch takes one (while there are
// any) and starts over inside the same launch, concurrently with everybody else, instead of waiting for a second launch.
struct BigScratch {
    uint32_t* key;
    double* pay;
    double* ais;
    StackE* stack;
    unsigned capK, capP, capA;
    int nSlots;
    unsigned long long* counter;  // slots handed out so far
    unsigned long long* started;  // not null: every CTA of the launch counts itself here when it starts (see k_wait_started)
    unsigned long long headEnd;   // > 0: the first so many entries of the list are run one per warp (maple_ctx_set_head_searches)
};

template <bool SCAN2, bool EXTRAS>
__device__ __forceinline__ void fsm_warp_loop(const DevModel& sm, const DevTree& T, const SearchParams& sp, int64_t n, const int32_t* __restrict__ nodes,
                                              SearchResult* __restrict__ out, ScratchD s, StackE* stack, int stackCap, unsigned long long* counter,
                                              long long* outCycles, int scanMinSize, int scanFlags, int poolBytes, unsigned long long* st,
                                              const int32_t* outIndex, int lanesPerWarp, ScanSmem& W, Scan2Smem& W2, uint32_t& mbarParity,
                                              const BigScratch& big, int warpId, int totalWarps, const ScanQueue& sq, int ownerBase,
                                              const DenseScores& ds, const EvalScratch& es, size_t evalSlot) {
    const double* myRow = nullptr;  // dense scoring pass: the current search's row of precomputed candidate scores
    // Scan service (sq.cap != 0, SCAN2 only): a lane that needs a subtree scan posts the job in its slot sq.jobs[ownerBase + lane]
    // and waits for a warp of the serving SMs to run it; meanwhile the other lanes of this warp go on with their co-walks.
    const bool service = EXTRAS && SCAN2 && sq.cap != 0;
    bool waitScan = false, localScan = false;
    unsigned long long nCompleted = 0;
    // First search of every owning lane: entry lane * totalWarps + warpId of the list, so that neighbours in the list -- the
    // longest searches when the list is sorted longest-first -- start in different warps (the lanes of a warp share its time).
    // The counter then starts behind those entries (the host sets it).  totalWarps == 0: everything comes from the counter.
    bool firstPull = totalWarps > 0;
    // Head of the list (big.headEnd entries, the longest searches when the list is sorted): pulled by lane 0 of every warp only,
    // so each of them has a warp to itself and the warps share them out one at a time; the other lanes wait until the counter
    // has passed the head, then everybody pulls as usual.
    bool headOpen = big.headEnd == 0;
    unsigned waitTick = 0;
    const ScratchD sOwn = s;
    StackE* const stackOwn = stack;
    bool usingBig = false;
    const int lane_ = int(threadIdx.x & 31);
    const bool l0 = lane_ == 0;
    const bool stats = st != nullptr;
    long long tk = clock64();
#define STAT_T(i) do { if (st) { const long long now_ = clock64(); if (l0) st[i] += (unsigned long long)(now_ - tk); tk = now_; } } while (0)
#define STAT_N(i, v) do { if (st && l0) st[i] += (unsigned long long)(v); } while (0)
    Fsm f;
    f.op = OP_NONE;
    f.pc = 0;
    // lanes beyond lanesPerWarp own no search: they only lend a hand in the whole-warp subtree scans (fewer searches per warp =
    // a long search shares its warp's time with fewer others)
    int stage = (int(threadIdx.x & 31) < lanesPerWarp) ? 0 : 3;  // 0 idle, 1 current-placement append pending, 2 search running, 3 no more work
    unsigned long long i = 0;
    int node = -1;
    double bestCurrentLK = 0.0;
    long long c0 = 0;
    SearchResult r;
    for (;;) {
        // ---------------- control (divergent, cheap)
        if (waitScan) {  // my posted scan: served yet?
            ScanJob* J = sq.jobs + ownerBase + lane_;
            const int state = ld_volatile_i32(&J->state);
            if (state == 2) {
                __threadfence();
                f.bestLKdiff = ld_cg(&J->bestOut);
                f.phase1 += ld_cg(&J->phase1);
                f.qN = ld_cg(&J->qN);
                f.scanNewBest = ld_cg(&J->newBest);
                const int e = ld_cg(&J->err);
                if (e) s.err = e;
                waitScan = false;
            } else if (state == 3) {  // declined (the removed list does not fit a server's pool): this warp runs it itself
                waitScan = false;
                localScan = true;
            }
        }
        if (stage == 2 && !waitScan && !localScan) {
            fsm_step(f, sm, T, sp, s, stack, stackCap, scanMinSize);
        } else if (stage == 1) {
            bestCurrentLK = f.resD;
            r.bestCurrentLK = bestCurrentLK;
            if (!(bestCurrentLK < sp.thresholdTopologyPlacement || T.dist[node] != 0.0)) {  // :9674
                f.rc = 1;
                f.op = OP_DONE;
            } else {
                const int parent = T.up[node];
                f.pc = 0;
                f.parent = parent;
                f.child = (T.child0[parent] == node) ? 0 : 1;
                f.bestLKdiff = bestCurrentLK;
                f.removedBLen = T.dist[node];
                f.phase1 = 0;
                f.rc = 0;
                s.topK = s.topP = 0;
                stage = 2;
                fsm_step(f, sm, T, sp, s, stack, stackCap, scanMinSize);
            }
        }
        while (stage != 3 && (stage == 0 || f.op == OP_DONE)) {
            if (stage != 0) {  // a search (or its pre-check) just ended
                if (stage == 2 && f.rc == 3 && !usingBig && big.nSlots > 0) {  // scratch exhausted: once more with a large slot
                    const unsigned long long slot = atomicAdd(big.counter, 1ULL);
                    if (slot < (unsigned long long)big.nSlots) {
                        usingBig = true;
                        s.key = big.key + slot * big.capK;
                        s.pay = big.pay + slot * big.capP;
                        s.ais = big.ais + slot * big.capA;
                        s.capK = big.capK; s.capP = big.capP; s.capA = big.capA; s.topK = s.topP = 0; s.err = 0;
                        stack = big.stack + slot * (size_t)stackCap;
                        f.pc = 0; f.op = OP_NONE;
                        f.bestLKdiff = bestCurrentLK;
                        f.removedBLen = T.dist[node];
                        f.phase1 = 0; f.rc = 0;
                        fsm_step(f, sm, T, sp, s, stack, stackCap, scanMinSize);
                        continue;
                    }
                }
                if (usingBig) {
                    usingBig = false;
                    s = sOwn;
                    stack = stackOwn;
                }
                if (stage == 2) fsm_finish(f, T, sp, node, bestCurrentLK, r);
                else r.status = f.rc;
                out[outIndex ? outIndex[i] : i] = r;
                if (outCycles) outCycles[i] = clock64() - c0;
                stage = 0;
                nCompleted++;
            }
            if (!headOpen && lane_ != 0) {  // the head is still being handed out: not my turn yet (look again every 64 iterations)
                firstPull = false;
                if ((waitTick++ & 63u) == 0u && ld_volatile_u64(counter) >= big.headEnd) headOpen = true;
                if (!headOpen) { f.op = OP_NONE; break; }
            }
            if (firstPull) {
                firstPull = false;
                i = (unsigned long long)lane_ * (unsigned long long)totalWarps + (unsigned long long)warpId;
            } else i = atomicAdd(counter, 1ULL);
            if (i >= (unsigned long long)n) { stage = 3; f.op = OP_NONE; break; }
            node = nodes[i];
            myRow = nullptr;
            if (EXTRAS && ds.rowOf) {
                const int row = ds.rowOf[i];
                if (row >= 0) myRow = ds.scores + (size_t)row * (size_t)ds.stride;
            }
            c0 = clock64();
            r.placement = -1; r.bestNode = -1; r.status = 1; r.phase1 = 0;
            r.improvement = r.bestCurrentLK = r.bestScore = r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
            f.op = OP_NONE;
            if (T.up[node] < 0) { out[outIndex ? outIndex[i] : i] = r; nCompleted++; continue; }
            s.topK = s.topP = 0;
            s.err = 0;
            const int parent = T.up[node];
            LRef vectUp = (T.child0[parent] == node) ? tree_list(T, 1, parent) : tree_list(T, 2, parent);
            if (n_mut(T, node)) vectUp = s_pass(sm, T, s, vectUp, node, false);
            const LRef own = tree_list(T, 0, node);
            if (!vectUp.k || !own.k) { r.status = s.err ? s.err : 2; out[outIndex ? outIndex[i] : i] = r; nCompleted++; continue; }
            f.a1 = vectUp; f.a2 = own; f.at1 = T.isTip[node] != 0; f.ab1 = T.dist[node];
            f.op = OP_APPEND;
            stage = 1;
        }
        // ---------------- co-walks, one kind at a time, lanes converged
        __syncwarp();
        STAT_T(0);
        if (stats) {
            const unsigned b1 = __ballot_sync(0xffffffffu, f.op == OP_APPEND), b2 = __ballot_sync(0xffffffffu, f.op == OP_MERGE),
                           b3 = __ballot_sync(0xffffffffu, f.op == OP_BLEN), b4 = __ballot_sync(0xffffffffu, f.op == OP_DIFFER);
            STAT_N(8, __popc(b1)); STAT_N(9, __popc(b2)); STAT_N(10, __popc(b3)); STAT_N(11, __popc(b4));
            STAT_N(12, b1 != 0); STAT_N(13, b2 != 0); STAT_N(14, b3 != 0); STAT_N(15, b4 != 0); STAT_N(16, 1);
            tk = clock64();
        }
        if (f.op == OP_APPEND) f.resD = f_append(sm, f.a1, f.a2, f.at1 != 0, f.ab1);
        __syncwarp();
        STAT_T(1);
        if (f.op == OP_MERGE) {
            Writer w;
            w.init(s.key + s.topK, s.pay + s.topP);
            if (f_merge(sm, f.a1, f.ab1, f.at1 != 0, f.a2, f.ab2, f.at2 != 0, f.aflags, w) == 0) f.resL = sc_commit(s, w.nk, w.np);
            else f.resL = lnull();
        }
        __syncwarp();
        STAT_T(2);
        if (f.op == OP_BLEN) f.resD = f_blen(sm, f.a1, f.a2, f.at1 != 0, s.ais);
        __syncwarp();
        STAT_T(3);
        if (f.op == OP_DIFFER) f.resB = f_differ(sm, f.a1, f.a2) ? 1 : 0;
        __syncwarp();
        STAT_T(4);
        // ---------------- subtree scans
        if (service) {
            // post the new requests; the serving warps take them from the ring
            if (f.op == OP_SCAN && !waitScan && !localScan) {
                ScanJob* J = sq.jobs + ownerBase + lane_;
                J->R = f.t1; J->pruned = f.pruned; J->sibling = f.sibling; J->failed0 = f.failedPasses;
                J->best = f.bestLKdiff; J->lastLK0 = f.lastLK; J->removedBLen = f.removedBLen;
                J->isRemovedTip = f.isRemovedTip;
                J->remK = f.removed.k; J->remP = f.removed.p;
                J->gpath = reinterpret_cast<PathE2*>(stack + f.spN);
                J->pathCap = int((size_t)(stackCap - f.spN) * sizeof(StackE) / sizeof(PathE2));
                J->qTop = s.key + s.capK;
                J->qCap = int(s.capK - s.topK) - 8;
                J->scoreRow = myRow;
                J->state = 1;
                __threadfence();
                const unsigned long long ticket = atomicAdd(sq.tail, 1ULL);
                *reinterpret_cast<volatile unsigned long long*>(sq.ring + (ticket & (sq.cap - 1))) =
                    ((ticket + 1ULL) << 32) | (unsigned long long)(ownerBase + lane_);
                waitScan = true;
                STAT_N(26, 1);
            }
            // nothing to do for anybody until a scan comes back: do not hammer the slots
            if (__all_sync(0xffffffffu, stage == 3 || waitScan)) spin_pause(400);
        }
        // scans this warp runs itself (no service, or declined by it): the whole warp works for one lane's search at a time
        for (unsigned pending = __ballot_sync(0xffffffffu, f.op == OP_SCAN && (!service || localScan)); pending; pending &= pending - 1) {
            const int src = __ffs(pending) - 1;
            if (SCAN2) {
                if (lane_ == src) {  // the request, for the whole warp to read
                    ScanJob& J = W2.job;
                    J.R = f.t1; J.pruned = f.pruned; J.sibling = f.sibling; J.failed0 = f.failedPasses;
                    J.best = f.bestLKdiff; J.lastLK0 = f.lastLK; J.removedBLen = f.removedBLen;
                    J.isRemovedTip = f.isRemovedTip;
                    J.remK = f.removed.k; J.remP = f.removed.p;
                    // deep per-depth state lives in the unused part of the DFS stack; the phase-2 queue grows down from the top of the key scratch
                    J.gpath = reinterpret_cast<PathE2*>(stack + f.spN);
                    J.pathCap = int((size_t)(stackCap - f.spN) * sizeof(StackE) / sizeof(PathE2));
                    J.qTop = s.key + s.capK;
                    J.qCap = int(s.capK - s.topK) - 8;
                    J.scoreRow = myRow;
                }
                __syncwarp();
                warp_scan_job2<EXTRAS>(sm, T, sp, W2, poolBytes, scanFlags, mbarParity, st, false);
                if (lane_ == src) {
                    const ScanJob& J = W2.job;
                    f.bestLKdiff = J.bestOut;
                    f.phase1 += J.phase1;
                    f.qN = J.qN;
                    f.scanNewBest = J.newBest;
                    if (J.err) s.err = J.err;
                    localScan = false;
                }
                __syncwarp();
            } else warp_scan_job(src, f, sm, T, sp, s, stack, stackCap, W, poolBytes, scanFlags, st);
        }
        // ---------------- queued phase-2 entries: one per lane
        for (unsigned pending = __ballot_sync(0xffffffffu, f.op == OP_EVALQ); pending; pending &= pending - 1)
            warp_eval_queue(__ffs(pending) - 1, f, sm, T, sp, s, es, evalSlot, st);
        STAT_T(5);
        if (__all_sync(0xffffffffu, stage == 3)) break;
    }
    if (EXTRAS && service) {  // this warp's searches are over: tell the servers, then help them until everybody is through
        for (int o = 16; o; o >>= 1) nCompleted += __shfl_xor_sync(0xffffffffu, nCompleted, o);
        if (l0 && nCompleted) atomicAdd(sq.doneSearches, nCompleted);
        __syncwarp();
        scan_server_loop(sm, T, sp, W2, poolBytes, scanFlags, mbarParity, st, sq, n);
    }
#undef STAT_T
#undef STAT_N
}

}  // namespace maple
