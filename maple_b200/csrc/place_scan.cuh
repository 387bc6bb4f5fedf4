// Placement of a new sample on a frozen tree by one WARP (findBestParentForNewSample, MAPLEv0.7.5.4.py:7912-8292).
//
// The walk of the reference (:7972-8093) only reads STORED lists: the candidate score at a node is
// appendProbNode(probVectTotUp[node], sample, True, oneMutBLen) and whether a node is visited was decided at its parent.
// In the pre-order the walk follows (DevTree::order: a node, the subtree of child 1, the subtree of child 0) the tree below
// the root is one contiguous range and a pruned subtree a contiguous sub-range.  So the warp takes a window of up to 96
// positions with at most 32 nodes that need a score, scores them one per lane (all lanes inside appendProbNode together),
// runs isMinorSequence for the leaves of the window the same way, and then ONE lane replays the reference's bookkeeping over
// the window in order -- running best (>=), failedPasses (reset on a new best, then +1 on a consecutive worsening, :8070),
// the bestNodes list (:8066, :8068), the stop rule, the jump over a pruned subtree, the early return when a leaf absorbs the
// sample (:7985-8003).  The replay is the reference's loop verbatim, on precomputed scores: nothing the reference would not
// have visited is counted or kept.  The refinement of the bestNodes entries (:8109-8187) is independent per entry: one entry
// per lane, then the sequential `>=` selection.
//
// Written phase by phase: every phase is a loop body over lanes (FOR_LANES) that only reads what earlier phases left in
// the per-warp block W and writes its own slots, with a warp barrier in between.  On the device a phase is executed by the 32
// lanes at once; compiled for the host (tests/hostsim) the same source runs the lanes of a phase one after the other, which
// is how this file is held to the reference's recorded placements without a GPU.
//
// Samples the scan does not cover fall back to place_sample() on lane 0 with the warp's whole scratch: trees with MAT
// mutations, --deeperSearchForLongBranches, a root without children, a sample list shorten() would change, a scored node
// without probVectTotUp (the reference raises there).
//
// Three entry points, kept apart until each has been timed on hardware:
//   place_sample_warp            variant 1: MAT-free trees, one-lane window replay -- the form that has run on a B200;
//   place_sample_warp_mat<false> variant 2: MAT trees covered (lane 0 walks above mutation-carrying nodes, mutation-free subtrees
//                                are scan jobs), one-lane window replay;
//   place_sample_warp_mat<true>  variant 3: variant 2 with every per-window step in parallel form (slots by prefix sum, replay by
//                                prefix maximum + pointer jumping, reductions as trees, leaf comparisons only where reached).
#pragma once
#include "place.cuh"

#ifdef MAPLE_HOST_LANES
#define FOR_LANES(lane) for (int lane = 0; lane < 32; lane++)
#define WARP_SYNC() ((void)0)
#else
#define FOR_LANES(lane) for (int lane = int(threadIdx.x & 31u), once_ = 1; once_; once_ = 0)
#define WARP_SYNC() __syncwarp()
#endif

namespace maple {

constexpr int kPWin = 96;   // pre-order positions per window
constexpr int kPPath = 48;  // per-depth states kept in the warp block (deeper ones in global scratch)

struct PlacePath {  // what a node hands to its children (:8079, :8086)
    double lk;
    int failed, pad;
};

// per-warp block (shared memory on the device)
struct PlaceWarp {
    double winScore[kPWin];
    int winNode[kPWin], winSize[kPWin], winInfo[kPWin];  // info: 1 needs a score, 2 leaf, 4 long branch without probVectTotUp, bits 8.. depth below the root
    int winMinor[kPWin];
    int slot[32];  // window position scored by lane k, or -1
    PlacePath path[kPPath];
    LRef diffs;
    double best, original;
    int bestNode, phase1, missed, nQ, nWin, pos;
    int state;  // 0 walking, 1 absorbed by leaf minorNode, 2 fall back to place_sample, 3 scratch exhausted
    int minorNode;
};

// per-warp global scratch
struct PlaceWarpScratch {
    double* pay;     // 32 lane slices of laneP doubles
    double* ais;     // 32 x laneA
    uint32_t* key;   // 32 x laneK
    PlaceBest* best;  // bestNodes, bestCap entries
    PlaceEval* eval;  // refinement results, bestCap entries
    int* evalRc;
    PlacePath* gpath;     // stackCap entries
    PlaceStackE* stack;   // stackCap entries (fallback walk)
    uint32_t* diffKey;    // the sample list: laneK keys ...
    double* diffPay;      // ... 6 * laneK doubles
    unsigned laneK, laneP, laneA;
    int bestCap, stackCap;
};

__device__ __forceinline__ ScratchD place_lane_scratch(const PlaceWarpScratch& ws, int lane) {
    ScratchD s;
    s.key = ws.key + (size_t)lane * ws.laneK;
    s.pay = ws.pay + (size_t)lane * ws.laneP;
    s.ais = ws.ais + (size_t)lane * ws.laneA;
    s.capK = ws.laneK; s.capP = ws.laneP; s.capA = ws.laneA; s.topK = s.topP = 0; s.err = 0;
    return s;
}

__device__ __forceinline__ ScratchD place_whole_scratch(const PlaceWarpScratch& ws) {
    ScratchD s;
    s.key = ws.key; s.pay = ws.pay; s.ais = ws.ais;
    s.capK = 32u * ws.laneK; s.capP = 32u * ws.laneP; s.capA = 32u * ws.laneA; s.topK = s.topP = 0; s.err = 0;
    return s;
}

__device__ __noinline__ double p_append_sitewise(const DevModel& m, LRef P, LRef C, double bLen) {
    return dev_append_sitewise<false>(m, P.k, P.p, C.k, C.p, true, bLen);
}

// r is written by lane 0
__device__ void place_sample_warp(const DevModel& m, const DevTree& t, const PlaceParams& pp, LRef in, PlaceWarp& W, const PlaceWarpScratch& ws,
                                  PlaceResult& r) {
    const int root = t.root;
    const double one = pp.oneMutBLen;
    // ---- preamble (lane 0): the sample list, the cost of hanging it from the root (:7962-7963)
    FOR_LANES(lane) {
        if (lane == 0) {
            W.state = 0;
            W.bestNode = root; W.phase1 = 0; W.missed = 0; W.nQ = 0;
            W.pos = 0; W.nWin = 0; W.minorNode = -1;
            W.diffs = lnull();
            const bool covered = t.order && !pp.deeperSearchForLongBranches && t.child0[root] >= 0 && !n_mut(t, root) &&
                                 !(t.mutStart && t.mutBelow[root]) && in.k && unsigned(in.nk) <= ws.laneK;
            if (!covered) W.state = 2;
            else {
                ScratchD s = place_lane_scratch(ws, 0);
                Writer w;
                w.init(ws.diffKey, ws.diffPay);
                dev_shorten<true>(m, in.k, in.p, w);  // a list shorten() leaves alone is copied as it is
                const LRef rootVect = w.nk == in.nk ? s_root_vector(m, t, s, tree_list(t, 0, root), 0.0, false) : lnull();
                if (!rootVect.k) W.state = 2;
                else {
                    W.diffs = LRef{ws.diffKey, ws.diffPay, w.nk};
                    W.best = W.original = f_append(m, rootVect, W.diffs, true, one);
                    W.path[1] = PlacePath{W.best, 0, 0};
                    W.pos = t.pre[root] + 1;
                }
            }
        }
    }
    WARP_SYNC();
    const int end = t.order ? t.pre[root] + t.size[root] : 0;
    const int d0 = t.order ? t.depth[root] : 0;
    while (W.state == 0 && W.pos < end) {
        const int pos = W.pos;
        // ---- window: records of positions pos .. pos+95
        FOR_LANES(lane) {
            for (int w = lane; w < kPWin; w += 32) {
                const int idx = pos + w;
                int info = 0, size = 1, node = -1;
                if (idx < end) {
                    const ScanNode rec = t.scan[idx];
                    node = rec.node;
                    size = rec.size;
                    const bool isLong = (rec.flags & SN_LONG) != 0, tot = (rec.flags & SN_TOT) != 0;
                    info = ((isLong && tot) ? 1 : 0) | ((rec.flags & SN_INNER) ? 0 : 2) | ((isLong && !tot) ? 4 : 0) | ((rec.depth - d0) << 8);
                }
                W.winInfo[w] = info; W.winSize[w] = size; W.winNode[w] = node;
            }
        }
        WARP_SYNC();
        // ---- at most 32 nodes to score: the window ends before the 33rd
        FOR_LANES(lane) {
            if (lane == 0) {
                int nWin = min(kPWin, end - pos), k = 0;
                for (int w = 0; w < nWin; w++) {
                    if (W.winInfo[w] & 1) {
                        if (k == 32) { nWin = w; break; }
                        W.slot[k++] = w;
                    }
                }
                for (; k < 32; k++) W.slot[k] = -1;
                W.nWin = nWin;
            }
        }
        WARP_SYNC();
        // ---- scores (:8050) and leaf comparisons (:7975-7984), one node per lane
        FOR_LANES(lane) {
            const int w = W.slot[lane];
            if (w >= 0) W.winScore[w] = p_append_sitewise(m, tree_list(t, 3, W.winNode[w]), W.diffs, one);
        }
        FOR_LANES(lane) {
            for (int w = lane; w < W.nWin; w += 32)
                if (W.winInfo[w] & 2) W.winMinor[w] = dev_is_minor(m.lRef, tree_list(t, 0, W.winNode[w]), W.diffs, pp.onlyFindIdentical != 0);
        }
        WARP_SYNC();
        // ---- the reference's loop body over the window, in order (lane 0)
        FOR_LANES(lane) {
            if (lane == 0) {
                int j = 0;
                const int nWin = W.nWin;
                double best = W.best;
                while (j < nWin) {
                    const int info = W.winInfo[j], rel = info >> 8, node = W.winNode[j];
                    const PlacePath pe = rel < kPPath ? W.path[rel] : ws.gpath[rel];
                    int failed = pe.failed;
                    double LK = pe.lk;
                    if (info & 2) {
                        const int cmp = W.winMinor[j];
                        if (cmp == 1) { W.state = 1; W.minorNode = node; break; }
                        if (cmp == 2) W.missed++;
                    }
                    if (info & 4) { W.state = 2; break; }
                    if (info & 1) {
                        LK = W.winScore[j];
                        W.phase1++;
                        const bool nb = LK >= best;
                        if (nb || LK > best - pp.thresholdLogLKoptimization) {
                            if (W.nQ >= ws.bestCap) { W.state = 3; break; }
                            PlaceBest& b = ws.best[W.nQ++];
                            b.t1 = node; b.score = LK; b.diffs = W.diffs;
                        }
                        if (nb) { best = LK; W.bestNode = node; failed = 0; }
                        if (LK < (pe.lk - pp.thresholdLogLKconsecutivePlacement)) failed++;
                    }
                    const bool within = LK > (best - pp.thresholdLogLK);
                    const bool go = pp.strictStopRules ? (failed <= pp.allowedFails && within) : (failed <= pp.allowedFails || within);
                    if (go && !(info & 2)) {
                        if (rel + 1 >= ws.stackCap) { W.state = 3; break; }
                        if (rel + 1 < kPPath) W.path[rel + 1] = PlacePath{LK, failed, 0};
                        else ws.gpath[rel + 1] = PlacePath{LK, failed, 0};
                        j += 1;
                    } else j += W.winSize[j];
                }
                W.best = best;
                W.pos = pos + j;
            }
        }
        WARP_SYNC();
    }
    // ---- endings that need no refinement
    if (W.state != 0) {
        FOR_LANES(lane) {
            if (lane == 0) {
                if (W.state == 2) {  // not covered by the scan: the straight-line walk with the whole scratch
                    ScratchD s = place_whole_scratch(ws);
                    place_sample(m, t, pp, in, s, ws.stack, ws.stackCap, ws.best, ws.bestCap, r);
                } else {
                    r.bestNode = W.state == 1 ? W.minorNode : -1;
                    r.status = W.state;
                    r.phase1 = W.phase1;
                    r.missedMinors = W.missed;
                    r.bestScore = W.state == 1 ? 1.0 : 0.0;
                    r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
                    if (W.state == 3) r.phase1 = r.missedMinors = 0;
                }
            }
        }
        WARP_SYNC();
        return;
    }
    // ---- refinement of the bestNodes entries within thresholdLogLKoptimization of the final best (:8109-8187), one per lane
    FOR_LANES(lane) {
        for (int i = lane; i < W.nQ; i += 32) {
            int rc = -1;  // -1: filtered out
            if (ws.best[i].score >= W.best - pp.thresholdLogLKoptimization) {
                ScratchD s = place_lane_scratch(ws, lane);
                rc = place_refine_entry(m, t, s, ws.best[i].t1, ws.best[i].diffs, ws.eval[i]);
            }
            ws.evalRc[i] = rc;
        }
    }
    WARP_SYNC();
    FOR_LANES(lane) {
        if (lane == 0) {
            int bestNode = W.bestNode, status = 0;
            double bestScore = W.best;
            double bTop = 0.0, bBottom = 0.0, bAppend = one;  // (False, False, oneMutBLen), :7930
            if (bestNode != root) {  // lengths recorded with the last new best of the walk (:8067)
                bTop = t.dist[bestNode] / 2;
                bBottom = t.dist[bestNode] / 2 / 2;
            }
            for (int i = 0; i < W.nQ; i++) {
                int rc = ws.evalRc[i];
                if (rc < 0) continue;
                PlaceEval e = ws.eval[i];
                if (rc == 3) {  // the lists around this branch did not fit a lane's slice (long upper lists near the root): whole scratch
                    ScratchD s = place_whole_scratch(ws);
                    rc = place_refine_entry(m, t, s, ws.best[i].t1, ws.best[i].diffs, e);
                }
                if (rc > 0) { status = rc; break; }
                if (e.score >= bestScore) {
                    bestNode = ws.best[i].t1;
                    bestScore = e.score;
                    bTop = e.top; bBottom = e.bottom; bAppend = e.append;
                }
            }
            r.phase1 = W.phase1;
            r.missedMinors = W.missed;
            r.status = status;
            if (status) {
                r.bestNode = -1;
                r.bestScore = r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
            } else {
                if (bestScore == -INFINITY) bestScore = W.original;
                r.bestNode = bestNode;
                r.bestScore = bestScore;
                r.bLenTop = bTop; r.bLenBottom = bBottom; r.bLenAppend = bAppend;
            }
        }
    }
    WARP_SYNC();
}

// ---- the same with MAT trees covered -------------------------------------------------------------------------------------
// On a tree with local references (mutations[node] lists, the reference's default) the sample list is re-referenced whenever
// the walk crosses a branch that carries mutations (:7969, :8082, :8089), so one list no longer serves the whole tree.  Here lane 0
// runs the reference's stack-driven walk itself and hands every popped node whose subtree carries no mutations below it (and is
// worth a window) to the warp as a scan job: the subtree is a contiguous pre-order range scored against ONE list, the one the
// stack entry carries, exactly as in place_sample_warp; what the job's root inherits (parentLK, failedPasses) comes from the
// entry.  Nodes outside such subtrees -- the ancestors of mutation-carrying nodes -- are processed by lane 0 in the straight-line
// form.  shorten() of the current list at a new best (:8065) is applied in place like the reference does; candidate scores and
// leaf comparisons do not depend on how R runs are split, the refinement does, and by then every list has been shortened if
// and only if a new best was found while it was the current one, as in the reference.
// Scratch: the walk allocates re-referenced lists from the bottom of the warp's scratch and never frees them (bestNodes
// entries point at them); the refinement lanes share what is left above.
constexpr int kPlaceScanMin = 4;  // subtrees smaller than this stay on lane 0

struct PlaceWarpMat {
    PlaceWarp w;
    ScratchD walk;       // lane 0's allocator over the whole scratch
    PlaceStackE cur;     // the entry being processed
    int sp, job, jobNewBest;
    // work arrays of the parallel window replay (place_replay_parallel)
    double bb[kPWin];       // running best before the node, taking every earlier score of the window
    double lkOut[kPWin];    // LKdiff the node hands to its children
    double dval[2][kPWin];  // double-buffered values of the scans / pointer jumps
    int fOut[kPWin];        // failedPasses the node hands to its children
    int par[kPWin];         // window position of the parent, -1: before the window
    int ival[2][kPWin], ptr[2][kPWin];
    int flags[kPWin];       // 1 new best, 2 descends, 4 reached
    double redD[32];
    int redI[32], redJ[32], redK[32];
    int act[8];
    int cutW, maxTarget, committed, qBase;
};

// Pointer jumping over the parent links of a window: every node with ptr >= 0 combines its value with its pointer target's and
// takes over the target's pointer, until all chains end at a resolved node (ptr < 0).  MODE 0: the value of the resolved ancestor
// (double), 1: sum (int), 2: logical and (int).  Values and pointers are double-buffered; returns the buffer that holds the
// result, or -1 when 8 rounds were not enough (cannot happen for 96 nodes; the caller then replays serially).
template <int MODE>
__device__ int place_jump(PlaceWarpMat& X, int nWin) {
    FOR_LANES(lane) {
        if (lane < 8) X.act[lane] = 0;
    }
    WARP_SYNC();
    int cur = 0;
    for (int round = 0; round < 8; round++) {
        const int nxt = cur ^ 1;
        FOR_LANES(lane) {
            for (int w = lane; w < nWin; w += 32) {
                const int p = X.ptr[cur][w];
                if (p < 0) {
                    X.ptr[nxt][w] = -1;
                    if (MODE == 0) X.dval[nxt][w] = X.dval[cur][w];
                    else X.ival[nxt][w] = X.ival[cur][w];
                } else {
                    const int pp = X.ptr[cur][p];
                    if (MODE == 0) X.dval[nxt][w] = X.dval[cur][p];  // meaningful once p is resolved, i.e. when pp < 0
                    else if (MODE == 1) X.ival[nxt][w] = X.ival[cur][w] + X.ival[cur][p];
                    else X.ival[nxt][w] = X.ival[cur][w] & X.ival[cur][p];
                    X.ptr[nxt][w] = pp;
                    if (pp >= 0) X.act[round] = 1;
                }
            }
        }
        WARP_SYNC();
        cur = nxt;
        if (!X.act[round]) return cur;
    }
    return -1;
}

// Inclusive prefix sum over the 32 lane values in X.ival[0][0..31] (Hillis-Steele, double-buffered); returns the buffer holding it.
__device__ int place_lane_scan(PlaceWarpMat& X) {
    int cur = 0;
    for (int o = 1; o < 32; o <<= 1) {
        const int nxt = cur ^ 1;
        FOR_LANES(lane) { X.ival[nxt][lane] = lane >= o ? X.ival[cur][lane] + X.ival[cur][lane - o] : X.ival[cur][lane]; }
        WARP_SYNC();
        cur = nxt;
    }
    return cur;
}

// Tree reductions over the per-lane partials redD (max), redI (sum), redJ (max), redK (min), in place; results in element 0.
// A round only writes elements below its stride and only reads elements at or above it from other lanes.
__device__ void place_lane_reduce(PlaceWarpMat& X) {
    for (int o = 16; o > 0; o >>= 1) {
        FOR_LANES(lane) {
            if (lane < o) {
                X.redD[lane] = fmax(X.redD[lane], X.redD[lane + o]);
                X.redI[lane] += X.redI[lane + o];
                X.redJ[lane] = max(X.redJ[lane], X.redJ[lane + o]);
                X.redK[lane] = min(X.redK[lane], X.redK[lane + o]);
            }
        }
        WARP_SYNC();
    }
}

// The nodes of the window that need a score, in order, one per lane; the window ends before the 33rd (parallel form of the loop
// in place_sample_warp).  Three consecutive positions per lane, a prefix sum over the lanes.
__device__ void place_assign_slots(PlaceWarpMat& X, int nWin0) {
    PlaceWarp& W = X.w;
    FOR_LANES(lane) {
        int c = 0;
        for (int w = 3 * lane; w < 3 * lane + 3 && w < nWin0; w++) c += W.winInfo[w] & 1;
        X.ival[0][lane] = c;
        W.slot[lane] = -1;
        if (lane == 0) W.nWin = nWin0;
    }
    WARP_SYNC();
    const int cur = place_lane_scan(X);
    FOR_LANES(lane) {
        int k = lane > 0 ? X.ival[cur][lane - 1] : 0;
        for (int w = 3 * lane; w < 3 * lane + 3 && w < nWin0; w++) {
            if (W.winInfo[w] & 1) {
                if (k < 32) W.slot[k] = w;
                else if (k == 32) W.nWin = w;
                k++;
            }
        }
    }
    WARP_SYNC();
}

// The window replay of place_sample_warp in parallel form.  What a node hands to its children -- LKdiff (its own score, or the
// inherited one when it is not scored) and failedPasses (reset on a new best, +1 on a consecutive worsening, else inherited) --
// are chains over ancestors, and "reached" is the AND of the ancestors' stop-rule outcomes: three pointer-jumping passes over the
// window's parent links.  The running best before a node is taken as the prefix maximum of ALL earlier scores of the window, which
// is right unless a node the walk does not reach holds a score above it; that is checked, and such a window is left to the serial
// replay (returns false, nothing committed).  Everything after the first reached leaf that absorbs the sample is void.
__device__ bool place_replay_parallel(const DevModel& m, const DevTree& t, const PlaceParams& pp, PlaceWarpMat& X, const PlaceWarpScratch& ws,
                                      int pos) {
    PlaceWarp& W = X.w;
    const int nWin = W.nWin;
    // ---- running best before every node: exclusive prefix maximum of the scores (Hillis-Steele over the window)
    FOR_LANES(lane) {
        for (int w = lane; w < nWin; w += 32) X.dval[0][w] = (W.winInfo[w] & 1) ? W.winScore[w] : -INFINITY;
    }
    WARP_SYNC();
    int cur = 0;
    for (int o = 1; o < nWin; o <<= 1) {
        const int nxt = cur ^ 1;
        FOR_LANES(lane) {
            for (int w = lane; w < nWin; w += 32) X.dval[nxt][w] = w >= o ? fmax(X.dval[cur][w], X.dval[cur][w - o]) : X.dval[cur][w];
        }
        WARP_SYNC();
        cur = nxt;
    }
    FOR_LANES(lane) {
        for (int w = lane; w < nWin; w += 32) X.bb[w] = w > 0 ? fmax(W.best, X.dval[cur][w - 1]) : W.best;
    }
    WARP_SYNC();
    // ---- pass 1: LKdiff handed down = own score, else the nearest scored ancestor's, else what came in from above the window
    FOR_LANES(lane) {
        for (int w = lane; w < nWin; w += 32) {
            const int info = W.winInfo[w], rel = info >> 8, p = X.par[w];
            if (info & 1) { X.ptr[0][w] = -1; X.dval[0][w] = W.winScore[w]; }
            else if (p < 0) { X.ptr[0][w] = -1; X.dval[0][w] = (rel < kPPath ? W.path[rel] : ws.gpath[rel]).lk; }
            else { X.ptr[0][w] = p; X.dval[0][w] = 0.0; }
        }
    }
    WARP_SYNC();
    cur = place_jump<0>(X, nWin);
    if (cur < 0) return false;
    FOR_LANES(lane) {
        for (int w = lane; w < nWin; w += 32) X.lkOut[w] = X.dval[cur][w];
    }
    WARP_SYNC();
    // ---- pass 2: failedPasses handed down.  Own effect: SET to the increment on a new best, else ADD the increment
    FOR_LANES(lane) {
        for (int w = lane; w < nWin; w += 32) {
            const int info = W.winInfo[w], rel = info >> 8, p = X.par[w];
            const bool scored = (info & 1) != 0;
            const PlacePath pe = p < 0 ? (rel < kPPath ? W.path[rel] : ws.gpath[rel]) : PlacePath{X.lkOut[p], 0, 0};
            const double sc = W.winScore[w];
            const bool nb = scored && sc >= X.bb[w];
            const int inc = (scored && sc < (pe.lk - pp.thresholdLogLKconsecutivePlacement)) ? 1 : 0;
            X.flags[w] = nb ? 1 : 0;
            if (nb) { X.ptr[0][w] = -1; X.ival[0][w] = inc; }
            else if (p < 0) { X.ptr[0][w] = -1; X.ival[0][w] = pe.failed + inc; }
            else { X.ptr[0][w] = p; X.ival[0][w] = inc; }
        }
    }
    WARP_SYNC();
    cur = place_jump<1>(X, nWin);
    if (cur < 0) return false;
    // ---- the stop rule of every node as if it were reached
    FOR_LANES(lane) {
        for (int w = lane; w < nWin; w += 32) {
            const int info = W.winInfo[w];
            const int failed = X.ival[cur][w];
            X.fOut[w] = failed;
            const double LK = X.lkOut[w];
            const double bestAfter = (info & 1) ? fmax(X.bb[w], W.winScore[w]) : X.bb[w];
            const bool within = LK > (bestAfter - pp.thresholdLogLK);
            const bool go = pp.strictStopRules ? (failed <= pp.allowedFails && within) : (failed <= pp.allowedFails || within);
            if (go && !(info & 2)) X.flags[w] |= 2;
        }
    }
    WARP_SYNC();
    // ---- pass 3: reached = every ancestor inside the window descends (the window's first node always is)
    FOR_LANES(lane) {
        for (int w = lane; w < nWin; w += 32) {
            const int p = X.par[w];
            X.ptr[0][w] = p;
            X.ival[0][w] = p < 0 ? 1 : ((X.flags[p] & 2) ? 1 : 0);
        }
    }
    WARP_SYNC();
    cur = place_jump<2>(X, nWin);
    if (cur < 0) return false;
    // ---- leaf comparisons (:7975-7984), only for the leaves the walk reaches
    FOR_LANES(lane) {
        for (int w = lane; w < nWin; w += 32)
            if ((W.winInfo[w] & 2) && X.ival[cur][w])
                W.winMinor[w] = dev_is_minor(m.lRef, tree_list(t, 0, W.winNode[w]), W.diffs, pp.onlyFindIdentical != 0);
    }
    WARP_SYNC();
    // ---- the first reached leaf that absorbs the sample, anomalies, and the check of the prefix-maximum assumption
    FOR_LANES(lane) {
        int cut = nWin, bad = 0, anomaly = nWin;
        for (int w = lane; w < nWin; w += 32) {
            const int info = W.winInfo[w];
            const bool ok = X.ival[cur][w] != 0;
            if (ok) {
                X.flags[w] |= 4;
                if ((info & 2) && W.winMinor[w] == 1) cut = min(cut, w);
                if (info & 4) anomaly = min(anomaly, w);
            } else if ((info & 1) && W.winScore[w] > X.bb[w]) bad = 1;
        }
        X.redD[lane] = 0.0; X.redI[lane] = bad; X.redJ[lane] = -anomaly; X.redK[lane] = cut;  // sum, max, min of the tree below
    }
    WARP_SYNC();
    place_lane_reduce(X);
    FOR_LANES(lane) {
        if (lane == 0) {
            const int cut = X.redK[0], bad = X.redI[0], anomaly = -X.redJ[0];
            X.cutW = cut;
            // a leaf that absorbs the sample ends the walk before anything else at that node; an anomaly before it sends the
            // sample to the straight-line walk; a wrong running best anywhere in the window sends the window to the serial replay
            X.committed = bad ? 0 : 1;
            if (!bad && anomaly < cut) { W.state = 2; X.committed = 2; }
        }
    }
    WARP_SYNC();
    if (X.committed == 0) return false;
    if (X.committed == 2) return true;
    const int cutW = X.cutW;
    // ---- commit: counters, the new running best and its node, the position the walk continues at (three nodes per lane, in order)
    FOR_LANES(lane) {
        int nScored = 0, nMissed = 0, nQueued = 0, lastNb = -1, target = 0;
        double mx = -INFINITY;
        for (int w = 3 * lane; w < 3 * lane + 3 && w < cutW; w++) {
            const int info = W.winInfo[w], fl = X.flags[w];
            if (!(fl & 4)) continue;
            if ((info & 2) && W.winMinor[w] == 2) nMissed++;
            if (info & 1) {
                const double sc = W.winScore[w];
                nScored++;
                mx = fmax(mx, sc);
                if (fl & 1) lastNb = w;
                if ((fl & 1) || sc > X.bb[w] - pp.thresholdLogLKoptimization) nQueued++;
            }
            target = max(target, (fl & 2) ? w + 1 : w + W.winSize[w]);
        }
        X.redD[lane] = mx;                        // max
        X.redI[lane] = nScored | (nMissed << 16);  // sum
        X.redJ[lane] = lastNb;                    // max (-1: none)
        X.redK[lane] = -target;                   // min, i.e. the largest target
        X.ival[0][lane] = nQueued;                // prefix sum: where the lane's bestNodes entries go
    }
    WARP_SYNC();
    place_lane_reduce(X);
    const int qcur = place_lane_scan(X);
    FOR_LANES(lane) {
        if (lane == 0) {
            const int nScored = X.redI[0] & 0xffff, nMissed = X.redI[0] >> 16, lastNb = X.redJ[0], target = -X.redK[0];
            const int q = X.ival[qcur][31];
            if (W.nQ + q > ws.bestCap) { W.state = 3; X.committed = 2; }
            else {
                X.qBase = W.nQ;
                W.nQ += q;
                W.phase1 += nScored;
                W.missed += nMissed;
                W.best = fmax(W.best, X.redD[0]);
                if (lastNb >= 0) { W.bestNode = W.winNode[lastNb]; X.jobNewBest = 1; }
                X.maxTarget = target;
                if (cutW < nWin) { W.state = 1; W.minorNode = W.winNode[cutW]; }
                else W.pos = pos + target;
            }
        }
    }
    WARP_SYNC();
    if (X.committed == 2) return true;
    // ---- bestNodes entries in window order, and the states the windows that follow inherit: those of the nodes whose subtree
    // reaches beyond the position the walk continues at (the ancestors of the next node; one per depth)
    const int maxTarget = X.maxTarget;
    FOR_LANES(lane) {
        int at = X.qBase + (lane > 0 ? X.ival[qcur][lane - 1] : 0);
        for (int w = 3 * lane; w < 3 * lane + 3 && w < cutW; w++) {
            const int info = W.winInfo[w], fl = X.flags[w];
            if (!(fl & 4)) continue;
            if (info & 1) {
                const double sc = W.winScore[w];
                if ((fl & 1) || sc > X.bb[w] - pp.thresholdLogLKoptimization) {
                    PlaceBest& b = ws.best[at++];
                    b.t1 = W.winNode[w]; b.score = sc; b.diffs = W.diffs;
                }
            }
            if ((fl & 2) && w + W.winSize[w] > maxTarget) {
                const int rel1 = (info >> 8) + 1;
                if (rel1 >= ws.stackCap) W.state = 3;
                else if (rel1 < kPPath) W.path[rel1] = PlacePath{X.lkOut[w], X.fOut[w], 0};
                else ws.gpath[rel1] = PlacePath{X.lkOut[w], X.fOut[w], 0};
            }
        }
    }
    WARP_SYNC();
    return true;
}

template <bool PAR>
__device__ void place_sample_warp_mat(const DevModel& m, const DevTree& t, const PlaceParams& pp, LRef in, PlaceWarpMat& X,
                                      const PlaceWarpScratch& ws, PlaceResult& r) {
    PlaceWarp& W = X.w;
    const int root = t.root;
    const double one = pp.oneMutBLen, eff = pp.effectivelyNon0BLen;
    // ---- preamble (lane 0): :7929-7971
    FOR_LANES(lane) {
        if (lane == 0) {
            W.state = 0;
            W.bestNode = root; W.phase1 = 0; W.missed = 0; W.nQ = 0;
            W.pos = 0; W.nWin = 0; W.minorNode = -1;
            W.diffs = lnull();
            X.sp = 0; X.job = 0; X.jobNewBest = 0;
            X.walk = place_whole_scratch(ws);
            const bool covered = t.order && !pp.deeperSearchForLongBranches && t.child0[root] >= 0 && in.k;
            if (!covered) W.state = 2;
            else {
                ScratchD& s = X.walk;
                LRef diffs = s_copy(s, in);
                if (diffs.k && n_mut(t, root)) diffs = s_pass(m, t, s, diffs, root, false);
                const LRef rootVect = diffs.k ? s_root_vector(m, t, s, tree_list(t, 0, root), 0.0, false) : lnull();
                if (!rootVect.k) W.state = s.err == 3 ? 3 : 2;
                else {
                    W.best = W.original = f_append(m, rootVect, diffs, true, one);
                    for (int i = 0; i < 2 && W.state == 0; i++) {
                        const int c = i == 0 ? t.child0[root] : t.child1[root];
                        LRef dc = diffs;
                        if (n_mut(t, c)) dc = s_pass(m, t, s, diffs, c, false);
                        if (!dc.k || X.sp >= ws.stackCap) { W.state = 3; break; }
                        PlaceStackE& e = ws.stack[X.sp++];
                        e.t1 = c; e.parentLK = W.best; e.failedPasses = 0; e.diffs = dc;
                    }
                }
            }
        }
    }
    WARP_SYNC();
    for (;;) {
        const int state0 = W.state, sp0 = X.sp;
        WARP_SYNC();  // every lane has read the loop condition before lane 0 pops
        if (state0 != 0 || sp0 == 0) break;
        // ---- lane 0 pops an entry: a scan job for the warp, or one node processed on the spot
        FOR_LANES(lane) {
            if (lane == 0) {
                const PlaceStackE E = ws.stack[--X.sp];
                X.cur = E;
                const int t1 = E.t1;
                X.job = (!(t.mutStart && t.mutBelow[t1]) && t.size[t1] >= kPlaceScanMin) ? 1 : 0;
                X.jobNewBest = 0;
                if (X.job) {
                    W.diffs = E.diffs;
                    W.path[0] = PlacePath{E.parentLK, E.failedPasses, 0};
                    W.pos = t.pre[t1];
                } else {
                    ScratchD& s = X.walk;
                    int failedPasses = E.failedPasses;
                    LRef d = E.diffs;
                    double LKdiff = E.parentLK;
                    bool stop = false;
                    if (t.child0[t1] < 0) {
                        const int cmp = dev_is_minor(m.lRef, tree_list(t, 0, t1), d, pp.onlyFindIdentical != 0);
                        if (cmp == 1) { W.state = 1; W.minorNode = t1; stop = true; }
                        else if (cmp == 2) W.missed++;
                    }
                    if (!stop && t.dist[t1] > eff && t.up[t1] >= 0) {
                        const LRef tot = tree_list(t, 3, t1);
                        if (!tot.k) { W.state = 2; stop = true; }
                        else {
                            LKdiff = p_append_sitewise(m, tot, d, one);
                            W.phase1++;
                            const bool nb = LKdiff >= W.best;
                            if (nb) f_shorten_inplace(m, d);  // :8065, before the entry is recorded
                            if (nb || LKdiff > W.best - pp.thresholdLogLKoptimization) {
                                if (W.nQ >= ws.bestCap) { W.state = 3; stop = true; }
                                else {
                                    PlaceBest& b = ws.best[W.nQ++];
                                    b.t1 = t1; b.score = LKdiff; b.diffs = d;
                                }
                            }
                            if (nb) { W.best = LKdiff; W.bestNode = t1; failedPasses = 0; }
                            if (LKdiff < (E.parentLK - pp.thresholdLogLKconsecutivePlacement)) failedPasses++;
                        }
                    }
                    if (!stop && t.child0[t1] >= 0) {
                        const bool within = LKdiff > (W.best - pp.thresholdLogLK);
                        const bool go = pp.strictStopRules ? (failedPasses <= pp.allowedFails && within) : (failedPasses <= pp.allowedFails || within);
                        if (go) {
                            for (int i = 0; i < 2; i++) {
                                const int c = i == 0 ? t.child0[t1] : t.child1[t1];
                                LRef dc = d;
                                if (n_mut(t, c)) dc = s_pass(m, t, s, d, c, false);
                                if (!dc.k || X.sp >= ws.stackCap) { W.state = 3; break; }
                                PlaceStackE& e = ws.stack[X.sp++];
                                e.t1 = c; e.parentLK = LKdiff; e.failedPasses = failedPasses; e.diffs = dc;
                            }
                        }
                    }
                }
            }
        }
        WARP_SYNC();
        if (!X.job) continue;
        // ---- scan job over the subtree of X.cur.t1 (same phases as place_sample_warp)
        const int jobRoot = X.cur.t1;
        const int end = t.pre[jobRoot] + t.size[jobRoot], d0 = t.depth[jobRoot];
        for (;;) {
            const int state1 = W.state, pos = W.pos;
            WARP_SYNC();
            if (state1 != 0 || pos >= end) break;
            FOR_LANES(lane) {
                for (int w = lane; w < kPWin; w += 32) {
                    const int idx = pos + w;
                    int info = 0, size = 1, node = -1;
                    if (idx < end) {
                        const ScanNode rec = t.scan[idx];
                        node = rec.node;
                        size = rec.size;
                        const bool isLong = (rec.flags & SN_LONG) != 0, tot = (rec.flags & SN_TOT) != 0;
                        info = ((isLong && tot) ? 1 : 0) | ((rec.flags & SN_INNER) ? 0 : 2) | ((isLong && !tot) ? 4 : 0) | ((rec.depth - d0) << 8);
                    }
                    W.winInfo[w] = info; W.winSize[w] = size; W.winNode[w] = node;
                    X.par[w] = idx < end ? max(t.scan[idx].parentPos - pos, -1) : -1;
                }
            }
            WARP_SYNC();
            if (PAR) place_assign_slots(X, min(kPWin, end - pos));
            else {
                FOR_LANES(lane) {
                    if (lane == 0) {
                        int nWin = min(kPWin, end - pos), k = 0;
                        for (int w = 0; w < nWin; w++) {
                            if (W.winInfo[w] & 1) {
                                if (k == 32) { nWin = w; break; }
                                W.slot[k++] = w;
                            }
                        }
                        for (; k < 32; k++) W.slot[k] = -1;
                        W.nWin = nWin;
                    }
                }
                WARP_SYNC();
            }
            FOR_LANES(lane) {
                const int w = W.slot[lane];
                if (w >= 0) W.winScore[w] = p_append_sitewise(m, tree_list(t, 3, W.winNode[w]), W.diffs, one);
            }
            WARP_SYNC();
            if (PAR && place_replay_parallel(m, t, pp, X, ws, pos)) continue;  // it compares only the leaves the walk reaches
            // the reference's loop body over the window, in order (lane 0), on the comparisons of every leaf of the window
            FOR_LANES(lane) {
                for (int w = lane; w < W.nWin; w += 32)
                    if (W.winInfo[w] & 2) W.winMinor[w] = dev_is_minor(m.lRef, tree_list(t, 0, W.winNode[w]), W.diffs, pp.onlyFindIdentical != 0);
            }
            WARP_SYNC();
            FOR_LANES(lane) {
                if (lane == 0) {
                    int j = 0;
                    const int nWin = W.nWin;
                    double best = W.best;
                    while (j < nWin) {
                        const int info = W.winInfo[j], rel = info >> 8, node = W.winNode[j];
                        const PlacePath pe = rel < kPPath ? W.path[rel] : ws.gpath[rel];
                        int failed = pe.failed;
                        double LK = pe.lk;
                        if (info & 2) {
                            const int cmp = W.winMinor[j];
                            if (cmp == 1) { W.state = 1; W.minorNode = node; break; }
                            if (cmp == 2) W.missed++;
                        }
                        if (info & 4) { W.state = 2; break; }
                        if (info & 1) {
                            LK = W.winScore[j];
                            W.phase1++;
                            const bool nb = LK >= best;
                            if (nb || LK > best - pp.thresholdLogLKoptimization) {
                                if (W.nQ >= ws.bestCap) { W.state = 3; break; }
                                PlaceBest& b = ws.best[W.nQ++];
                                b.t1 = node; b.score = LK; b.diffs = W.diffs;
                            }
                            if (nb) { best = LK; W.bestNode = node; failed = 0; X.jobNewBest = 1; }
                            if (LK < (pe.lk - pp.thresholdLogLKconsecutivePlacement)) failed++;
                        }
                        const bool within = LK > (best - pp.thresholdLogLK);
                        const bool go = pp.strictStopRules ? (failed <= pp.allowedFails && within) : (failed <= pp.allowedFails || within);
                        if (go && !(info & 2)) {
                            if (rel + 1 >= ws.stackCap) { W.state = 3; break; }
                            if (rel + 1 < kPPath) W.path[rel + 1] = PlacePath{LK, failed, 0};
                            else ws.gpath[rel + 1] = PlacePath{LK, failed, 0};
                            j += 1;
                        } else j += W.winSize[j];
                    }
                    W.best = best;
                    W.pos = pos + j;
                }
            }
            WARP_SYNC();
        }
        // the list of this job was the current one at a new best: shorten it in place (:8065), once
        FOR_LANES(lane) {
            if (lane == 0 && X.jobNewBest) {
                LRef d = X.cur.diffs;
                f_shorten_inplace(m, d);
            }
        }
        WARP_SYNC();
    }
    if (W.state != 0) {
        FOR_LANES(lane) {
            if (lane == 0) {
                if (W.state == 2) {
                    ScratchD s = place_whole_scratch(ws);
                    place_sample(m, t, pp, in, s, ws.stack, ws.stackCap, ws.best, ws.bestCap, r);
                } else {
                    r.bestNode = W.state == 1 ? W.minorNode : -1;
                    r.status = W.state;
                    r.phase1 = W.state == 1 ? W.phase1 : 0;
                    r.missedMinors = W.state == 1 ? W.missed : 0;
                    r.bestScore = W.state == 1 ? 1.0 : 0.0;
                    r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
                }
            }
        }
        WARP_SYNC();
        return;
    }
    // ---- refinement, one entry per lane, in what the walk left of the scratch
    const unsigned usedK = (X.walk.topK + 3u) & ~3u, usedP = (X.walk.topP + 1u) & ~1u;
    const unsigned totalK = 32u * ws.laneK, totalP = 32u * ws.laneP;
    const unsigned sliceK = usedK < totalK ? ((totalK - usedK) / 32u) & ~3u : 0u, sliceP = usedP < totalP ? ((totalP - usedP) / 32u) & ~1u : 0u;
    FOR_LANES(lane) {
        for (int i = lane; i < W.nQ; i += 32) {
            int rc = -1;
            if (ws.best[i].score >= W.best - pp.thresholdLogLKoptimization) {
                ScratchD s;
                s.key = ws.key + usedK + (size_t)lane * sliceK;
                s.pay = ws.pay + usedP + (size_t)lane * sliceP;
                s.ais = ws.ais + (size_t)lane * ws.laneA;
                s.capK = sliceK; s.capP = sliceP; s.capA = ws.laneA; s.topK = s.topP = 0; s.err = 0;
                rc = place_refine_entry(m, t, s, ws.best[i].t1, ws.best[i].diffs, ws.eval[i]);
            }
            ws.evalRc[i] = rc;
        }
    }
    WARP_SYNC();
    FOR_LANES(lane) {
        if (lane == 0) {
            int bestNode = W.bestNode, status = 0;
            double bestScore = W.best;
            double bTop = 0.0, bBottom = 0.0, bAppend = one;
            if (bestNode != root) {
                bTop = t.dist[bestNode] / 2;
                bBottom = t.dist[bestNode] / 2 / 2;
            }
            for (int i = 0; i < W.nQ; i++) {
                int rc = ws.evalRc[i];
                if (rc < 0) continue;
                PlaceEval e = ws.eval[i];
                if (rc == 3) {  // did not fit a lane's slice: everything above the walk's lists
                    ScratchD s;
                    s.key = ws.key + usedK; s.pay = ws.pay + usedP; s.ais = ws.ais;
                    s.capK = totalK - min(usedK, totalK); s.capP = totalP - min(usedP, totalP); s.capA = 32u * ws.laneA; s.topK = s.topP = 0; s.err = 0;
                    rc = place_refine_entry(m, t, s, ws.best[i].t1, ws.best[i].diffs, e);
                }
                if (rc > 0) { status = rc; break; }
                if (e.score >= bestScore) {
                    bestNode = ws.best[i].t1;
                    bestScore = e.score;
                    bTop = e.top; bBottom = e.bottom; bAppend = e.append;
                }
            }
            r.phase1 = W.phase1;
            r.missedMinors = W.missed;
            r.status = status;
            if (status) {
                r.bestNode = -1;
                r.bestScore = r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
            } else {
                if (bestScore == -INFINITY) bestScore = W.original;
                r.bestNode = bestNode;
                r.bestScore = bestScore;
                r.bLenTop = bTop; r.bLenBottom = bBottom; r.bLenAppend = bAppend;
            }
        }
    }
    WARP_SYNC();
}

}  // namespace maple
