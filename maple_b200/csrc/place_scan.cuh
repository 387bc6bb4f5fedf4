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

}  // namespace maple
