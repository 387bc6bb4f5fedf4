// updatePartials (MAPLEv0.7.5.4.py:5479-5815) and the default, sequential mode of traverseTreeToOptimizeBranchLengths
// (:8727-8890) on the device, with the genome lists staying in the arena.
//
// Both are sequential by definition: updatePartials is a LIFO work list in which every step re-derives the lists of one node
// from the CURRENT lists of its neighbours and stops propagating where areVectorsDifferent says nothing changed (a tolerance
// test, so the result depends on the order); the sweep re-estimates one branch at a time and calls updatePartials after every
// accepted change, so later estimates see earlier ones (Gauss-Seidel).  To give the reference's lists and lengths bit for bit
// they run here as the reference runs them: one lane, the same order, the same merges (dev_merge / dev_differ / dev_blen of
// likelihood.cuh).  What the device adds is that nothing leaves HBM: new lists are appended at the arena's tail and the four
// per-list tables are re-pointed, as `tree.probVect[node] = newList` does in the reference.
//
// Time trees, HnZ and the `testing` mode are not part of the path.
#pragma once
#include "search.cuh"

namespace maple {

struct ArenaW {  // the bound arena, writable (same memory as DevTree::key / pay / keyStart / payStart / nkeys / npay)
    uint32_t* key;
    double* pay;
    int64_t* keyStart;
    int64_t* payStart;
    int32_t* nkeys;
    int32_t* npay;
    long long* tails;  // [0] entries used in key, [1] doubles used in pay
    long long capK, capP;
};

struct UpdateState {
    const DevModel* m;
    DevTree t;       // read views of the same arrays
    ArenaW a;
    double* dist;    // == t.dist, writable
    uint8_t* dirty;
    ScratchD s;      // temporaries of one step
    int32_t* work;   // the nodeList stack: pairs (node, direction); direction 2 = change comes from the parent, 0/1 = from that child
    int workCap, nWork;
    int err;         // 0 ok, 2 the reference would raise, 3 arena / scratch / work list exhausted
};

__device__ inline void up_push(UpdateState& u, int node, int direction) {
    if (u.nWork >= u.workCap) { u.err = 3; return; }
    u.work[2 * u.nWork] = node;
    u.work[2 * u.nWork + 1] = direction;
    u.nWork++;
}

__device__ inline LRef up_list(const UpdateState& u, int fam, int node) { return tree_list(u.t, fam, node); }

// tree.<family>[node] = v  (None when v is null); with `doShorten` the stored copy is shorten()ed as the reference does after
// most assignments (:5542, :5644, :5704 ...)
__device__ inline void up_store(UpdateState& u, int fam, int node, LRef v, bool doShorten) {
    const int64_t id = (int64_t)fam * u.t.nNodes + node;
    if (!v.k) {
        u.a.keyStart[id] = -1; u.a.payStart[id] = -1; u.a.nkeys[id] = 0; u.a.npay[id] = 0;
        return;
    }
    const long long tk = u.a.tails[0], tp = u.a.tails[1];
    if (tk + v.nk + 8 > u.a.capK || tp + 6LL * v.nk + 8 > u.a.capP) { u.err = 3; return; }
    Writer w;
    w.init(u.a.key + tk, u.a.pay + tp);
    if (doShorten) dev_shorten<false>(*u.m, v.k, v.p, w);
    else {
        Cursor<false> c;
        c.init(v.k, v.p);
        for (;;) {
            double vec[4] = {0, 0, 0, 0};
            if (c.type == T_O) c.vec(vec);
            w.put(c.type, c.nl, c.flag, c.nuc, c.end, c.l0(), c.l1(), vec);
            if (c.end == u.m->lRef) break;
            c.next();
        }
    }
    u.a.keyStart[id] = tk; u.a.payStart[id] = tp; u.a.nkeys[id] = w.nk; u.a.npay[id] = w.np;
    u.a.tails[0] = tk + ((w.nk + 3) & ~3);
    u.a.tails[1] = tp + ((w.np + 1) & ~1);
}

__device__ inline LRef up_merge(UpdateState& u, LRef a, double b1, bool t1, LRef b, double b2, bool t2, bool upDown) {
    if (!a.k || !b.k) { if (!u.err) u.err = 2; return lnull(); }  // the reference would raise on a None operand
    const LRef r = s_merge(*u.m, u.s, a, b1, t1, b, b2, t2, upDown);
    if (u.s.err == 3) u.err = 3;
    u.s.err = 0;
    return r;
}

__device__ inline LRef up_pass(UpdateState& u, LRef v, int node, bool dirIsUp) {
    if (!v.k || !n_mut(u.t, node)) return v;
    const LRef r = s_pass(*u.m, u.t, u.s, v, node, dirIsUp);
    if (!r.k) u.err = 3;
    return r;
}

// updateBLen (:5385-5414)
__device__ inline void up_update_blen(UpdateState& u, int cNode, bool addToList) {
    const int node = u.t.up[cNode];
    const int cNum = (u.t.child0[node] == cNode) ? 0 : 1;
    LRef vectUp = up_list(u, cNum == 0 ? 1 : 2, node);
    vectUp = up_pass(u, vectUp, cNode, false);
    const LRef low = up_list(u, 0, cNode);
    if (!vectUp.k || !low.k) { if (!u.err) u.err = 2; return; }
    if (unsigned(vectUp.nk) + unsigned(low.nk) + 1u > u.s.capA) { u.err = 3; return; }
    u.dist[cNode] = f_blen(*u.m, vectUp, low, u.t.isTip[cNode] != 0, u.s.ais);  // python False counts as 0
    u.dirty[node] = 1;
    u.dirty[cNode] = 1;
    if (addToList) {
        up_push(u, cNode, 2);
        up_push(u, node, cNum);
    }
}

__device__ inline bool up_differ(UpdateState& u, LRef a, LRef b) {
    if (!b.k || !a.k) return true;  // (a None first operand makes the reference raise; it cannot occur on a set-up tree)
    return f_differ(*u.m, a, b);
}

// the work list must hold the entries to start from (last entry is taken first, like nodeList.pop())
__device__ void dev_update_partials(UpdateState& u) {
    const DevTree& t = u.t;
    while (u.nWork > 0 && !u.err) {
        u.nWork--;
        const int node = u.work[2 * u.nWork], direction = u.work[2 * u.nWork + 1];
        u.s.topK = u.s.topP = 0;
        bool updatedBLen = false, madeChange = false;
        u.dirty[node] = 1;
        const int parent = t.up[node];
        int childNumUp = 0;
        LRef vectUpUp = lnull();
        if (parent >= 0) {
            childNumUp = (t.child0[parent] == node) ? 0 : 1;
            vectUpUp = up_list(u, childNumUp == 0 ? 1 : 2, parent);
            vectUpUp = up_pass(u, vectUpUp, node, false);
        }
        const bool isTip = t.isTip[node] != 0;
        if (direction == 2) {  // the change comes from the parent (:5523-5661)
            if (u.dist[node] != 0.0) {
                LRef newTot = up_merge(u, vectUpUp, u.dist[node] / 2, false, up_list(u, 0, node), u.dist[node] / 2, isTip, true);
                if (u.err) return;
                if (!newTot.k) {
                    up_update_blen(u, node, false);
                    up_push(u, parent, childNumUp);
                    newTot = up_merge(u, vectUpUp, u.dist[node] / 2, false, up_list(u, 0, node), u.dist[node] / 2, isTip, true);
                    madeChange = true;
                    if (!newTot.k && !u.err) u.err = 2;
                }
                if (u.err) return;
                up_store(u, 3, node, newTot, true);
            } else up_store(u, 3, node, lnull(), false);
            if (t.child0[node] >= 0 && !u.err) {
                const int c0 = t.child0[node], c1 = t.child1[node];
                const LRef child0Vect = up_pass(u, up_list(u, 0, c0), c0, true), child1Vect = up_pass(u, up_list(u, 0, c1), c1, true);
                const bool isTip0 = t.isTip[c0] != 0, isTip1 = t.isTip[c1] != 0;
                LRef newUpRight = up_merge(u, vectUpUp, u.dist[node], false, child1Vect, u.dist[c1], isTip1, true), newUpLeft = lnull();
                if (u.err) return;
                if (!newUpRight.k) {
                    if (u.dist[node] == 0.0 && u.dist[c1] == 0.0) {
                        up_update_blen(u, node, false);
                        if (u.dist[node] == 0.0) {
                            up_update_blen(u, c1, true);
                            updatedBLen = true;
                        } else {
                            up_store(u, 3, node, up_merge(u, vectUpUp, u.dist[node] / 2, false, up_list(u, 0, node), u.dist[node] / 2, isTip, true), false);
                            newUpRight = up_merge(u, vectUpUp, u.dist[node], false, child1Vect, u.dist[c1], isTip1, true);
                            up_push(u, parent, childNumUp);
                            madeChange = true;
                        }
                    } else u.err = 2;  // "Strange: None vector from non-zero distances"
                }
                if (u.err) return;
                if (!updatedBLen) {
                    newUpLeft = up_merge(u, vectUpUp, u.dist[node], false, child0Vect, u.dist[c0], isTip0, true);
                    if (u.err) return;
                    if (!newUpLeft.k) {
                        if (u.dist[node] == 0.0 && u.dist[c0] == 0.0) {
                            up_update_blen(u, node, false);
                            if (u.dist[node] == 0.0) {
                                up_update_blen(u, c0, true);
                                updatedBLen = true;
                            } else {
                                up_store(u, 3, node, up_merge(u, vectUpUp, u.dist[node] / 2, false, up_list(u, 0, node), u.dist[node] / 2, isTip, true), false);
                                newUpRight = up_merge(u, vectUpUp, u.dist[node], false, child1Vect, u.dist[c1], isTip1, true);
                                newUpLeft = up_merge(u, vectUpUp, u.dist[node], false, child0Vect, u.dist[c0], isTip0, true);
                                up_push(u, parent, childNumUp);
                                madeChange = true;
                            }
                        } else u.err = 2;
                    }
                }
                if (u.err) return;
                if (!updatedBLen) {
                    bool upRightChanged = false, upLeftChanged = false;
                    if (madeChange || up_differ(u, up_list(u, 1, node), newUpRight)) {
                        if (!newUpRight.k) { u.err = 2; return; }
                        up_store(u, 1, node, newUpRight, true);
                        upRightChanged = true;
                    }
                    if (madeChange || up_differ(u, up_list(u, 2, node), newUpLeft)) {
                        if (!newUpLeft.k) { u.err = 2; return; }
                        up_store(u, 2, node, newUpLeft, true);
                        upLeftChanged = true;
                    }
                    if (upRightChanged) up_push(u, c0, 2);
                    if (upLeftChanged) up_push(u, c1, 2);
                }
            }
        } else {  // the change comes from child number `direction` (:5663-5814)
            const int childNum = direction, cN = childNum == 0 ? t.child0[node] : t.child1[node], oN = childNum == 0 ? t.child1[node] : t.child0[node];
            double childDist = u.dist[cN];
            const double otherChildDist = u.dist[oN];
            const LRef otherChildVect = up_pass(u, up_list(u, 0, oN), oN, true), probVectDown = up_pass(u, up_list(u, 0, cN), cN, true);
            const bool isTipC = t.isTip[cN] != 0, otherIsTip = t.isTip[oN] != 0;
            const LRef otherVectUp = up_list(u, childNum ? 1 : 2, node);
            LRef oldProbVect = lnull(), newUpVect = lnull();
            if (u.err) return;
            // lower likelihoods
            const LRef newVect = up_merge(u, otherChildVect, otherChildDist, otherIsTip, probVectDown, childDist, isTipC, false);
            if (u.err) return;
            if (!newVect.k) {
                if (childDist == 0.0 && otherChildDist == 0.0) {
                    up_update_blen(u, cN, false);
                    if (u.dist[cN] == 0.0) {
                        up_update_blen(u, oN, true);
                        updatedBLen = true;
                    } else {
                        childDist = u.dist[cN];
                        up_store(u, 0, node, up_merge(u, otherChildVect, otherChildDist, otherIsTip, probVectDown, childDist, isTipC, false), false);
                        up_push(u, cN, 2);
                        madeChange = true;
                    }
                } else u.err = 2;
            } else {
                oldProbVect = up_list(u, 0, node);  // (arena lists are never overwritten: the old one stays readable)
                up_store(u, 0, node, newVect, true);
            }
            if (u.err) return;
            // total likelihoods at the middle of the branch above
            if (!updatedBLen && u.dist[node] != 0.0 && parent >= 0 && vectUpUp.k) {
                const LRef newTot = up_merge(u, vectUpUp, u.dist[node] / 2, false, up_list(u, 0, node), u.dist[node] / 2, false, true);
                if (u.err) return;
                if (!newTot.k) {
                    up_update_blen(u, node, false);
                    up_store(u, 0, node, up_merge(u, otherChildVect, otherChildDist, otherIsTip, probVectDown, childDist, isTipC, false), false);
                    up_push(u, cN, 2);
                    up_store(u, 3, node, up_merge(u, vectUpUp, u.dist[node] / 2, false, up_list(u, 0, node), u.dist[node] / 2, false, true), false);
                    madeChange = true;
                } else up_store(u, 3, node, newTot, true);
            } else if (u.dist[node] == 0.0) up_store(u, 3, node, lnull(), false);
            if (u.err) return;
            // likelihoods passed on to the sibling
            if (!updatedBLen && otherVectUp.k) {
                if (parent >= 0) newUpVect = up_merge(u, vectUpUp, u.dist[node], false, probVectDown, childDist, isTipC, true);
                else {
                    newUpVect = s_root_vector(*u.m, t, u.s, probVectDown, childDist, isTipC);
                    if (u.s.err == 3) u.err = 3;
                }
                if (u.err) return;
                if (!newUpVect.k) {
                    if (u.dist[node] == 0.0 && childDist == 0.0 && parent >= 0) {
                        up_update_blen(u, node, false);
                        if (u.dist[node] == 0.0) {
                            up_update_blen(u, cN, true);
                            updatedBLen = true;
                        } else {
                            up_store(u, 3, node, up_merge(u, vectUpUp, u.dist[node] / 2, false, up_list(u, 0, node), u.dist[node] / 2, false, true), false);
                            up_push(u, cN, 2);
                            madeChange = true;
                            newUpVect = up_merge(u, vectUpUp, u.dist[node], false, probVectDown, childDist, isTipC, true);
                        }
                    } else u.err = 2;
                }
            }
            if (u.err) return;
            if (!updatedBLen) {
                bool upChanged = false, downChanged = false;
                if (otherVectUp.k) {
                    if (madeChange || up_differ(u, otherVectUp, newUpVect)) {
                        if (!newUpVect.k) { u.err = 2; return; }
                        upChanged = true;
                        up_store(u, childNum ? 1 : 2, node, newUpVect, true);
                    }
                }
                if (madeChange || up_differ(u, up_list(u, 0, node), oldProbVect)) downChanged = true;
                if (parent >= 0 && downChanged) up_push(u, parent, childNumUp);
                if (upChanged) up_push(u, oN, 2);
            }
        }
    }
}

// The loop of traverseTreeToOptimizeBranchLengths(tree, root, fastPass=False) below the root's children (:8815-8886); the scan of
// the root's own two branches (:8745-8814) is done by the caller with batch calls and two dev_update_partials runs.
// walk: scratch for the nodesToTraverse stack (nNodes ints).  Returns the number of updated lengths.
__device__ int dev_sweep_sequential(UpdateState& u, int32_t* walk, int walkCap) {
    const DevTree& t = u.t;
    int nWalk = 0, updates = 0;
    const int root = t.root;
    if (t.child0[root] < 0) return 0;
    for (int side = 0; side < 2; side++) {
        const int c = side == 0 ? t.child0[root] : t.child1[root];
        if (t.child0[c] >= 0) { walk[nWalk++] = t.child0[c]; walk[nWalk++] = t.child1[c]; }
    }
    while (nWalk > 0 && !u.err) {
        const int node = walk[--nWalk];
        if (u.dirty[node]) {
            const int parent = t.up[node];
            const int child = (t.child0[parent] == node) ? 0 : 1;
            u.s.topK = u.s.topP = 0;
            LRef upVect = up_list(u, child == 0 ? 1 : 2, parent);
            upVect = up_pass(u, upVect, node, false);
            const LRef low = up_list(u, 0, node);
            if (!upVect.k || !low.k) { if (!u.err) u.err = 2; break; }
            if (unsigned(upVect.nk) + unsigned(low.nk) + 1u > u.s.capA) { u.err = 3; break; }
            const double bestLength = f_blen(*u.m, upVect, low, t.isTip[node] != 0, u.s.ais);
            const double d = u.dist[node];
            bool change = false;
            if (bestLength != 0.0 || d != 0.0) change = bestLength == 0.0 || d == 0.0 || d / bestLength > 1.01 || d / bestLength < 0.99;
            if (change) {
                u.dist[node] = bestLength;
                updates++;
                up_push(u, node, 2);
                up_push(u, parent, child);
                dev_update_partials(u);
            } else u.dirty[node] = 0;
        }
        if (t.child0[node] >= 0) {
            if (nWalk + 2 > walkCap) { u.err = 3; break; }
            walk[nWalk++] = t.child0[node];
            walk[nWalk++] = t.child1[node];
        }
    }
    return updates;
}

}  // namespace maple
