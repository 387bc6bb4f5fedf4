// TEST INFRASTRUCTURE: the lane-level device code of maple_b200/csrc compiled for the host (shim/cuda_runtime.h) behind the
// C oracle's own entry-point names and signatures (oracle/maple_oracle.c), so the python wrapper of the oracle
// (oracle/oracle.py) can drive either library and the golden-vector tests run unchanged against the kernel SOURCE.
// Built by tests/hostsim/build.py with g++ -O2 -ffp-contract=off (the CUDA build uses -fmad=false for the same reason:
// a*b+c must round twice like CPython).  Never loaded by anything under maple_b200/.
#include "cuda_runtime.h"

#include "place_scan.cuh"
#include "search_fsm.cuh"
#include "update.cuh"

#include <vector>

using namespace maple;

struct OrTree {  // oracle/oracle.py: OrTree
    int32_t nNodes, root;
    const int32_t *up, *child0, *child1;
    const double* dist;
    const uint8_t* isTip;
    const int32_t *mutStart, *mut;
    const uint32_t* key;
    const double* pay;
    const int64_t *keyStart, *payStart;
    const int32_t* nkeys;
};

static DevTree dev_tree(const OrTree* t) {
    DevTree T;
    memset(&T, 0, sizeof T);
    T.nNodes = t->nNodes; T.root = t->root;
    T.up = t->up; T.child0 = t->child0; T.child1 = t->child1; T.dist = t->dist; T.isTip = t->isTip;
    T.mutStart = t->mutStart; T.mut = t->mut;
    T.key = t->key; T.pay = t->pay; T.keyStart = t->keyStart; T.payStart = t->payStart; T.nkeys = t->nkeys;
    return T;  // order == nullptr: no warp scans (they do not exist on the host)
}

static int tree_height(const OrTree* t) {
    std::vector<int> depth(t->nNodes, 0), stack{t->root};
    int h = 0;
    while (!stack.empty()) {
        const int v = stack.back();
        stack.pop_back();
        if (depth[v] > h) h = depth[v];
        for (int c : {t->child0[v], t->child1[v]})
            if (c >= 0) { depth[c] = depth[v] + 1; stack.push_back(c); }
    }
    return h;
}

extern "C" {

double or_append(const DevModel* m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC, int isTipC, double bLen) {
    return dev_append<true>(*m, kP, pP, kC, pC, isTipC != 0, bLen);
}
// the two forms the warp scans of the search kernel use per lane (search_fsm.cuh: f_append_sitewise / f_append_q4)
double hs_append_sitewise(const DevModel* m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC, int isTipC, double bLen) {
    return dev_append_sitewise<false>(*m, kP, pP, kC, pC, isTipC != 0, bLen);
}
double hs_append_q4(const DevModel* m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC, int isTipC, double bLen) {
    return dev_append_q4<false>(*m, kP, pP, kC, pC, isTipC != 0, bLen);
}

int or_merge(const DevModel* m, const uint32_t* k1, const double* p1, double bLen1, int fromTip1, const uint32_t* k2, const double* p2,
             double bLen2, int fromTip2, int flags, int numMinor1, int numMinor2, uint32_t* outKey, double* outPay, int32_t* outNk,
             int32_t* outNp, double* outLk) {
    Writer w;
    w.init(outKey, outPay);
    double lk = 0.0;
    const int st = dev_merge<true>(*m, k1, p1, bLen1, fromTip1 != 0, k2, p2, bLen2, fromTip2 != 0, flags, numMinor1, numMinor2, w, &lk);
    *outNk = st == 0 ? w.nk : 0;
    *outNp = st == 0 ? w.np : 0;
    if (outLk) *outLk = lk;
    return st;
}

void or_shorten(const DevModel* m, const uint32_t* k, const double* p, uint32_t* outKey, double* outPay, int32_t* outNk, int32_t* outNp) {
    Writer w;
    w.init(outKey, outPay);
    dev_shorten<false>(*m, k, p, w);
    *outNk = w.nk;
    *outNp = w.np;
}

int or_blen(const DevModel* m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC, int fromTipC, double* ais, double* out) {
    double v = 0.0;
    const int st = dev_blen<true>(*m, kP, pP, kC, pC, fromTipC != 0, ais, &v);
    *out = v;
    return st;
}

int or_differ(const DevModel* m, const uint32_t* k1, const double* p1, const uint32_t* k2, const double* p2) {
    return dev_differ<true>(*m, k1, p1, k2, p2) ? 1 : 0;
}

void or_pass_branch(const DevModel* m, const uint32_t* k, const double* p, const int32_t* mut, int nMut, int dirIsUp, uint32_t* outKey,
                    double* outPay, int32_t* outNk, int32_t* outNp) {
    Writer w;
    w.init(outKey, outPay);
    dev_pass_branch(m->lRef, k, p, mut, nMut, dirIsUp != 0, w);
    *outNk = w.nk;
    *outNp = w.np;
}

void or_root_vector(const DevModel* m, const uint32_t* k, const double* p, double bLen, int isFromTip, uint32_t* outKey, double* outPay,
                    int32_t* outNk, int32_t* outNp) {
    Writer w;
    w.init(outKey, outPay);
    dev_root_vector<true>(*m, k, p, bLen, isFromTip != 0, w);
    *outNk = w.nk;
    *outNp = w.np;
}

double or_prob_root(const DevModel* m, const uint32_t* k, const double* p) { return dev_prob_root<true>(*m, k, p); }

int or_is_minor(const DevModel* m, const uint32_t* k1, const double* p1, const uint32_t* k2, const double* p2, int onlyFindIdentical) {
    return dev_is_minor(m->lRef, LRef{k1, p1, 0}, LRef{k2, p2, 0}, onlyFindIdentical != 0);
}

// the straight-line search of k_spr_search (search variant 1), one node after the other
void or_search_batch(const DevModel* m, const OrTree* t, const SearchParams* sp, int64_t n, const int32_t* nodes, int64_t scratchKeys,
                     int32_t /*lazyMode: the device semantics are the pre-filled ones*/, SearchResult* out) {
    const DevTree T = dev_tree(t);
    const unsigned capK = (unsigned)scratchKeys, capP = 2 * capK + 6 * 1024, capA = 2048;
    const int stackCap = (2 * tree_height(t) + 32 + 63) & ~63;
    std::vector<uint32_t> key(capK + 64);
    std::vector<double> pay(capP + 64), ais(capA);
    std::vector<StackE> stack(stackCap);
    ScratchD s;
    s.key = key.data(); s.pay = pay.data(); s.ais = ais.data();
    s.capK = capK; s.capP = capP; s.capA = capA; s.topK = s.topP = 0; s.err = 0;
    for (int64_t i = 0; i < n; i++) search_node(*m, T, *sp, nodes[i], s, stack.data(), stackCap, out[i]);
}

// One lane of k_spr_search_fsm (the default search kernel) with the warp scans switched off (search variant 2): the per-lane
// control code (fsm_step, fsm_finish) and the co-walk service loop of the kernel, one search after the other.
static long long* g_opCounts = nullptr;  // optional [n][5]: append, merge, blen, differ co-walks and control steps per search
void hs_set_op_counts(long long* p) { g_opCounts = p; }

void hs_search_batch_fsm(const DevModel* m, const OrTree* t, const SearchParams* sp, int64_t n, const int32_t* nodes, int64_t scratchKeys,
                         SearchResult* out) {
    const DevTree T = dev_tree(t);
    const unsigned capK = (unsigned)scratchKeys, capP = 2 * capK + 6 * 1024, capA = 2048;
    const int stackCap = (2 * tree_height(t) + 32 + 63) & ~63;
    std::vector<uint32_t> key(capK + 64);
    std::vector<double> pay(capP + 64), ais(capA);
    std::vector<StackE> stack(stackCap);
    ScratchD s;
    s.key = key.data(); s.pay = pay.data(); s.ais = ais.data();
    s.capK = capK; s.capP = capP; s.capA = capA; s.topK = s.topP = 0; s.err = 0;
    for (int64_t i = 0; i < n; i++) {
        const int node = nodes[i];
        SearchResult r;
        r.placement = -1; r.bestNode = -1; r.status = 1; r.phase1 = 0;
        r.improvement = r.bestCurrentLK = r.bestScore = r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
        out[i] = r;
        if (T.up[node] < 0) continue;
        s.topK = s.topP = 0;
        s.err = 0;
        const int parent = T.up[node];
        LRef vectUp = (T.child0[parent] == node) ? tree_list(T, 1, parent) : tree_list(T, 2, parent);
        if (n_mut(T, node)) vectUp = s_pass(*m, T, s, vectUp, node, false);
        const LRef own = tree_list(T, 0, node);
        if (!vectUp.k || !own.k) { out[i].status = s.err ? s.err : 2; continue; }
        const double bestCurrentLK = f_append(*m, vectUp, own, T.isTip[node] != 0, T.dist[node]);
        r.bestCurrentLK = bestCurrentLK;
        if (!(bestCurrentLK < sp->thresholdTopologyPlacement || T.dist[node] != 0.0)) { r.status = 1; out[i] = r; continue; }  // :9674
        Fsm f;
        memset(&f, 0, sizeof f);
        f.op = OP_NONE; f.pc = 0;
        f.parent = parent;
        f.child = (T.child0[parent] == node) ? 0 : 1;
        f.bestLKdiff = bestCurrentLK;
        f.removedBLen = T.dist[node];
        f.phase1 = 0; f.rc = 0;
        s.topK = s.topP = 0;
        for (;;) {
            fsm_step(f, *m, T, *sp, s, stack.data(), stackCap, 0 /* no warp scans */);
            if (g_opCounts) { g_opCounts[5 * i + 4]++; if (f.op >= OP_APPEND && f.op <= OP_DIFFER) g_opCounts[5 * i + f.op - 1]++; }
            if (f.op == OP_DONE) break;
            if (f.op == OP_APPEND) f.resD = f_append(*m, f.a1, f.a2, f.at1 != 0, f.ab1);
            else if (f.op == OP_MERGE) {
                Writer w;
                w.init(s.key + s.topK, s.pay + s.topP);
                if (f_merge(*m, f.a1, f.ab1, f.at1 != 0, f.a2, f.ab2, f.at2 != 0, f.aflags, w) == 0) f.resL = sc_commit(s, w.nk, w.np);
                else f.resL = lnull();
            } else if (f.op == OP_BLEN) f.resD = f_blen(*m, f.a1, f.a2, f.at1 != 0, s.ais);
            else if (f.op == OP_DIFFER) f.resB = f_differ(*m, f.a1, f.a2) ? 1 : 0;
            else { f.rc = 2; break; }  // OP_SCAN cannot be requested with scanMinSize == 0
        }
        fsm_finish(f, T, *sp, node, bestCurrentLK, r);
        out[i] = r;
    }
}

// k_place_samples, one sample after the other
void or_place_batch(const DevModel* m, const OrTree* t, const PlaceParams* pp, int64_t n, const uint32_t* key, const double* pay,
                    const int64_t* keyStart, const int64_t* payStart, const int32_t* nkeys, int64_t scratchKeys, PlaceResult* out) {
    const DevTree T = dev_tree(t);
    const unsigned capK = (unsigned)scratchKeys, capP = 2 * capK + 6 * 1024, capA = 2048;
    const int stackCap = tree_height(t) + 8, bestCap = (int)(capK / 4 > 1024 ? capK / 4 : 1024);
    std::vector<uint32_t> sk(capK + 64);
    std::vector<double> spay(capP + 64), ais(capA);
    std::vector<PlaceStackE> stack(stackCap);
    std::vector<PlaceBest> best(bestCap);
    ScratchD s;
    s.key = sk.data(); s.pay = spay.data(); s.ais = ais.data();
    s.capK = capK; s.capP = capP; s.capA = capA; s.topK = s.topP = 0; s.err = 0;
    for (int64_t i = 0; i < n; i++)
        place_sample(*m, T, *pp, LRef{key + keyStart[i], pay + payStart[i], nkeys[i]}, s, stack.data(), stackCap, best.data(), bestCap, out[i]);
}

// batch drivers with the oracle's signatures (plain loops: this library is about the source, not about speed)
void or_append_batch(const DevModel* m, const uint32_t* key, const double* pay, const int64_t* keyStart, const int64_t* payStart, int64_t n,
                     const int32_t* pIdx, const int32_t* cIdx, const uint8_t* isTip, const double* bLen, double* out) {
    for (int64_t i = 0; i < n; i++) {
        const int p = pIdx[i], c = cIdx[i];
        out[i] = dev_append<true>(*m, key + keyStart[p], pay + payStart[p], key + keyStart[c], pay + payStart[c], isTip[i] != 0, bLen[i]);
    }
}

void or_merge_batch(const DevModel* m, const uint32_t* key, const double* pay, const int64_t* keyStart, const int64_t* payStart, int64_t n,
                    const int32_t* idx1, const double* bLen1, const uint8_t* tip1, const int32_t* idx2, const double* bLen2,
                    const uint8_t* tip2, const uint8_t* flags, const int32_t* numMinor1, const int32_t* numMinor2, uint32_t* outKey,
                    double* outPay, const int64_t* outKeyStart, const int64_t* outPayStart, int32_t* outNk, int32_t* outNp, double* outLk,
                    int32_t* outStatus) {
    for (int64_t i = 0; i < n; i++) {
        const int a = idx1[i], b = idx2[i];
        double lk = 0.0;
        outStatus[i] = or_merge(m, key + keyStart[a], pay + payStart[a], bLen1[i], tip1[i], key + keyStart[b], pay + payStart[b], bLen2[i],
                                tip2[i], flags[i], numMinor1 ? numMinor1[i] : 0, numMinor2 ? numMinor2[i] : 0, outKey + outKeyStart[i],
                                outPay + outPayStart[i], &outNk[i], &outNp[i], &lk);
        if (outLk) outLk[i] = lk;
    }
}

void or_blen_batch(const DevModel* m, const uint32_t* key, const double* pay, const int64_t* keyStart, const int64_t* payStart,
                   const int32_t* nkeys, int64_t n, const int32_t* pIdx, const int32_t* cIdx, const uint8_t* fromTip, double* out,
                   int32_t* outStatus) {
    std::vector<double> ais;
    for (int64_t i = 0; i < n; i++) {
        const int p = pIdx[i], c = cIdx[i];
        ais.assign(nkeys[p] + nkeys[c] + 1, 0.0);
        outStatus[i] = or_blen(m, key + keyStart[p], pay + payStart[p], key + keyStart[c], pay + payStart[c], fromTip[i], ais.data(), &out[i]);
    }
}

void or_differ_batch(const DevModel* m, const uint32_t* key, const double* pay, const int64_t* keyStart, const int64_t* payStart, int64_t n,
                     const int32_t* idx1, const int32_t* idx2, uint8_t* out) {
    for (int64_t i = 0; i < n; i++) {
        const int a = idx1[i], b = idx2[i];
        out[i] = (uint8_t)or_differ(m, key + keyStart[a], pay + payStart[a], keyStart[b] < 0 ? nullptr : key + keyStart[b],
                                    keyStart[b] < 0 ? nullptr : pay + payStart[b]);
    }
}

void or_shorten_slots(const DevModel* m, int64_t n, uint32_t* key, double* pay, const int64_t* keyStart, const int64_t* payStart, int32_t* nk,
                      int32_t* np_, const int32_t* status) {
    for (int64_t i = 0; i < n; i++) {
        if (status && status[i] != 0) continue;
        int32_t a = 0, b = 0;
        or_shorten(m, key + keyStart[i], pay + payStart[i], key + keyStart[i], pay + payStart[i], &a, &b);
        nk[i] = a;
        np_[i] = b;
    }
}

// k_place_samples_warp (placement variant 1): place_scan.cuh with the lanes of every phase run in turn.  The derived tree
// arrays are what maple_tree_bind computes on the host, the ScanNode records what k_scan_prepare computes per position.
void hs_place_batch_scan(const DevModel* m, const OrTree* t, const PlaceParams* pp, int64_t n, const uint32_t* key, const double* pay,
                         const int64_t* keyStart, const int64_t* payStart, const int32_t* nkeys, int64_t scratchKeys, const int32_t* npay,
                         int32_t matVariant /* 0: place_sample_warp, 1: place_sample_warp_mat, 2: the same with the parallel window replay */, PlaceResult* out) {
    DevTree T = dev_tree(t);
    const size_t N = (size_t)t->nNodes;
    std::vector<int32_t> order(N, -1), pre(N, -1), size(N, 1), depth(N, 0), st{t->root};
    std::vector<uint8_t> mutBelow(N, 0);
    size_t cnt = 0;
    int height = 0;
    while (!st.empty()) {
        const int v = st.back();
        st.pop_back();
        pre[v] = (int32_t)cnt;
        order[cnt++] = v;
        if (depth[v] > height) height = depth[v];
        if (t->child0[v] >= 0) {
            depth[t->child0[v]] = depth[t->child1[v]] = depth[v] + 1;
            st.push_back(t->child0[v]);
            st.push_back(t->child1[v]);
        }
    }
    for (size_t i = cnt; i-- > 1;) {
        const int v = order[i], p = t->up[v];
        size[p] += size[v];
        if (mutBelow[v] || (t->mutStart && t->mutStart[v + 1] > t->mutStart[v])) mutBelow[p] = 1;
    }
    for (size_t i = cnt; i < N; i++) order[i] = -1;
    T.order = order.data(); T.pre = pre.data(); T.size = size.data(); T.depth = depth.data(); T.mutBelow = mutBelow.data();
    T.npay = npay;
    std::vector<ScanNode> scan(N);
    T.scan = scan.data();
    for (size_t i = 0; i < N; i++) scan[i] = make_scan_node(T, pp->effectivelyNon0BLen, (int)i);
    PlaceWarpScratch ws;
    ws.laneK = 512; ws.laneP = 6 * 512; ws.laneA = 512;
    const unsigned capK = (unsigned)scratchKeys;
    ws.bestCap = (int)(capK / 4 > 1024 ? capK / 4 : 1024);
    ws.stackCap = height + 8;
    std::vector<double> payS(32 * (size_t)ws.laneP), aisS(32 * (size_t)ws.laneA), diffPay(6 * (size_t)ws.laneK);
    std::vector<uint32_t> keyS(32 * (size_t)ws.laneK + 64), diffKey(ws.laneK + 64);
    std::vector<PlaceBest> best(ws.bestCap);
    std::vector<PlaceEval> eval(ws.bestCap);
    std::vector<int> evalRc(ws.bestCap);
    std::vector<PlacePath> gpath(ws.stackCap);
    std::vector<PlaceStackE> stack(ws.stackCap);
    ws.pay = payS.data(); ws.ais = aisS.data(); ws.key = keyS.data(); ws.best = best.data(); ws.eval = eval.data(); ws.evalRc = evalRc.data();
    ws.gpath = gpath.data(); ws.stack = stack.data(); ws.diffKey = diffKey.data(); ws.diffPay = diffPay.data();
    PlaceWarpMat X;
    for (int64_t i = 0; i < n; i++) {
        memset(&X, 0, sizeof X);
        const LRef in{key + keyStart[i], pay + payStart[i], nkeys[i]};
        if (matVariant == 2) place_sample_warp_mat<true>(*m, T, *pp, in, X, ws, out[i]);
        else if (matVariant) place_sample_warp_mat<false>(*m, T, *pp, in, X, ws, out[i]);
        else place_sample_warp(*m, T, *pp, in, X.w, ws, out[i]);
    }
}

// updatePartials (mode 0: the work list `entries`, pairs (node, direction), last pair first) or the sequential branch-length sweep
// (mode 1) of update.cuh on host arrays: the arena is writable and has room behind its tails.  Returns the status (0 ok, 2 the
// reference would raise, 3 capacity); *updates = lengths changed by the sweep.
int hs_update(const DevModel* m, const OrTree* t, uint32_t* key, double* pay, int64_t* keyStart, int64_t* payStart, int32_t* nkeys, int32_t* npay,
              long long* tails, long long capK, long long capP, double* dist, uint8_t* dirty, int mode, int nEntries, const int32_t* entries,
              int32_t* updates) {
    DevTree T = dev_tree(t);
    T.key = key; T.pay = pay; T.keyStart = keyStart; T.payStart = payStart; T.nkeys = nkeys; T.dist = dist;
    std::vector<int32_t> npay_(4 * (size_t)t->nNodes + 8, 0);
    T.npay = npay ? npay : npay_.data();
    UpdateState u;
    u.m = m; u.t = T;
    u.a.key = key; u.a.pay = pay; u.a.keyStart = keyStart; u.a.payStart = payStart; u.a.nkeys = nkeys; u.a.npay = const_cast<int32_t*>(T.npay);
    u.a.tails = tails; u.a.capK = capK; u.a.capP = capP;
    u.dist = dist; u.dirty = dirty;
    const unsigned sK = 1u << 16, sP = 6u << 16, sA = 1u << 16;
    std::vector<uint32_t> sk(sK + 64);
    std::vector<double> sp(sP + 64), sa(sA);
    u.s.key = sk.data(); u.s.pay = sp.data(); u.s.ais = sa.data(); u.s.capK = sK; u.s.capP = sP; u.s.capA = sA; u.s.topK = u.s.topP = 0; u.s.err = 0;
    std::vector<int32_t> work(8 * (size_t)t->nNodes + 128), walk((size_t)t->nNodes + 8);
    u.work = work.data(); u.workCap = (int)(work.size() / 2); u.nWork = 0; u.err = 0;
    int n = 0;
    if (mode == 0) {
        for (int i = 0; i < nEntries; i++) up_push(u, entries[2 * i], entries[2 * i + 1]);
        dev_update_partials(u);
    } else n = dev_sweep_sequential(u, walk.data(), (int)walk.size());
    if (updates) *updates = n;
    return u.err;
}

int or_num_threads(void) { return 1; }

}  // extern "C"
