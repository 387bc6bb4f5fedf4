// TEST INFRASTRUCTURE: the WARP-level device code of the search kernel -- fsm_warp_loop, i.e. the whole body of
// k_spr_search_fsm with its subtree scans (warp_scan_job of search_fsm.cuh, warp_scan_job2 of scan2.cuh) -- compiled for the host
// behind shim_warp/cuda_runtime.h, which emulates the 32 lanes of a warp as coroutines that meet at the *_sync intrinsics.
// The driver below does on the host what maple_tree_bind, k_scan_prepare / k_scan_count / k_scan_build / k_scan_nsa and the
// kernel's prologue do on the device, with the same device functions.  Never loaded by anything under maple_b200/.
#include "cuda_runtime.h"

#include "search_fsm.cuh"

#include <algorithm>
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

extern "C" {

// scanForm: 0 = no warp scans, 1 = first form (warp_scan_job), 2 = second form (warp_scan_job2).
// stats: optional 32 counters (as maple_search_stats).  Returns 0, or -1 if the second form is not available for this tree.
int hw_search_batch(const DevModel* m, const OrTree* t, const SearchParams* sp, int64_t n, const int32_t* nodes, int64_t scratchKeys,
                    const int32_t* npay, int32_t scanForm, int32_t scanMinSize, int32_t lanesPerWarp, int32_t poolBytes, int32_t scanFlags,
                    int32_t bigSlots /* large scratch slots (8x) for searches that exhaust theirs, 0 = none */,
                    int32_t ownerWarps /* scan service: warps that own searches ... */, int32_t serverWarps /* ... and warps that only serve scans; 0 = no service */,
                    int32_t denseRows /* dense scoring pass: rows of the score matrix (searches that get one), 0 = off */,
                    int32_t evalSlice /* warp-wide evaluation of queued phase-2 entries: scratch entries per lane, 0 = the owning lane does it */,
                    int32_t headEnd /* head of the list handed out one per warp (BigScratch::headEnd), 0 = off */,
                    SearchResult* out, unsigned long long* stats) {
    DevTree T;
    memset(&T, 0, sizeof T);
    T.nNodes = t->nNodes; T.root = t->root;
    T.up = t->up; T.child0 = t->child0; T.child1 = t->child1; T.dist = t->dist; T.isTip = t->isTip;
    T.mutStart = t->mutStart; T.mut = t->mut;
    T.key = t->key; T.pay = t->pay; T.keyStart = t->keyStart; T.payStart = t->payStart; T.nkeys = t->nkeys;
    T.npay = npay;
    const size_t N = (size_t)t->nNodes;
    // what maple_tree_bind derives
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
    T.order = order.data(); T.pre = pre.data(); T.size = size.data(); T.depth = depth.data(); T.mutBelow = mutBelow.data();
    std::vector<ScanNode> scan(N);
    std::vector<ScanRec> recs(N);
    std::vector<uint32_t> units(N), offs(N);
    std::vector<uint4> arena;
    if (scanForm == 1) {
        T.scan = scan.data();
        for (size_t i = 0; i < N; i++) scan[i] = make_scan_node(T, sp->effectivelyNon0BLen, (int)i);
    } else if (scanForm == 2) {
        if (height >= 65535) return -1;
        uint64_t tot = 0;
        for (size_t i = 0; i < N; i++) {
            units[i] = scan_count_units(T, (int)i, m->U != 0);
            if (units[i] == ~0u) units[i] = 0;
            const uint64_t u = (units[i] & 0xffffu) + (units[i] >> 16);
            if (u == 0) { offs[i] = ~0u; continue; }
            offs[i] = (uint32_t)tot;
            tot += u;
        }
        arena.assign(tot + 4, uint4{0, 0, 0, 0});
        T.scanOff = offs.data();
        for (size_t i = 0; i < N; i++) recs[i] = scan_build_rec(*m, T, sp->effectivelyNon0BLen, (int)i, units[i], arena.data());
        for (size_t i = 0; i < N; i++) scan_fill_nsa(T, recs.data(), (int)i);
        T.scan2 = recs.data();
        T.scanArena = arena.data();
    }
    // what maple_spr_search_batch sizes
    const unsigned capK = ((unsigned)scratchKeys + 3u) & ~3u, capP = 2 * capK + 6 * 1024, capA = capK / 4 > 2048 ? capK / 4 : 2048;
    const int stackCap = (2 * height + 32 + 63) & ~63;
    if (lanesPerWarp < 1) lanesPerWarp = 1;
    if (lanesPerWarp > 32) lanesPerWarp = 32;
    const bool service = scanForm == 2 && serverWarps > 0;
    if (!service) ownerWarps = 1;
    if (ownerWarps < 1) ownerWarps = 1;
    const size_t owners = (size_t)ownerWarps * lanesPerWarp;
    std::vector<uint32_t> key(owners * capK + 64);
    std::vector<double> pay(owners * capP + 64), ais(owners * capA);
    std::vector<StackE> stack(owners * stackCap);
    const size_t fixed = scanForm == 2 ? sizeof(Scan2Smem) : sizeof(ScanSmem);
    const int nWarps = service ? ownerWarps + serverWarps : 1;
    const size_t warpSmem = (fixed + (size_t)poolBytes + 64) / 16;
    std::vector<uint4> smem(warpSmem * nWarps);
    // one warp without the service: lane k starts with entry k (fsm_warp_loop); with it everything comes from the counter
    unsigned long long counter = service ? 0ULL : (unsigned long long)(headEnd > 0 ? 1 : lanesPerWarp), bigCounter = 0;
    unsigned long long wst[kNumSearchStats] = {0};
    BigScratch big;
    memset(&big, 0, sizeof big);
    big.capK = capK * 8; big.capP = 2 * big.capK + 6 * 1024; big.capA = capA * 8;
    std::vector<uint32_t> bkey((size_t)bigSlots * big.capK + 64);
    std::vector<double> bpay((size_t)bigSlots * big.capP + 64), bais((size_t)bigSlots * big.capA + 1);
    std::vector<StackE> bstack((size_t)bigSlots * stackCap + 1);
    big.key = bkey.data(); big.pay = bpay.data(); big.ais = bais.data(); big.stack = bstack.data();
    big.nSlots = bigSlots; big.counter = &bigCounter;
    big.headEnd = (!service && headEnd > 0) ? (unsigned long long)headEnd : 0ULL;
    ScanQueue sq;
    memset(&sq, 0, sizeof sq);
    unsigned long long qctl[4] = {0, 0, 0, 0};
    unsigned cap = 1024;
    while (cap < 4 * owners) cap <<= 1;
    std::vector<unsigned long long> ring(cap, 0ULL);
    std::vector<ScanJob> jobs(owners);
    memset(jobs.data(), 0, owners * sizeof(ScanJob));
    if (service) {
        sq.head = &qctl[0]; sq.tail = &qctl[1]; sq.doneSearches = &qctl[2]; sq.ownerCounter = &qctl[3];
        sq.ring = ring.data(); sq.jobs = jobs.data(); sq.cap = cap; sq.maxOwners = (int)owners;
    }
    // dense scoring pass: what k_dense_cols, k_dense_prepare and k_dense_score do
    DenseScores ds;
    memset(&ds, 0, sizeof ds);
    std::vector<int32_t> colPos(N), rowOf((size_t)n, -1), rowEntry((size_t)(denseRows > 0 ? denseRows : 1));
    std::vector<double> rowBLen((size_t)(denseRows > 0 ? denseRows : 1)), scores;
    std::vector<uint4> cArena((size_t)(denseRows > 0 ? denseRows : 1) * kDenseCUnits);
    if (scanForm == 2 && denseRows > 0) {
        int nCols = 0;
        for (size_t i = 0; i < N; i++) {
            const bool has = (recs[i].flags & (SR_SCORED | SR_STAGED)) == (SR_SCORED | SR_STAGED);
            recs[i].col = has ? nCols : -1;
            if (has) colPos[nCols++] = (int32_t)i;
        }
        unsigned long long rowCounter = 0;
        for (int64_t i = 0; i < n; i++)
            dense_prepare_entry(*m, T, *sp, i, nodes, denseRows, &rowCounter, rowOf.data(), rowEntry.data(), cArena.data(), rowBLen.data());
        const int nRows = (int)std::min<unsigned long long>(rowCounter, (unsigned long long)denseRows);
        const long long stride = (nCols + 31) & ~31;
        scores.assign((size_t)nRows * stride + 1, NAN);
        const int densePool = poolBytes > 2048 ? poolBytes : 2048;
        std::vector<uint4> dsm((sizeof(DenseSmem) + (size_t)densePool + 64) / 16);
        DenseSmem& DW = *reinterpret_cast<DenseSmem*>(dsm.data());
        const int nTiles = (nCols + 31) / 32;
        hostwarp::run_warps(1, [&](int) {
            uint32_t parity = 0;
            for (int row0 = 0; row0 < nRows; row0 += kDenseCBlock)
                for (int tile = 0; tile < nTiles; tile++)
                    dense_score_task(*m, T, DW, densePool, parity, tile, nCols, colPos.data(), row0, std::min(nRows, row0 + kDenseCBlock), cArena.data(),
                                     rowBLen.data(), scores.data(), stride);
        });
        ds.scores = scores.data();
        ds.rowOf = rowOf.data();
        ds.stride = stride;
        if (stats) stats[30] += (unsigned long long)nRows;
    }
    // per-lane slices for the warp-wide evaluation of queued phase-2 entries (evalSlice = entries per slice, 0 = none)
    EvalScratch es;
    memset(&es, 0, sizeof es);
    std::vector<uint32_t> ek;
    std::vector<double> ep, ea;
    if (evalSlice > 0) {
        es.capK = (unsigned)evalSlice; es.capP = 6u * (unsigned)evalSlice; es.capA = (unsigned)evalSlice;
        ek.resize((size_t)nWarps * 32 * es.capK + 64);
        ep.resize((size_t)nWarps * 32 * es.capP + 64);
        ea.resize((size_t)nWarps * 32 * es.capA + 64);
        es.key = ek.data(); es.pay = ep.data(); es.ais = ea.data();
    }
    hostwarp::run_warps(nWarps, [&](int w) {
        const int lane = int(threadIdx.x & 31);
        ScanSmem& W = *reinterpret_cast<ScanSmem*>(smem.data() + warpSmem * w);
        Scan2Smem& W2 = *reinterpret_cast<Scan2Smem*>(smem.data() + warpSmem * w);
        uint32_t parity = 0;
        if (service && w >= ownerWarps) {
            scan_server_loop(*m, T, *sp, W2, poolBytes, scanFlags, parity, stats ? wst : nullptr, sq, n);
            return;
        }
        const int ownerBase = w * lanesPerWarp;
        const size_t tid = (size_t)ownerBase + (size_t)(lane < lanesPerWarp ? lane : 0);
        ScratchD s;
        s.key = key.data() + tid * capK;
        s.pay = pay.data() + tid * capP;
        s.ais = ais.data() + tid * capA;
        s.capK = capK; s.capP = capP; s.capA = capA; s.topK = 0; s.topP = 0; s.err = 0;
        StackE* stk = stack.data() + tid * (size_t)stackCap;
        if (scanForm == 2 && !service && !ds.rowOf)  // the default kernel's instantiation (scan service and dense pass compiled out)
            fsm_warp_loop<true, false>(*m, T, *sp, n, nodes, out, s, stk, stackCap, &counter, nullptr, scanMinSize, scanFlags, poolBytes,
                                stats ? wst : nullptr, nullptr, lanesPerWarp, W, W2, parity, big, 0, 1, sq, ownerBase, ds, es, (size_t)w * 32 + (size_t)lane);
        else if (scanForm == 2)
            fsm_warp_loop<true, true>(*m, T, *sp, n, nodes, out, s, stk, stackCap, &counter, nullptr, scanMinSize, scanFlags, poolBytes,
                                stats ? wst : nullptr, nullptr, lanesPerWarp, W, W2, parity, big, 0, service ? 0 : 1, sq, ownerBase, ds, es, (size_t)w * 32 + (size_t)lane);
        else
            fsm_warp_loop<false, false>(*m, T, *sp, n, nodes, out, s, stk, stackCap, &counter, nullptr, scanForm == 1 ? scanMinSize : 0, scanFlags,
                                 poolBytes, stats ? wst : nullptr, nullptr, lanesPerWarp, W, W2, parity, big, 0, 1, sq, 0, ds, es, (size_t)w * 32 + (size_t)lane);
    });
    if (stats) {
        for (int i = 0; i < kNumSearchStats; i++) stats[i] += wst[i];
        stats[31] += bigCounter;  // requests for a large slot
    }
    return 0;
}

// appendProbNode through the scan-format copies (scan2.cuh: scan_build_p, scan_build_c, scan_walk), one pair.
// Returns NaN if the removed-side copy does not fit its shared-memory slots.
double hw_scan_append(const DevModel* m, const uint32_t* kP, const double* pP, int nkP, const uint32_t* kC, const double* pC, int isTipC,
                      double bLen, int convertSlow) {
    const int capE = 1 << 16;
    std::vector<uint2> eP(nkP + 2), eC(capE);
    std::vector<double> yP(12 * (size_t)nkP + 8), yC(8 * (size_t)capE);
    uint32_t slow = 0;
    scan_build_p(*m, kP, pP, nkP, eP.data(), yP.data(), &slow);
    if (scan_build_c(*m, kC, pC, bLen, isTipC != 0, eC.data(), capE, yC.data(), 8 * capE) < 0) return NAN;
    if (convertSlow && !(m->U && isTipC))  // what a job does to its staged copy (warp_scan_job2): the factors of the O entries below the shortcut
        for (int r = 0; r < std::min(int(slow >> 30), 2); r++) {
            const uint32_t idx = (slow >> (15 * r)) & 0x7fffu;
            if (idx == 0x7fffu) continue;
            double* pay = yP.data() + ((eP[idx].y & SA_OFF) >> 3);
            pay[-1] = scan_slow_o_factor(eP[idx].x, pay, bLen);
            eP[idx].y |= SA_FAST_PS;
        }
    const double one = 1.0;
    if (__double2hiint(kMinCarryOver) != kMinCarryOverHi) return NAN;  // scan_walk's screen of the carry-over test
    double r = 0.0;
    hostwarp::run_warp([&]() { if ((threadIdx.x & 31) == 0) r = scan_walk(*m, eP.data(), yP.data(), eC.data(), yC.data(), isTipC != 0, bLen, &one, [&]() { return ScanOrig{kP, pP}; }); });
    return r;
}

}  // extern "C"
