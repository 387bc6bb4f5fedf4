// Batch kernels + C ABI (include/maple_b200.h) for the SPR-likelihood hot path, sm_100a.
//
// Thread mapping: one thread per (list, list) pair, threads of a warp take consecutive pairs of
// the batch (callers order batches so that neighbours share the child list and walk similar
// parent lists).  The 4x4 rate matrix, root frequencies and scalars are staged in shared memory
// once per CTA; per-site rate / error tables (239 kB each at lRef 29903) are read through the
// read-only path at the few informative sites only and stay L2-resident.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/maple_b200.h"
#include "likelihood.cuh"
#include "search.cuh"
#include "search_fsm.cuh"
#include "place.cuh"
#include "place_scan.cuh"
#include "scan2.cuh"
#include "update.cuh"

using namespace maple;

struct maple_ctx {
    int device = 0;
    DevModel model{};
    bool haveModel = false, haveLists = false;
    double *dSiteRates = nullptr, *dErrorRates = nullptr, *dCumRate = nullptr, *dCumErr = nullptr, *dPiLogErrCum = nullptr;
    int32_t* dCumBases = nullptr;
    const uint32_t* key = nullptr;
    const double* pay = nullptr;
    const int64_t *keyStart = nullptr, *payStart = nullptr;
    int64_t nLists = 0;
    int64_t launches = 0;
    int numSMs = 148;
    // tree bound for the search (device pointers, caller-owned)
    DevTree tree{};
    bool haveTree = false;
    int32_t* treeDerived = nullptr;  // order | pre | size | depth (int32[nNodes] each) then mutBelow (uint8[nNodes]); owned
    int treeHeight = 0;
    bool treeHasMut = false;  // some reachable node carries MAT mutations
    unsigned long long* searchStats = nullptr;  // device counters of the search kernel (maple_search_stats)
    bool statsOn = false;
    int lanesPerWarp = 0;               // searches per warp (1..32); 0 = chosen per launch from the number of searches
    int headSearches = 0;               // the first so many entries of a batch are run one per warp (maple_ctx_set_head_searches)
    int criticalSearches = 0;           // the first so many entries of a batch run on an SM of their own each (maple_ctx_set_critical_searches)
    cudaStream_t criticalStream = nullptr;
    cudaEvent_t criticalEvA = nullptr, criticalEvB = nullptr;
    unsigned long long* criticalCounter = nullptr;
    size_t criticalHogSmem = 0;
    bool scanReplaySequential = false;  // A/B: node-by-node window replay instead of the pointer-jumping one
    bool scanAppendSitewise = true;   // A/B: break-at-every-site appendProbNode in the scans instead of the queued one
    int fsmMinBlocks = 7;            // __launch_bounds__ minimum CTAs per SM of the state-machine kernel (7 -> 128 registers, 6 -> 168)
    int scanMinSize = 8;             // subtrees of at least this many nodes are scanned by the whole warp (0 = never)
    // second form of the scans (scan2.cuh): per-position sizes / offsets of the scan-format lists (computed at maple_tree_bind), the
    // arena that holds them and the per-position records (both rewritten before every search launch); owned
    bool scan2Ok = false, scanOld = false, scanOldEnv = false;  // scanOld: A/B switch back to the first form (MAPLE_SCAN_OLD=1 or search variant 4)
    uint32_t *scanUnits = nullptr, *scanOffsets = nullptr;
    uint4* scanArena = nullptr;
    ScanRec* scanRecs = nullptr;
    uint32_t scanMaxUnits = 0;    // largest scan-format list, 16-byte units
    size_t scanNumLists = 0;      // lists that have a scan-format copy
    bool scanAllStaged = false;   // every probVectTotUp list has a scan-format copy
    // scan service (scan2.cuh: ScanQueue): SMs whose CTAs own the searches; the CTAs of all other SMs only serve subtree scans.
    // -1 = chosen per launch from the stop rules, 0 = off (every warp scans for its own lanes: the default -- measured on the
    // 100 000-sequence rounds the service is 10-15 % slower in the deep round and 2.5x slower in the fast one)
    int fsmSMs = 0;
    void* queueMem = nullptr;
    size_t queueBytes = 0;
    // dense scoring pass (scan2.cuh: DenseScores): -1 = whenever it applies and the stop rules are the non-strict ones, 0 = never
    // (the default: at 100 000 sequences the pass takes 1.86 s and the searches that read it 0.9 s, against 2.2 s for searches that
    // score in place -- see DESIGN.md), 1 = whenever it applies.  The score matrix takes at most denseBudget bytes of HBM;
    // searches beyond it scan the usual way.
    int denseMode = 0;
    size_t denseBudget = (size_t)64 << 30;
    void* evalMem = nullptr;      // per-lane scratch slices of warp_eval_queue (search_fsm.cuh)
    size_t evalBytes = 0;
    bool evalQueueWarp = true;    // MAPLE_EVALQ=0: the owning lane evaluates its queued phase-2 entries itself (A/B)
    void* updateMem = nullptr;    // scratch of k_update_partials
    size_t updateBytes = 0;
    void* denseMem = nullptr;     // scores | removed-list copies | row tables | column table | counters
    size_t denseBytes = 0;
    // per-thread scratch of the search kernel (owned by the context)
    void* searchScratch = nullptr;
    size_t searchScratchBytes = 0;
    unsigned long long* searchCounter = nullptr;
    void* placeScratch = nullptr;  // per-thread scratch of the sample-placement kernel
    size_t placeScratchBytes = 0;
    void* retryScratch = nullptr;  // node list + scratch of the on-device retry of overflowed searches
    size_t retryScratchBytes = 0;
    unsigned long long* retryCounters = nullptr;  // [0] overflowed searches, [1] work counter of the retry launch
    int placeVariant = 3;   // 0 = one sample per thread (place.cuh), 1 = one sample per warp with windowed scans (place_scan.cuh), 2 = + MAT trees, 3 (default) = 2 with the parallel window replay
    int searchVariant = 0;  // 0 = state machine + warp-cooperative subtree scans (default), 1 = straight-line kernel, 2 = state machine only, 3 = scans with the queued-site appendProbNode and the node-by-node replay
    // device staging for the host-buffer entry point
    void* devStage = nullptr;
    size_t devStageBytes = 0;
    std::string err;
};

static thread_local std::string g_err;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            return MAPLE_E_CUDA;                                                          \
        }                                                                                 \
    } while (0)

constexpr int kThreads = 128;

// MAPLE_TIME_KERNELS=1: device time of the phases of maple_spr_search_batch, printed to stderr (synchronises; diagnostics only)
struct PhaseTimer {
    bool on;
    cudaStream_t st;
    std::vector<std::pair<const char*, cudaEvent_t>> ev;
    PhaseTimer(cudaStream_t s) : on(getenv("MAPLE_TIME_KERNELS") != nullptr), st(s) { mark("start"); }
    void mark(const char* name) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back({name, e});
    }
    void report() {
        if (!on) return;
        cudaEventSynchronize(ev.back().second);
        fprintf(stderr, "[maple_b200]");
        for (size_t i = 1; i < ev.size(); i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
            fprintf(stderr, " %s %.2f ms |", ev[i].first, ms);
        }
        fprintf(stderr, "\n");
        for (auto& e : ev) cudaEventDestroy(e.second);
    }
};

struct Arena {
    const uint32_t* key;
    const double* pay;
    const int64_t* keyStart;
    const int64_t* payStart;
};

__device__ __forceinline__ void stage_model(DevModel& sm, const DevModel& gm) {
    // 32-bit words of the parameter struct -> shared memory
    const int nWords = sizeof(DevModel) / 4;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&gm);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sm);
    for (int i = threadIdx.x; i < nWords; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads) k_append(const __grid_constant__ DevModel gm, Arena A, int64_t n, const int32_t* __restrict__ pIdx,
                                                     const int32_t* __restrict__ cIdx, const uint8_t* __restrict__ isTip,
                                                     const double* __restrict__ bLen, double* __restrict__ out) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = pIdx[i], c = cIdx[i];
        const int64_t kp = __ldg(A.keyStart + p), kc = __ldg(A.keyStart + c);
        out[i] = dev_append<true>(sm, A.key + kp, A.pay + __ldg(A.payStart + p), A.key + kc, A.pay + __ldg(A.payStart + c),
                                  isTip[i] != 0, bLen[i]);
    }
}

__global__ void __launch_bounds__(kThreads) k_merge(const __grid_constant__ DevModel gm, Arena A, int64_t n, const int32_t* __restrict__ idx1,
                                                    const double* __restrict__ bLen1, const uint8_t* __restrict__ tip1,
                                                    const int32_t* __restrict__ idx2, const double* __restrict__ bLen2,
                                                    const uint8_t* __restrict__ tip2, const uint8_t* __restrict__ flags,
                                                    const int32_t* __restrict__ nm1, const int32_t* __restrict__ nm2, uint32_t* outKey,
                                                    double* outPay, const int64_t* __restrict__ outKeyStart,
                                                    const int64_t* __restrict__ outPayStart, int32_t* __restrict__ outNk,
                                                    int32_t* __restrict__ outNp, double* __restrict__ outLk, int32_t* __restrict__ outStatus,
                                                    int doShorten) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = idx1[i], b = idx2[i];
        Writer w;
        w.init(outKey + outKeyStart[i], outPay + outPayStart[i]);
        double lk = 0.0;
        int st = dev_merge<true>(sm, A.key + __ldg(A.keyStart + a), A.pay + __ldg(A.payStart + a), bLen1[i], tip1[i] != 0,
                                 A.key + __ldg(A.keyStart + b), A.pay + __ldg(A.payStart + b), bLen2[i], tip2[i] != 0, flags[i],
                                 nm1 ? nm1[i] : 0, nm2 ? nm2[i] : 0, w, &lk);
        if (st == 0 && doShorten) {
            Writer w2;
            w2.init(w.key, w.pay);
            dev_shorten<false>(sm, w.key, w.pay, w2);  // in place: the writer never overtakes the reader
            w = w2;
        }
        outNk[i] = st == 0 ? w.nk : 0;
        outNp[i] = st == 0 ? w.np : 0;
        outStatus[i] = st;
        if (outLk) outLk[i] = lk;
    }
}

__global__ void __launch_bounds__(kThreads) k_blen(const __grid_constant__ DevModel gm, Arena A, int64_t n, const int32_t* __restrict__ pIdx,
                                                   const int32_t* __restrict__ cIdx, const uint8_t* __restrict__ fromTip, double* scratch,
                                                   const int64_t* __restrict__ scratchStart, double* __restrict__ out,
                                                   int32_t* __restrict__ outStatus) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = pIdx[i], c = cIdx[i];
        double v = 0.0;
        outStatus[i] = dev_blen<true>(sm, A.key + __ldg(A.keyStart + p), A.pay + __ldg(A.payStart + p), A.key + __ldg(A.keyStart + c),
                                      A.pay + __ldg(A.payStart + c), fromTip[i] != 0, scratch + scratchStart[i], &v);
        out[i] = v;
    }
}

__global__ void __launch_bounds__(kThreads) k_differ(const __grid_constant__ DevModel gm, Arena A, int64_t n, const int32_t* __restrict__ idx1,
                                                     const int32_t* __restrict__ idx2, uint8_t* __restrict__ out) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = idx1[i], b = idx2[i];
        const int64_t kb = __ldg(A.keyStart + b);
        out[i] = dev_differ<true>(sm, A.key + __ldg(A.keyStart + a), A.pay + __ldg(A.payStart + a), kb < 0 ? nullptr : A.key + kb,
                                  kb < 0 ? nullptr : A.pay + __ldg(A.payStart + b))
                     ? 1
                     : 0;
    }
}

__global__ void __launch_bounds__(kThreads) k_root_vector(const __grid_constant__ DevModel gm, Arena A, int64_t n, const int32_t* __restrict__ idx,
                                                          const double* __restrict__ bLen, const uint8_t* __restrict__ fromTip, uint32_t* outKey,
                                                          double* outPay, const int64_t* __restrict__ outKeyStart,
                                                          const int64_t* __restrict__ outPayStart, int32_t* __restrict__ outNk,
                                                          int32_t* __restrict__ outNp, int doShorten) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = idx[i];
        Writer w;
        w.init(outKey + outKeyStart[i], outPay + outPayStart[i]);
        dev_root_vector<true>(sm, A.key + __ldg(A.keyStart + a), A.pay + __ldg(A.payStart + a), bLen[i], fromTip[i] != 0, w);
        if (doShorten) {
            Writer w2;
            w2.init(w.key, w.pay);
            dev_shorten<false>(sm, w.key, w.pay, w2);
            w = w2;
        }
        outNk[i] = w.nk;
        outNp[i] = w.np;
    }
}

__global__ void __launch_bounds__(kThreads) k_prob_root(const __grid_constant__ DevModel gm, Arena A, int64_t n, const int32_t* __restrict__ idx,
                                                        double* __restrict__ out) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = idx[i];
        out[i] = dev_prob_root<true>(sm, A.key + __ldg(A.keyStart + a), A.pay + __ldg(A.payStart + a));
    }
}

// passGenomeListThroughBranch (:3749-3877): list idx[i] through the MAT mutations of node mutNode[i]
__global__ void __launch_bounds__(kThreads) k_pass_branch(int lRef, Arena A, int64_t n, const int32_t* __restrict__ idx,
                                                          const int32_t* __restrict__ mutNode, const uint8_t* __restrict__ dirIsUp,
                                                          const int32_t* __restrict__ mutStart, const int32_t* __restrict__ mut, uint32_t* outKey,
                                                          double* outPay, const int64_t* __restrict__ outKeyStart,
                                                          const int64_t* __restrict__ outPayStart, int32_t* __restrict__ outNk,
                                                          int32_t* __restrict__ outNp) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = idx[i], nd = mutNode[i];
        Writer w;
        w.init(outKey + outKeyStart[i], outPay + outPayStart[i]);
        const int m0 = mutStart[nd], nm = mutStart[nd + 1] - m0;
        dev_pass_branch(lRef, A.key + A.keyStart[a], A.pay + A.payStart[a], mut + 3 * (size_t)m0, nm, dirIsUp[i] != 0, w);
        outNk[i] = w.nk;
        outNp[i] = w.np;
    }
}

// gather-copy of whole lists between arenas: one warp per list, coalesced
__global__ void __launch_bounds__(256) k_lists_copy(int64_t n, const uint32_t* __restrict__ srcKey, const double* __restrict__ srcPay,
                                                    const int64_t* __restrict__ srcKeyStart, const int64_t* __restrict__ srcPayStart,
                                                    const int32_t* __restrict__ nk, const int32_t* __restrict__ np, uint32_t* __restrict__ dstKey,
                                                    double* __restrict__ dstPay, const int64_t* __restrict__ dstKeyStart,
                                                    const int64_t* __restrict__ dstPayStart) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nWarps) {
        const int64_t sk = srcKeyStart[i], dk = dstKeyStart[i];
        if (sk < 0 || dk < 0) continue;
        const int k = nk[i], p = np[i];
        const uint32_t* s1 = srcKey + sk;
        uint32_t* d1 = dstKey + dk;
        for (int j = lane; j < k; j += 32) d1[j] = s1[j];
        const double* s2 = srcPay + srcPayStart[i];
        double* d2 = dstPay + dstPayStart[i];
        for (int j = lane; j < p; j += 32) d2[j] = s2[j];
    }
}

constexpr int kSearchThreads = 64;

// searches whose per-lane scratch ran out (status 3) are collected for a second launch with a few lanes and 8x the scratch
__global__ void __launch_bounds__(256) k_collect_overflow(int64_t n, const SearchResult* __restrict__ out, const int32_t* __restrict__ nodes,
                                                          int32_t* __restrict__ retryNodes, int32_t* __restrict__ retryIdx,
                                                          unsigned long long* retryCount, int cap) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || out[i].status != 3) return;
    const unsigned long long at = atomicAdd(retryCount, 1ULL);
    if (at < (unsigned long long)cap) { retryNodes[at] = nodes[i]; retryIdx[at] = (int32_t)i; }
}

// one thread per pre-order position: the ScanNode records of the bound tree and arena for this launch's effectivelyNon0BLen
__global__ void __launch_bounds__(256) k_scan_prepare(const __grid_constant__ DevTree T, double eff, ScanNode* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T.nNodes) return;
    out[i] = make_scan_node(T, eff, i);
}

// ---- scan-format copies of the probVectTotUp lists (scan2.cuh), one thread per pre-order position
// sizes, at maple_tree_bind: 16-byte units of the copy of the list at each position (0 = no list / too large to stage)
__global__ void __launch_bounds__(256) k_scan_count(const __grid_constant__ DevTree T, uint32_t* __restrict__ units, bool U) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T.nNodes) units[i] = scan_count_units(T, i, U);
}

// per launch: the copies themselves (they hold Q * siteRate, so they follow the model) and the records
__global__ void __launch_bounds__(128) k_scan_build(const __grid_constant__ DevModel gm, const __grid_constant__ DevTree T, double eff,
                                                    const uint32_t* __restrict__ units, uint4* __restrict__ arena, ScanRec* __restrict__ recs) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T.nNodes) recs[i] = scan_build_rec(sm, T, eff, i, units[i], arena);
}

// nearest scored proper ancestor of every position (what a node inherits only changes at scored nodes)
__global__ void __launch_bounds__(256) k_scan_nsa(const __grid_constant__ DevTree T, ScanRec* __restrict__ recs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T.nNodes) scan_fill_nsa(T, recs, i);
}

// updatePartials / the sequential branch-length sweep (update.cuh): one lane, the reference's order
__global__ void __launch_bounds__(32) k_update_partials(const __grid_constant__ DevModel gm, const __grid_constant__ DevTree T, ArenaW a, double* dist,
                                                        uint8_t* dirty, uint32_t* scrKey, double* scrPay, double* scrAis, unsigned capK, unsigned capP,
                                                        unsigned capA, int32_t* work, int workCap, int32_t* walk, int walkCap,
                                                        const int32_t* __restrict__ entries, int nEntries, int mode, int32_t* result) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    if (threadIdx.x != 0) return;
    UpdateState u;
    u.m = &sm; u.t = T; u.a = a; u.dist = dist; u.dirty = dirty;
    u.s.key = scrKey; u.s.pay = scrPay; u.s.ais = scrAis; u.s.capK = capK; u.s.capP = capP; u.s.capA = capA; u.s.topK = u.s.topP = 0; u.s.err = 0;
    u.work = work; u.workCap = workCap; u.nWork = 0; u.err = 0;
    int updates = 0;
    if (mode == 0) {
        for (int i = 0; i < nEntries; i++) up_push(u, entries[2 * i], entries[2 * i + 1]);
        dev_update_partials(u);
    } else updates = dev_sweep_sequential(u, walk, walkCap);
    result[0] = u.err;
    result[1] = updates;
}

// ---- dense scoring pass (scan2.cuh)
// columns: compact index of the positions whose node is scored and has a scan-format copy.  One block; positions in chunks.
__global__ void __launch_bounds__(1024) k_dense_cols(int nNodes, ScanRec* __restrict__ recs, int32_t* __restrict__ colPos, int32_t* __restrict__ nCols) {
    __shared__ int warpSum[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int start = 0; start < nNodes; start += 1024) {
        const int i = start + threadIdx.x;
        const bool has = i < nNodes && (recs[i].flags & (SR_SCORED | SR_STAGED)) == (SR_SCORED | SR_STAGED);
        const unsigned b = __ballot_sync(0xffffffffu, has);
        if (lane == 0) warpSum[wid] = __popc(b);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < wid; w++) before += warpSum[w];
        int total = 0;
        for (int w = 0; w < 32; w++) total += warpSum[w];
        const int col = base + before + __popc(b & ((1u << lane) - 1u));
        if (i < nNodes) recs[i].col = has ? col : -1;
        if (has) colPos[col] = i;
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *nCols = base;
}

__global__ void __launch_bounds__(128) k_dense_prepare(const __grid_constant__ DevModel gm, const __grid_constant__ DevTree T,
                                                       const __grid_constant__ SearchParams sp, int64_t n, const int32_t* __restrict__ nodes,
                                                       int maxRows, unsigned long long* rowCounter, int32_t* __restrict__ rowOf,
                                                       int32_t* __restrict__ rowEntry, uint4* __restrict__ cArena, double* __restrict__ rowBLen) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dense_prepare_entry(sm, T, sp, i, nodes, maxRows, rowCounter, rowOf, rowEntry, cArena, rowBLen);
}

constexpr int kDenseThreads = 128;
// persistent: every warp pulls (tile of 32 columns, block of kDenseCBlock rows) tasks; the numbers of rows and columns live on the device
__global__ void __launch_bounds__(kDenseThreads, 3) k_dense_score(const __grid_constant__ DevModel gm, const __grid_constant__ DevTree T, int poolBytes,
                                                                  const int32_t* __restrict__ nColsDev, const int32_t* __restrict__ colPos,
                                                                  const unsigned long long* __restrict__ rowCounter, int maxRows,
                                                                  const uint4* __restrict__ cArena, const double* __restrict__ rowBLen,
                                                                  double* __restrict__ scores, long long stride, unsigned long long* taskCounter) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    extern __shared__ uint4 dynSmem[];
    const int warpSmem = int(sizeof(DenseSmem) - sizeof(uint4)) + poolBytes;
    DenseSmem& W = *reinterpret_cast<DenseSmem*>(reinterpret_cast<char*>(dynSmem) + (threadIdx.x >> 5) * warpSmem);
    uint32_t parity = 0;
#ifdef __CUDA_ARCH__
    if ((threadIdx.x & 31) == 0) mbar_init(&W.mbar);
#endif
    __syncwarp();
    const int nCols = *nColsDev;
    const int nRows = (int)min((unsigned long long)maxRows, *rowCounter);
    const long long nTiles = (nCols + 31) / 32, nBlocks = (nRows + kDenseCBlock - 1) / kDenseCBlock;
    const long long nTasks = nTiles * nBlocks;
    for (;;) {
        long long task = 0;
        if ((threadIdx.x & 31) == 0) task = (long long)atomicAdd(taskCounter, 1ULL);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= nTasks) break;
        // consecutive tasks share the block of removed lists (L2) and take neighbouring tiles
        const int blk = int(task / nTiles), tile = int(task % nTiles);
        const int row0 = blk * kDenseCBlock, row1 = min(nRows, row0 + kDenseCBlock);
        dense_score_task(sm, T, W, poolBytes, parity, tile, nCols, colPos, row0, row1, cArena, rowBLen, scores, stride);
    }
}

// one SPR search per thread; threads pull the next pruned node from a global counter (searches differ ~10x in length)
__global__ void __launch_bounds__(kSearchThreads) k_spr_search(const __grid_constant__ DevModel gm, const __grid_constant__ DevTree T,
                                                               const __grid_constant__ SearchParams sp, int64_t n,
                                                               const int32_t* __restrict__ nodes, SearchResult* __restrict__ out, uint32_t* scrKey,
                                                               double* scrPay, double* scrAis, StackE* scrStack, unsigned capK, unsigned capP,
                                                               unsigned capA, int stackCap, unsigned long long* counter, long long* outCycles) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    ScratchD s;
    s.key = scrKey + tid * capK;
    s.pay = scrPay + tid * capP;
    s.ais = scrAis + tid * capA;
    s.capK = capK; s.capP = capP; s.capA = capA; s.topK = 0; s.topP = 0; s.err = 0;
    StackE* stack = scrStack + tid * (size_t)stackCap;
    for (;;) {
        const unsigned long long i = atomicAdd(counter, 1ULL);
        if (i >= (unsigned long long)n) break;
        SearchResult r;
        const long long c0 = clock64();
        search_node(sm, T, sp, nodes[i], s, stack, stackCap, r);
        out[i] = r;
        if (outCycles) outCycles[i] = clock64() - c0;
    }
}

// Holds the stream until `want` CTAs of a launch on another stream have started (BigScratch::started), or ~0.1 s have passed:
// the launch that follows then finds those CTAs resident and is scheduled around them.  (Without it the launch that follows
// may take every SM first, and CTAs that need a whole SM each then wait for it to end.  The first such launch of a process has
// been seen to take well over 2 ms to start.)
__global__ void k_wait_started(const unsigned long long* started, unsigned long long want) {
    const long long t0 = clock64();
    while (ld_volatile_u64(started) < want && clock64() - t0 < 200000000LL) spin_pause(500);
}

// The same searches as k_spr_search, one per lane, but as resumable state machines (search_fsm.cuh): every loop
// iteration each lane advances its control code to the next co-walk request, then the warp runs each kind of
// co-walk once for all lanes that requested it.
template <int MINB, bool SCAN2, bool EXTRAS>
__global__ void __launch_bounds__(kSearchThreads, MINB) k_spr_search_fsm(const __grid_constant__ DevModel gm, const __grid_constant__ DevTree T,
                                                                   const __grid_constant__ SearchParams sp, int64_t n,
                                                                   const int32_t* __restrict__ nodes, SearchResult* __restrict__ out,
                                                                   uint32_t* scrKey, double* scrPay, double* scrAis, StackE* scrStack,
                                                                   unsigned capK, unsigned capP, unsigned capA, int stackCap,
                                                                   unsigned long long* counter, long long* outCycles, int scanMinSize,
                                                                   int scanFlags, int poolBytes, unsigned long long* stats,
                                                                   const unsigned long long* nDev, const int32_t* outIndex, int lanesPerWarp,
                                                                   const __grid_constant__ BigScratch big, const __grid_constant__ ScanQueue sq,
                                                                   int fsmSMs, const __grid_constant__ DenseScores ds, const __grid_constant__ EvalScratch es) {
    __shared__ DevModel sm;
    __shared__ unsigned long long wst[kSearchThreads / 32][kNumSearchStats];
    unsigned long long* st = nullptr;  // per-warp counters (lane 0 adds), flushed to `stats` at the end
    if (stats) {
        for (int i = threadIdx.x; i < (kSearchThreads / 32) * kNumSearchStats; i += blockDim.x) (&wst[0][0])[i] = 0ULL;
        st = wst[threadIdx.x >> 5];
        __syncthreads();
    }
    extern __shared__ uint4 dynSmem[];
    const int warpSmem = int((SCAN2 ? sizeof(Scan2Smem) : sizeof(ScanSmem)) - sizeof(uint4)) + poolBytes;
    char* const warpBase = reinterpret_cast<char*>(dynSmem) + (threadIdx.x >> 5) * warpSmem;
    ScanSmem& W = *reinterpret_cast<ScanSmem*>(warpBase);
    Scan2Smem& W2 = *reinterpret_cast<Scan2Smem*>(warpBase);
    uint32_t mbarParity = 0;
    if (SCAN2) {
#ifdef __CUDA_ARCH__
        if ((threadIdx.x & 31) == 0) mbar_init(&W2.mbar);
#endif
        __syncwarp();
    }
    (void)W;
    if (big.started && threadIdx.x == 0) atomicAdd(big.started, 1ULL);
    if (nDev) n = (int64_t)min((unsigned long long)n, *nDev);  // retry launch: the list length lives on the device
    stage_model(sm, gm);
    const int lane_ = int(threadIdx.x & 31);
    size_t tid;
    int ownerBase = 0;
    if (EXTRAS && SCAN2 && sq.cap != 0) {
        // Scan service: the CTAs on the first fsmSMs SMs (and CTA 0, so that somebody owns the searches wherever the CTAs land) run
        // the searches' state machines, 32 to a warp, and post their subtree scans; every other CTA only serves scans.  The split
        // is by SM so that an SM's instruction cache holds one of the two code paths, not both.
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        bool owns = int(smid) < fsmSMs || blockIdx.x == 0;
        if (owns) {
            if (lane_ == 0) ownerBase = int(atomicAdd(sq.ownerCounter, (unsigned long long)lanesPerWarp));
            ownerBase = __shfl_sync(0xffffffffu, ownerBase, 0);
            if (ownerBase + lanesPerWarp > sq.maxOwners) owns = false;  // more owning warps than the host sized scratch for
        }
        if (!owns) {
            scan_server_loop(sm, T, sp, W2, poolBytes, scanFlags, mbarParity, st, sq, n);
            goto flush;
        }
        tid = (size_t)ownerBase + (lane_ < lanesPerWarp ? lane_ : 0);
    } else {
        // scratch is laid out for the lanes that own searches only
        tid = ((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5) * (size_t)lanesPerWarp + (lane_ < lanesPerWarp ? lane_ : 0);
    }
    {
    ScratchD s;
    s.key = scrKey + tid * capK;
    s.pay = scrPay + tid * capP;
    s.ais = scrAis + tid * capA;
    s.capK = capK; s.capP = capP; s.capA = capA; s.topK = 0; s.topP = 0; s.err = 0;
    StackE* stack = scrStack + tid * (size_t)stackCap;
    fsm_warp_loop<SCAN2, EXTRAS>(sm, T, sp, n, nodes, out, s, stack, stackCap, counter, outCycles, scanMinSize, scanFlags, poolBytes, st, outIndex,
                         lanesPerWarp, W, W2, mbarParity, big, int((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5),
                         (nDev || sq.cap != 0) ? 0 : int(gridDim.x * (blockDim.x >> 5)), sq, ownerBase, ds, es,
                         blockIdx.x * (size_t)blockDim.x + threadIdx.x);
    }
flush:
    if (stats) {
        __syncthreads();
        for (int i = threadIdx.x; i < kNumSearchStats; i += blockDim.x) {
            unsigned long long v = 0;
            for (int w = 0; w < kSearchThreads / 32; w++) v += wst[w][i];
            if (v) atomicAdd(stats + i, v);
        }
    }
}

// one new-sample placement per thread (place.cuh); threads pull samples from a global counter
__global__ void __launch_bounds__(kSearchThreads) k_place_samples(const __grid_constant__ DevModel gm, const __grid_constant__ DevTree T,
                                                                  const __grid_constant__ PlaceParams pp, int64_t n,
                                                                  const int32_t* __restrict__ sampleLists, PlaceResult* __restrict__ out,
                                                                  char* scratch, size_t laneBytes, unsigned capK, unsigned capP, unsigned capA,
                                                                  int stackCap, int bestCap, unsigned long long* counter) {
    __shared__ DevModel sm;
    stage_model(sm, gm);
    char* base = scratch + (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * laneBytes;
    ScratchD s;
    s.pay = reinterpret_cast<double*>(base);
    s.ais = s.pay + capP;
    PlaceBest* best = reinterpret_cast<PlaceBest*>(s.ais + capA);
    PlaceStackE* stack = reinterpret_cast<PlaceStackE*>(best + bestCap);
    s.key = reinterpret_cast<uint32_t*>(stack + stackCap);
    s.capK = capK; s.capP = capP; s.capA = capA; s.topK = 0; s.topP = 0; s.err = 0;
    for (;;) {
        const unsigned long long i = atomicAdd(counter, 1ULL);
        if (i >= (unsigned long long)n) break;
        const int64_t id = sampleLists[i];
        PlaceResult r;
        const int64_t ks = T.keyStart[id];
        if (ks < 0) { r.bestNode = -1; r.status = 2; r.phase1 = r.missedMinors = 0; r.bestScore = r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0; }
        else place_sample(sm, T, pp, LRef{T.key + ks, T.pay + T.payStart[id], T.nkeys[id]}, s, stack, stackCap, best, bestCap, r);
        out[i] = r;
    }
}

// one new-sample placement per WARP (place_scan.cuh); warps pull samples from a global counter
constexpr int kPlaceWarpThreads = 128;
__global__ void __launch_bounds__(kPlaceWarpThreads) k_place_samples_warp(const __grid_constant__ DevModel gm, const __grid_constant__ DevTree T,
                                                                          const __grid_constant__ PlaceParams pp, int64_t n,
                                                                          const int32_t* __restrict__ sampleLists, PlaceResult* __restrict__ out,
                                                                          char* scratch, size_t warpBytes, unsigned laneK, unsigned laneP,
                                                                          unsigned laneA, int stackCap, int bestCap, unsigned long long* counter) {
    __shared__ DevModel sm;
    __shared__ PlaceWarp warps[kPlaceWarpThreads / 32];
    __shared__ long long ticket[kPlaceWarpThreads / 32];
    stage_model(sm, gm);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PlaceWarp& W = warps[wid];
    char* base = scratch + (blockIdx.x * (size_t)(kPlaceWarpThreads / 32) + wid) * warpBytes;
    PlaceWarpScratch ws;
    ws.laneK = laneK; ws.laneP = laneP; ws.laneA = laneA; ws.bestCap = bestCap; ws.stackCap = stackCap;
    ws.pay = reinterpret_cast<double*>(base);
    ws.ais = ws.pay + 32 * (size_t)laneP;
    ws.diffPay = ws.ais + 32 * (size_t)laneA;
    ws.eval = reinterpret_cast<PlaceEval*>(ws.diffPay + 6 * (size_t)laneK);
    ws.gpath = reinterpret_cast<PlacePath*>(ws.eval + bestCap);
    ws.best = reinterpret_cast<PlaceBest*>(ws.gpath + stackCap);
    ws.stack = reinterpret_cast<PlaceStackE*>(ws.best + bestCap);
    ws.key = reinterpret_cast<uint32_t*>(ws.stack + stackCap);
    ws.diffKey = ws.key + 32 * (size_t)laneK;
    ws.evalRc = reinterpret_cast<int*>(ws.diffKey + laneK);
    for (;;) {
        if (lane == 0) ticket[wid] = (long long)atomicAdd(counter, 1ULL);
        __syncwarp();
        const long long i = ticket[wid];
        __syncwarp();
        if (i >= n) break;
        const int64_t id = sampleLists[i];
        const int64_t ks = T.keyStart[id];
        PlaceResult r;
        r.bestNode = -1; r.status = 2; r.phase1 = r.missedMinors = 0; r.bestScore = r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
        if (ks >= 0) place_sample_warp(sm, T, pp, LRef{T.key + ks, T.pay + T.payStart[id], T.nkeys[id]}, W, ws, r);
        if (lane == 0) out[i] = r;
        __syncwarp();
    }
}

// the same with MAT trees covered (place_sample_warp_mat): lane 0 walks above mutation-carrying nodes, mutation-free subtrees are scan jobs
template <bool PAR>
__global__ void __launch_bounds__(kPlaceWarpThreads) k_place_samples_warp_mat(const __grid_constant__ DevModel gm, const __grid_constant__ DevTree T,
                                                                              const __grid_constant__ PlaceParams pp, int64_t n,
                                                                              const int32_t* __restrict__ sampleLists, PlaceResult* __restrict__ out,
                                                                              char* scratch, size_t warpBytes, unsigned laneK, unsigned laneP,
                                                                              unsigned laneA, int stackCap, int bestCap, unsigned long long* counter) {
    __shared__ DevModel sm;
    __shared__ PlaceWarpMat warps[kPlaceWarpThreads / 32];
    __shared__ long long ticket[kPlaceWarpThreads / 32];
    stage_model(sm, gm);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PlaceWarpMat& X = warps[wid];
    char* base = scratch + (blockIdx.x * (size_t)(kPlaceWarpThreads / 32) + wid) * warpBytes;
    PlaceWarpScratch ws;
    ws.laneK = laneK; ws.laneP = laneP; ws.laneA = laneA; ws.bestCap = bestCap; ws.stackCap = stackCap;
    ws.pay = reinterpret_cast<double*>(base);
    ws.ais = ws.pay + 32 * (size_t)laneP;
    ws.diffPay = ws.ais + 32 * (size_t)laneA;
    ws.eval = reinterpret_cast<PlaceEval*>(ws.diffPay + 6 * (size_t)laneK);
    ws.gpath = reinterpret_cast<PlacePath*>(ws.eval + bestCap);
    ws.best = reinterpret_cast<PlaceBest*>(ws.gpath + stackCap);
    ws.stack = reinterpret_cast<PlaceStackE*>(ws.best + bestCap);
    ws.key = reinterpret_cast<uint32_t*>(ws.stack + stackCap);
    ws.diffKey = ws.key + 32 * (size_t)laneK;
    ws.evalRc = reinterpret_cast<int*>(ws.diffKey + laneK);
    for (;;) {
        if (lane == 0) ticket[wid] = (long long)atomicAdd(counter, 1ULL);
        __syncwarp();
        const long long i = ticket[wid];
        __syncwarp();
        if (i >= n) break;
        const int64_t id = sampleLists[i];
        const int64_t ks = T.keyStart[id];
        PlaceResult r;
        r.bestNode = -1; r.status = 2; r.phase1 = r.missedMinors = 0; r.bestScore = r.bLenTop = r.bLenBottom = r.bLenAppend = 0.0;
        if (ks >= 0) place_sample_warp_mat<PAR>(sm, T, pp, LRef{T.key + ks, T.pay + T.payStart[id], T.nkeys[id]}, X, ws, r);
        if (lane == 0) out[i] = r;
        __syncwarp();
    }
}

// grid: whole waves of CTAs (multiples of the SM count), capped by the work
static int grid_for(const maple_ctx* ctx, int64_t n, int ctasPerSM) {
    int64_t need = (n + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)ctx->numSMs * ctasPerSM;
    if (need >= cap) return (int)cap;
    return (int)(need < 1 ? 1 : need);
}

extern "C" {

int maple_version(void) { return 100; }

const char* maple_last_error(const maple_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int maple_ctx_create(maple_ctx** out, int device, int32_t lRef, const double rootFreqs[4], int32_t flags) {
    if (!out || !rootFreqs || lRef <= 0 || lRef >= (1 << 24)) {
        g_err = "maple_ctx_create: bad argument (lRef must be in 1..2^24-1)";
        return MAPLE_E_ARG;
    }
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev <= device) {
        g_err = std::string("maple_ctx_create: no usable CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return MAPLE_E_NOGPU;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        g_err = "maple_ctx_create: cudaSetDevice failed";
        return MAPLE_E_NOGPU;
    }
    maple_ctx* ctx = new maple_ctx();
    ctx->device = device;
    if (const char* e = getenv("MAPLE_FSM_MINB")) ctx->fsmMinBlocks = atoi(e);
    if (const char* e = getenv("MAPLE_SCAN_OLD")) ctx->scanOldEnv = atoi(e) != 0;
    if (const char* e = getenv("MAPLE_FSM_SMS")) ctx->fsmSMs = atoi(e);
    if (const char* e = getenv("MAPLE_DENSE")) ctx->denseMode = atoi(e);
    if (const char* e = getenv("MAPLE_EVALQ")) ctx->evalQueueWarp = atoi(e) != 0;
    if (const char* e = getenv("MAPLE_DENSE_GB")) ctx->denseBudget = (size_t)atoll(e) << 30;
    ctx->scanOld = ctx->scanOldEnv;
    if (const char* e = getenv("MAPLE_LANES_PER_WARP")) { ctx->lanesPerWarp = atoi(e); if (ctx->lanesPerWarp < 0 || ctx->lanesPerWarp > 32) ctx->lanesPerWarp = 0; }
    if (const char* e = getenv("MAPLE_SCAN_REPLAY")) ctx->scanReplaySequential = strcmp(e, "sequential") == 0;
    if (const char* e = getenv("MAPLE_SCAN_APPEND")) ctx->scanAppendSitewise = strcmp(e, "q4") != 0;
    cudaDeviceGetAttribute(&ctx->numSMs, cudaDevAttrMultiProcessorCount, device);
    memset(&ctx->model, 0, sizeof(DevModel));
    ctx->model.lRef = lRef;
    ctx->model.U = (flags & MAPLE_F_USING_ERROR_RATE) ? 1 : 0;
    ctx->model.errSS = (flags & MAPLE_F_ERROR_SITE_SPECIFIC) ? 1 : 0;
    ctx->model.rateVar = (flags & MAPLE_F_RATE_VARIATION) ? 1 : 0;
    for (int i = 0; i < 4; i++) ctx->model.pi[i] = rootFreqs[i];
    ctx->model.thresholdProb = 1e-8;
    ctx->model.thresholdDiffForUpdate = 1e-5;
    ctx->model.thresholdFoldChangeUpdate = 1.01;
    ctx->model.minBLenSensitivity = 0.001 / lRef;
    *out = ctx;
    return MAPLE_OK;
}

int maple_ctx_destroy(maple_ctx* ctx) {
    if (!ctx) return MAPLE_OK;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->dSiteRates);
    cudaFree(ctx->dErrorRates);
    cudaFree(ctx->dCumRate);
    cudaFree(ctx->dCumErr);
    cudaFree(ctx->dPiLogErrCum);
    cudaFree(ctx->dCumBases);
    cudaFree(ctx->devStage);
    cudaFree(ctx->searchScratch);
    cudaFree(ctx->searchCounter);
    cudaFree(ctx->criticalCounter);
    if (ctx->criticalStream) cudaStreamDestroy(ctx->criticalStream);
    if (ctx->criticalEvA) cudaEventDestroy(ctx->criticalEvA);
    if (ctx->criticalEvB) cudaEventDestroy(ctx->criticalEvB);
    cudaFree(ctx->retryScratch);
    cudaFree(ctx->placeScratch);
    cudaFree(ctx->retryCounters);
    cudaFree(ctx->treeDerived);
    cudaFree(ctx->searchStats);
    cudaFree(ctx->scanUnits);
    cudaFree(ctx->scanArena);
    cudaFree(ctx->scanRecs);
    cudaFree(ctx->queueMem);
    cudaFree(ctx->denseMem);
    cudaFree(ctx->updateMem);
    cudaFree(ctx->evalMem);
    delete ctx;
    return MAPLE_OK;
}

static int upload(maple_ctx* ctx, double** dst, const double* src, size_t n) {
    if (!*dst) CK(cudaMalloc((void**)dst, n * sizeof(double)));
    CK(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
    return MAPLE_OK;
}

int maple_ctx_set_model(maple_ctx* ctx, const double Q[16], const double* siteRates, double errorRate, const double* errorRates,
                        const double* cumulativeRate, const double* cumulativeErrorRate, double totError) {
    if (!ctx || !Q || !cumulativeRate) return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    DevModel& m = ctx->model;
    if (m.rateVar && !siteRates) { ctx->err = "set_model: rate variation needs siteRates"; return MAPLE_E_ARG; }
    if (m.U && m.errSS && (!errorRates || !cumulativeErrorRate)) { ctx->err = "set_model: site-specific errors need errorRates and cumulativeErrorRate"; return MAPLE_E_ARG; }
    for (int i = 0; i < 16; i++) m.Q[i] = Q[i];
    m.errorRate = errorRate;
    m.totError = totError;
    int rc;
    const size_t L = (size_t)m.lRef;
    if (m.rateVar) { if ((rc = upload(ctx, &ctx->dSiteRates, siteRates, L))) return rc; }
    if (m.U && m.errSS) {
        if ((rc = upload(ctx, &ctx->dErrorRates, errorRates, L))) return rc;
        if ((rc = upload(ctx, &ctx->dCumErr, cumulativeErrorRate, L + 1))) return rc;
    }
    if ((rc = upload(ctx, &ctx->dCumRate, cumulativeRate, L + 1))) return rc;
    m.siteRates = ctx->dSiteRates;
    m.errorRates = ctx->dErrorRates;
    m.cumRate = ctx->dCumRate;
    m.cumErr = ctx->dCumErr;
    ctx->haveModel = true;
    return MAPLE_OK;
}

int maple_ctx_set_thresholds(maple_ctx* ctx, double thresholdProb, double thresholdDiffForUpdate, double thresholdFoldChangeUpdate,
                             double minBLenSensitivity) {
    if (!ctx) return MAPLE_E_ARG;
    ctx->model.thresholdProb = thresholdProb;
    ctx->model.thresholdDiffForUpdate = thresholdDiffForUpdate;
    ctx->model.thresholdFoldChangeUpdate = thresholdFoldChangeUpdate;
    ctx->model.minBLenSensitivity = minBLenSensitivity;
    return MAPLE_OK;
}

int maple_lists_bind(maple_ctx* ctx, const uint32_t* key, const double* pay, const int64_t* key_start, const int64_t* pay_start,
                     int64_t nLists) {
    if (!ctx || !key || !pay || !key_start || !pay_start || nLists <= 0) return MAPLE_E_ARG;
    ctx->key = key;
    ctx->pay = pay;
    ctx->keyStart = key_start;
    ctx->payStart = pay_start;
    ctx->nLists = nLists;
    ctx->haveLists = true;
    return MAPLE_OK;
}

static int ready(maple_ctx* ctx) {
    if (!ctx) return MAPLE_E_ARG;
    if (!ctx->haveModel || !ctx->haveLists) {
        ctx->err = "model or lists not set (maple_ctx_set_model / maple_lists_bind)";
        return MAPLE_E_STATE;
    }
    return MAPLE_OK;
}

int maple_append_prob_batch(maple_ctx* ctx, int64_t n, const int32_t* pIdx, const int32_t* cIdx, const uint8_t* isTipC,
                            const double* bLen, double* out, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !pIdx || !cIdx || !isTipC || !bLen || !out) return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    Arena A{ctx->key, ctx->pay, ctx->keyStart, ctx->payStart};
    k_append<<<grid_for(ctx, n, 16), kThreads, 0, (cudaStream_t)stream>>>(ctx->model, A, n, pIdx, cIdx, isTipC, bLen, out);
    ctx->launches++;
    CK(cudaGetLastError());
    return MAPLE_OK;
}

static int ensure_stage(maple_ctx* ctx, size_t bytes) {
    if (bytes > ctx->devStageBytes) {
        cudaFree(ctx->devStage);
        ctx->devStage = nullptr;
        ctx->devStageBytes = 0;
        CK(cudaMalloc(&ctx->devStage, bytes + bytes / 4));
        ctx->devStageBytes = bytes + bytes / 4;
    }
    return MAPLE_OK;
}

int maple_append_prob_batch_host(maple_ctx* ctx, int64_t n, const int32_t* pIdx, const int32_t* cIdx, const uint8_t* isTipC,
                                 const double* bLen, double* out) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !pIdx || !cIdx || !isTipC || !bLen || !out) return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    // device staging block: bLen[n] f64 | out[n] f64 | pIdx[n] i32 | cIdx[n] i32 | isTip[n] u8.  The copies go
    // straight from / to the caller's buffers (DMA when they are page-locked, driver-staged otherwise).
    const size_t N = (size_t)n;
    const size_t oB = 0, oO = oB + 8 * N, oP = oO + 8 * N, oC = oP + 4 * N, oT = oC + 4 * N, total = oT + ((N + 15) / 16) * 16;
    if ((rc = ensure_stage(ctx, total))) return rc;
    char* d = (char*)ctx->devStage;
    cudaStream_t s = 0;
    CK(cudaMemcpyAsync(d + oB, bLen, 8 * N, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d + oP, pIdx, 4 * N, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d + oC, cIdx, 4 * N, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d + oT, isTipC, N, cudaMemcpyHostToDevice, s));
    rc = maple_append_prob_batch(ctx, n, (const int32_t*)(d + oP), (const int32_t*)(d + oC), (const uint8_t*)(d + oT),
                                 (const double*)(d + oB), (double*)(d + oO), s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, d + oO, 8 * N, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MAPLE_OK;
}

int maple_merge_batch(maple_ctx* ctx, int64_t n, const int32_t* idx1, const double* bLen1, const uint8_t* fromTip1, const int32_t* idx2,
                      const double* bLen2, const uint8_t* fromTip2, const uint8_t* flags, const int32_t* numMinor1,
                      const int32_t* numMinor2, uint32_t* out_key, double* out_pay, const int64_t* out_key_start,
                      const int64_t* out_pay_start, int32_t* out_nkeys, int32_t* out_npay, double* out_lk, int32_t* out_status,
                      int32_t shorten, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !idx1 || !bLen1 || !fromTip1 || !idx2 || !bLen2 || !fromTip2 || !flags || !out_key || !out_pay || !out_key_start ||
        !out_pay_start || !out_nkeys || !out_npay || !out_status)
        return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    Arena A{ctx->key, ctx->pay, ctx->keyStart, ctx->payStart};
    k_merge<<<grid_for(ctx, n, 8), kThreads, 0, (cudaStream_t)stream>>>(ctx->model, A, n, idx1, bLen1, fromTip1, idx2, bLen2, fromTip2, flags,
                                                                       numMinor1, numMinor2, out_key, out_pay, out_key_start,
                                                                       out_pay_start, out_nkeys, out_npay, out_lk, out_status, shorten);
    ctx->launches++;
    CK(cudaGetLastError());
    return MAPLE_OK;
}

int maple_blen_batch(maple_ctx* ctx, int64_t n, const int32_t* pIdx, const int32_t* cIdx, const uint8_t* fromTipC, double* scratch,
                     const int64_t* scratch_start, double* out, int32_t* out_status, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !pIdx || !cIdx || !fromTipC || !scratch || !scratch_start || !out || !out_status) return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    Arena A{ctx->key, ctx->pay, ctx->keyStart, ctx->payStart};
    k_blen<<<grid_for(ctx, n, 8), kThreads, 0, (cudaStream_t)stream>>>(ctx->model, A, n, pIdx, cIdx, fromTipC, scratch, scratch_start, out,
                                                                      out_status);
    ctx->launches++;
    CK(cudaGetLastError());
    return MAPLE_OK;
}

int maple_vectors_differ_batch(maple_ctx* ctx, int64_t n, const int32_t* idx1, const int32_t* idx2, uint8_t* out, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !idx1 || !idx2 || !out) return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    Arena A{ctx->key, ctx->pay, ctx->keyStart, ctx->payStart};
    k_differ<<<grid_for(ctx, n, 16), kThreads, 0, (cudaStream_t)stream>>>(ctx->model, A, n, idx1, idx2, out);
    ctx->launches++;
    CK(cudaGetLastError());
    return MAPLE_OK;
}

int maple_root_vector_batch(maple_ctx* ctx, int64_t n, const int32_t* idx, const double* bLen, const uint8_t* isFromTip, uint32_t* out_key,
                            double* out_pay, const int64_t* out_key_start, const int64_t* out_pay_start, int32_t* out_nkeys,
                            int32_t* out_npay, int32_t shorten, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !idx || !bLen || !isFromTip || !out_key || !out_pay || !out_key_start || !out_pay_start || !out_nkeys || !out_npay)
        return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    Arena A{ctx->key, ctx->pay, ctx->keyStart, ctx->payStart};
    k_root_vector<<<grid_for(ctx, n, 8), kThreads, 0, (cudaStream_t)stream>>>(ctx->model, A, n, idx, bLen, isFromTip, out_key, out_pay,
                                                                             out_key_start, out_pay_start, out_nkeys, out_npay, shorten);
    ctx->launches++;
    CK(cudaGetLastError());
    return MAPLE_OK;
}

int maple_ctx_set_root_tables(maple_ctx* ctx, const int32_t* cumulativeBases, const double* rootFreqsLogErrorCumulative) {
    if (!ctx || !cumulativeBases) return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    DevModel& m = ctx->model;
    const size_t L = (size_t)m.lRef;
    if (m.U && !rootFreqsLogErrorCumulative) { ctx->err = "set_root_tables: the error model needs rootFreqsLogErrorCumulative"; return MAPLE_E_ARG; }
    if (!ctx->dCumBases) CK(cudaMalloc((void**)&ctx->dCumBases, (L + 1) * 4 * sizeof(int32_t)));
    CK(cudaMemcpy(ctx->dCumBases, cumulativeBases, (L + 1) * 4 * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (rootFreqsLogErrorCumulative) {
        int rc = upload(ctx, &ctx->dPiLogErrCum, rootFreqsLogErrorCumulative, L + 1);
        if (rc) return rc;
    }
    m.cumBases = ctx->dCumBases;
    m.piLogErrCum = ctx->dPiLogErrCum;
    return MAPLE_OK;
}

int maple_prob_root_batch(maple_ctx* ctx, int64_t n, const int32_t* idx, double* out, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (!ctx->model.cumBases) { ctx->err = "maple_prob_root_batch: root tables not set (maple_ctx_set_root_tables)"; return MAPLE_E_STATE; }
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !idx || !out) return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    Arena A{ctx->key, ctx->pay, ctx->keyStart, ctx->payStart};
    k_prob_root<<<grid_for(ctx, n, 8), kThreads, 0, (cudaStream_t)stream>>>(ctx->model, A, n, idx, out);
    ctx->launches++;
    CK(cudaGetLastError());
    return MAPLE_OK;
}

int maple_pass_branch_batch(maple_ctx* ctx, int64_t n, const int32_t* idx, const int32_t* mutNode, const uint8_t* dirIsUp,
                            const int32_t* mutStart, const int32_t* mut, uint32_t* out_key, double* out_pay, const int64_t* out_key_start,
                            const int64_t* out_pay_start, int32_t* out_nkeys, int32_t* out_npay, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !idx || !mutNode || !dirIsUp || !mutStart || !mut || !out_key || !out_pay || !out_key_start || !out_pay_start || !out_nkeys ||
        !out_npay)
        return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    Arena A{ctx->key, ctx->pay, ctx->keyStart, ctx->payStart};
    k_pass_branch<<<grid_for(ctx, n, 8), kThreads, 0, (cudaStream_t)stream>>>(ctx->model.lRef, A, n, idx, mutNode, dirIsUp, mutStart, mut, out_key,
                                                                             out_pay, out_key_start, out_pay_start, out_nkeys, out_npay);
    ctx->launches++;
    CK(cudaGetLastError());
    return MAPLE_OK;
}

int maple_lists_copy(maple_ctx* ctx, int64_t n, const uint32_t* src_key, const double* src_pay, const int64_t* src_key_start,
                     const int64_t* src_pay_start, const int32_t* nkeys, const int32_t* npay, uint32_t* dst_key, double* dst_pay,
                     const int64_t* dst_key_start, const int64_t* dst_pay_start, void* stream) {
    if (!ctx) return MAPLE_E_ARG;
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !src_key || !src_pay || !src_key_start || !src_pay_start || !nkeys || !npay || !dst_key || !dst_pay || !dst_key_start ||
        !dst_pay_start)
        return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int64_t blocks = (n * 32 + 255) / 256;
    int64_t cap = (int64_t)ctx->numSMs * 8;
    k_lists_copy<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(n, src_key, src_pay, src_key_start, src_pay_start,
                                                                                      nkeys, npay, dst_key, dst_pay, dst_key_start,
                                                                                      dst_pay_start);
    ctx->launches++;
    CK(cudaGetLastError());
    return MAPLE_OK;
}

int maple_tree_bind(maple_ctx* ctx, int32_t nNodes, int32_t root, const int32_t* up, const int32_t* child0, const int32_t* child1,
                    const double* dist, const uint8_t* isTip, const int32_t* mutStart, const int32_t* mut, const int32_t* nkeys, const int32_t* npay) {
    if (!ctx || nNodes <= 0 || root < 0 || root >= nNodes || !up || !child0 || !child1 || !dist || !isTip || !nkeys) return MAPLE_E_ARG;
    if (!ctx->haveLists || ctx->nLists < 4 * (int64_t)nNodes) {
        ctx->err = "maple_tree_bind: bind an arena with 4*nNodes lists first (list id = family*nNodes + node)";
        return MAPLE_E_STATE;
    }
    DevTree& t = ctx->tree;
    t.nNodes = nNodes; t.root = root; t.up = up; t.child0 = child0; t.child1 = child1; t.dist = dist; t.isTip = isTip;
    t.mutStart = mutStart; t.mut = mut; t.nkeys = nkeys; t.npay = npay;
    // Derived arrays for the warp-cooperative subtree scans: the order in which findBestParentTopology walks down a
    // subtree (children pushed 0 then 1, so child 1 is explored first, :7112-7170), subtree sizes and depths.
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)nNodes;
    std::vector<int32_t> hUp(n), hC0(n), hC1(n), hMs;
    CK(cudaMemcpy(hUp.data(), up, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hC0.data(), child0, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hC1.data(), child1, n * 4, cudaMemcpyDeviceToHost));
    if (mutStart) {
        hMs.resize(n + 1);
        CK(cudaMemcpy(hMs.data(), mutStart, (n + 1) * 4, cudaMemcpyDeviceToHost));
    }
    std::vector<int32_t> der(4 * n + (n + 3) / 4, 0);
    int32_t *order = der.data(), *pre = order + n, *size = pre + n, *depth = size + n;
    uint8_t* mutBelow = reinterpret_cast<uint8_t*>(depth + n);
    for (size_t i = 0; i < n; i++) { pre[i] = -1; size[i] = 1; }
    std::vector<int32_t> st;
    st.push_back(root);
    depth[root] = 0;
    size_t cnt = 0;
    int height = 0;
    while (!st.empty()) {
        const int32_t v = st.back();
        st.pop_back();
        if (cnt >= n || pre[v] >= 0) { ctx->err = "maple_tree_bind: up/child arrays do not describe a tree"; return MAPLE_E_ARG; }
        pre[v] = (int32_t)cnt;
        order[cnt++] = v;
        if (depth[v] > height) height = depth[v];
        if (hC0[v] >= 0) {
            if (hC0[v] >= nNodes || hC1[v] < 0 || hC1[v] >= nNodes) { ctx->err = "maple_tree_bind: bad child index"; return MAPLE_E_ARG; }
            depth[hC0[v]] = depth[hC1[v]] = depth[v] + 1;
            st.push_back(hC0[v]);
            st.push_back(hC1[v]);
        }
    }
    for (size_t i = cnt; i-- > 1;) {  // children come after their parent in pre-order
        const int32_t v = order[i], p = hUp[v];
        size[p] += size[v];
        if (mutBelow[v] || (mutStart && hMs[v + 1] > hMs[v])) mutBelow[p] = 1;
    }
    cudaFree(ctx->treeDerived);
    ctx->treeDerived = nullptr;
    CK(cudaMalloc((void**)&ctx->treeDerived, der.size() * 4 + 32 + n * sizeof(ScanNode)));
    CK(cudaMemcpy(ctx->treeDerived, der.data(), der.size() * 4, cudaMemcpyHostToDevice));
    t.order = ctx->treeDerived; t.pre = t.order + n; t.size = t.pre + n; t.depth = t.size + n;
    t.mutBelow = reinterpret_cast<const uint8_t*>(t.depth + n);
    t.scan = reinterpret_cast<const ScanNode*>((reinterpret_cast<uintptr_t>(ctx->treeDerived + der.size()) + 31) & ~uintptr_t(31));
    ctx->treeHeight = height;
    ctx->treeHasMut = mutStart && (mutBelow[root] || hMs[root + 1] > hMs[root]);
    ctx->haveTree = true;
    // second form of the scans: where the scan-format copy of each probVectTotUp list goes (dense, in pre-order)
    ctx->scan2Ok = false;
    cudaFree(ctx->scanUnits); cudaFree(ctx->scanArena); cudaFree(ctx->scanRecs);
    ctx->scanUnits = nullptr; ctx->scanOffsets = nullptr; ctx->scanArena = nullptr; ctx->scanRecs = nullptr;
    t.scan2 = nullptr; t.scanArena = nullptr; t.scanOff = nullptr;
    if (height < 65535) {
        CK(cudaMalloc((void**)&ctx->scanUnits, 2 * n * sizeof(uint32_t)));
        ctx->scanOffsets = ctx->scanUnits + n;
        DevTree T = t;
        T.key = ctx->key; T.pay = ctx->pay; T.keyStart = ctx->keyStart; T.payStart = ctx->payStart;
        k_scan_count<<<(unsigned)((n + 255) / 256), 256>>>(T, ctx->scanUnits, ctx->model.U != 0);
        ctx->launches++;
        std::vector<uint32_t> hu(n), ho(n);
        CK(cudaMemcpy(hu.data(), ctx->scanUnits, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        uint64_t tot = 0;
        ctx->scanMaxUnits = 0;
        ctx->scanNumLists = 0;
        ctx->scanAllStaged = true;
        bool fixUnits = false;
        for (size_t i = 0; i < n; i++) {
            if (hu[i] == ~0u) { hu[i] = 0; ctx->scanAllStaged = false; fixUnits = true; }  // a list too large for a scan-format copy
            const uint64_t u = (hu[i] & 0xffffu) + (hu[i] >> 16);
            if (u == 0 || tot + u >= 0xffffffffull) {
                if (u) ctx->scanAllStaged = false;
                ho[i] = ~0u;
                continue;
            }
            ho[i] = (uint32_t)tot;
            tot += u;
            ctx->scanNumLists++;
            if (u > ctx->scanMaxUnits) ctx->scanMaxUnits = (uint32_t)u;
        }
        if (fixUnits) CK(cudaMemcpy(ctx->scanUnits, hu.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->scanOffsets, ho.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CK(cudaMalloc((void**)&ctx->scanArena, (size_t)(tot + 4) * sizeof(uint4)));
        CK(cudaMalloc((void**)&ctx->scanRecs, n * sizeof(ScanRec)));
        t.scan2 = ctx->scanRecs; t.scanArena = ctx->scanArena; t.scanOff = ctx->scanOffsets;
        ctx->scan2Ok = true;
    }
    return MAPLE_OK;
}

int maple_spr_search_batch(maple_ctx* ctx, const maple_search_params* p, int64_t n, const int32_t* nodes, maple_search_result* out,
                           int32_t scratch_keys_per_search, int32_t max_concurrent_searches, int64_t* out_cycles, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (!ctx->haveTree) { ctx->err = "maple_spr_search_batch: no tree bound (maple_tree_bind)"; return MAPLE_E_STATE; }
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !p || !nodes || !out) return MAPLE_E_ARG;
    static_assert(sizeof(maple_search_result) == sizeof(SearchResult), "result record layout");
    static_assert(sizeof(maple_search_params) == sizeof(SearchParams), "params layout");
    CK(cudaSetDevice(ctx->device));
    SearchParams sp;
    memcpy(&sp, p, sizeof sp);
    DevTree T = ctx->tree;
    T.key = ctx->key; T.pay = ctx->pay; T.keyStart = ctx->keyStart; T.payStart = ctx->payStart;
    const unsigned capK = (unsigned)((scratch_keys_per_search > 0 ? scratch_keys_per_search : 8192) + 3) & ~3u;
    // coefficient scratch of the branch-length solver: one value per entry of the two lists, which may be scratch lists, so it
    // grows with the list scratch (2048 at the default 8192 entries)
    const unsigned capP = 2 * capK + 6 * 1024, capA = capK / 4 > 2048 ? capK / 4 : 2048;
    // the DFS keeps at most one pending entry per level on the way up and one per level on the way down; what is left of the
    // stack doubles as the per-depth state of the subtree scans (5 states per free entry)
    const int stackCap = (2 * ctx->treeHeight + 32 + 63) & ~63;
    int blocksPerSM = 0;
    if (ctx->searchVariant == 1) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, k_spr_search, kSearchThreads, 0));
    // shared memory: a fixed part per warp plus as much list pool as the targeted CTAs per SM leave (227 KB per SM, 1 KB reserved per CTA)
    const bool scan2 = ctx->searchVariant == 0 && ctx->scan2Ok && !ctx->scanOld && T.order && ctx->scanMinSize > 0;
    const int ctasWanted = ctx->fsmMinBlocks == 6 ? 6 : 8;
    const int fixedPerWarp = int((scan2 ? sizeof(Scan2Smem) : sizeof(ScanSmem)) - sizeof(uint4));
    int poolBytes = ((227 * 1024 / ctasWanted - 2048) / (kSearchThreads / 32) - fixedPerWarp) & ~15;
    if (poolBytes > (scan2 ? 16384 : 12288)) poolBytes = scan2 ? 16384 : 12288;
    const size_t fsmSmem = (kSearchThreads / 32) * (size_t)(fixedPerWarp + poolBytes);
    using FsmKernel = void (*)(const DevModel, const DevTree, const SearchParams, int64_t, const int32_t*, SearchResult*, uint32_t*, double*, double*,
                               StackE*, unsigned, unsigned, unsigned, int, unsigned long long*, long long*, int, int, int, unsigned long long*,
                               const unsigned long long*, const int32_t*, int, const BigScratch, const ScanQueue, int, const DenseScores, const EvalScratch);
    // Register budget = resident warps.  __launch_bounds__(64, 7) makes ptxas settle on 128 registers with few spills, which lets
    // 8 CTAs (16 warps) share an SM: 3.1 s for the deep round at 100 k sequences against 4.4 s for the 168-register build
    // (12 warps) on the same box.  MAPLE_FSM_MINB=6 selects the latter for A/B runs.  (Register allocation of this kernel is
    // touchy: check `-Xptxas -v` after changing it -- a 128-register build with ~2 kB of spills is as slow as 168 registers.)
    // Keep the three instantiations: with only <6> and <7> present the same <7> comes out with 2 kB of spills.  96-register builds
    // (<9>, <10>: 10 CTAs per SM, half the list pool) were measured too: 5.5 s against 3.1-3.4 s.
    // <.., SCAN2 = second form of the subtree scans, EXTRAS = scan service + dense scoring pass compiled in (chosen below, once it
    // is known whether either is on for this launch)>
    FsmKernel fsmKernel = k_spr_search_fsm<6, false, false>;
    if (ctx->fsmMinBlocks == 8) fsmKernel = k_spr_search_fsm<8, false, false>;
    if (ctx->fsmMinBlocks == 7) fsmKernel = k_spr_search_fsm<7, false, false>;
    if (scan2) {
        fsmKernel = k_spr_search_fsm<6, true, false>;
        if (ctx->fsmMinBlocks == 8) fsmKernel = k_spr_search_fsm<8, true, false>;
        if (ctx->fsmMinBlocks == 7) fsmKernel = k_spr_search_fsm<7, true, false>;
    }
    if (scan2 && (ctx->fsmSMs != 0 || ctx->denseMode != 0))
        fsmKernel = ctx->fsmMinBlocks == 6 ? k_spr_search_fsm<6, true, true> : k_spr_search_fsm<7, true, true>;
    if (scan2 && !ctx->criticalStream) {
        // what a launch with critical searches needs, set up with the first batch of a context whether or not it has any: the
        // first use must not cost more than the later ones (callers compare the two launch shapes by their time)
        CK(cudaStreamCreateWithFlags(&ctx->criticalStream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->criticalEvA, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->criticalEvB, cudaEventDisableTiming));
        CK(cudaMalloc((void**)&ctx->criticalCounter, sizeof(unsigned long long)));
        int maxOptin = 0;
        CK(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, k_spr_search_fsm<7, true, false>));
        ctx->criticalHogSmem = (size_t)maxOptin - fa.sharedSizeBytes;
        CK(cudaFuncGetAttributes(&fa, k_wait_started));  // (loads it)
    }
    if (ctx->searchVariant != 1) {
        CK(cudaFuncSetAttribute(fsmKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(scan2 && ctx->criticalHogSmem > fsmSmem ? ctx->criticalHogSmem : fsmSmem)));
        CK(cudaFuncSetAttribute(fsmKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, fsmKernel, kSearchThreads, fsmSmem));
    }
    if (blocksPerSM < 1) blocksPerSM = 1;
    int64_t threads = (int64_t)ctx->numSMs * blocksPerSM * kSearchThreads;
    if (max_concurrent_searches > 0 && threads > max_concurrent_searches) threads = max_concurrent_searches < 32 ? 32 : max_concurrent_searches;
    // Searches per warp.  The subtree scans -- most of the work of a deep round -- are executed by whole warps, so the unit that
    // has to be kept busy is the warp, not the lane: every resident warp should own searches, and each owning lane should get
    // a few searches in turn (dynamic balance) rather than one.  With few searches (a shard of a multi-GPU round) this spreads
    // them over all warps instead of packing 32 into each of a few.
    // Scan service: searches owned by the CTAs of fsmSMs SMs (32 to a warp), subtree scans served by all other CTAs.  Needs the
    // whole grid resident (it is: sized from the occupancy query) and every list to fit a server's pool next to the removed list.
    int fsmSMs = 0;
    if (scan2 && ctx->scanAllStaged && (size_t)ctx->scanMaxUnits * 16 + 64 <= (size_t)poolBytes / 2 && max_concurrent_searches == 0 &&
        ctx->numSMs >= 8 && ctx->lanesPerWarp <= 0) {
        fsmSMs = ctx->fsmSMs;
        if (fsmSMs == 0) goto noService;
        // the strict rules of the fast round leave little to scan (most of the work is on the lanes), the deep rounds are nearly all scans
        if (fsmSMs < 0) fsmSMs = sp.strictTopologyStopRules ? ctx->numSMs / 2 : ctx->numSMs / 6;
        if (fsmSMs > ctx->numSMs - 4) fsmSMs = ctx->numSMs - 4;
        // few searches: no more owning warps than searches / 8
        const int64_t wantWarps = (n + 255) / 256;
        const int warpsPerSM = blocksPerSM * (kSearchThreads / 32);
        if ((int64_t)fsmSMs * warpsPerSM > wantWarps) fsmSMs = (int)((wantWarps + warpsPerSM - 1) / warpsPerSM);
        if (fsmSMs < 1) fsmSMs = 1;
    }
noService:
    // Searches that run on an SM of their own (maple_ctx_set_critical_searches): the first nCrit entries of the list go to a launch
    // of their own -- one single-warp CTA each, which asks for a whole SM's shared memory so that nothing else is scheduled next
    // to it -- and the rest of the list to the usual launch on the remaining SMs.
    int nCrit = 0;
    if (scan2 && fsmSMs == 0 && ctx->criticalSearches > 0 && ctx->denseMode == 0 && max_concurrent_searches == 0 && ctx->searchVariant == 0) {
        nCrit = ctx->criticalSearches;
        if (nCrit > ctx->numSMs / 2) nCrit = ctx->numSMs / 2;
        if ((int64_t)nCrit * 4 > n) nCrit = 0;
    }
    const int32_t* const nodesAll = nodes;
    const int64_t nAll = n;
    if (nCrit) {
        nodes += nCrit;
        n -= nCrit;
        threads = (int64_t)(ctx->numSMs - nCrit) * blocksPerSM * kSearchThreads;
    }
    int lpw = 32;
    if (fsmSMs > 0) {
        lpw = 32;
    } else if (ctx->searchVariant != 1) {
        lpw = ctx->lanesPerWarp;
        if (lpw <= 0) {
            const int64_t warps = threads / 32;
            lpw = (int)((n + warps * 3 - 1) / (warps * 3));
            lpw = lpw < 2 ? 2 : lpw > 32 ? 32 : lpw;
        }
        if (threads > (n + lpw - 1) / lpw * 32) threads = (n + lpw - 1) / lpw * 32;
    } else if (threads > n) threads = n;
    int blocks = (int)((threads + kSearchThreads - 1) / kSearchThreads);
    threads = (int64_t)blocks * kSearchThreads;
    // head of the list run one per warp (maple_ctx_set_head_searches): relative to the main launch's part of the list
    unsigned long long headEnd = 0;
    if (scan2 && fsmSMs == 0 && ctx->searchVariant == 0 && ctx->headSearches > nCrit && max_concurrent_searches == 0 && lpw > 1) {
        headEnd = (unsigned long long)(ctx->headSearches - nCrit);
        if ((int64_t)headEnd > n) headEnd = (unsigned long long)n;
        if (headEnd < (unsigned long long)(threads / 32)) headEnd = 0;  // fewer than one per warp: the static first pull does the same
    }
    const size_t perThread = (size_t)capK * 4 + (size_t)capP * 8 + (size_t)capA * 8 + (size_t)stackCap * sizeof(StackE);
    // lanes that own a search (and scratch); with the scan service: the warps of fsmSMs SMs and of CTA 0
    const size_t owners = fsmSMs > 0 ? ((size_t)fsmSMs * blocksPerSM + 1) * (kSearchThreads / 32) * 32 : (size_t)threads / 32 * lpw;
    const size_t mainBytes = (perThread * owners + 255) & ~size_t(255);
    // (room for the searches of a critical launch always: switching it on must not cost a reallocation)
    const size_t need = mainBytes + perThread * (size_t)(ctx->numSMs / 2) + 256;
    if (need > ctx->searchScratchBytes) {
        cudaFree(ctx->searchScratch);
        ctx->searchScratch = nullptr;
        ctx->searchScratchBytes = 0;
        CK(cudaMalloc(&ctx->searchScratch, need));
        ctx->searchScratchBytes = need;
    }
    if (!ctx->searchCounter) CK(cudaMalloc((void**)&ctx->searchCounter, sizeof(unsigned long long)));
    {   // the state-machine kernel hands the first `owners` entries out statically (fsm_warp_loop), the counter serves the rest
        // (with a head of the list run one per warp, only lane 0 of every warp has a static first entry)
        const unsigned long long first = (ctx->searchVariant != 1 && fsmSMs == 0) ? (unsigned long long)(headEnd ? owners / lpw : owners) : 0ULL;
        CK(cudaMemcpyAsync(ctx->searchCounter, &first, sizeof first, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    }
    char* base = (char*)ctx->searchScratch;
    double* scrPay = (double*)base;
    double* scrAis = (double*)(base + owners * capP * 8);
    StackE* scrStack = (StackE*)(base + owners * (capP + capA) * 8);
    uint32_t* scrKey = (uint32_t*)(base + owners * ((size_t)(capP + capA) * 8 + (size_t)stackCap * sizeof(StackE)));
    // A few slots with 8x the scratch, shared by the launch: a search that exhausts its lane's scratch takes one and starts over
    // inside the same launch (fsm_warp_loop), so the rare long search does not cost a serial launch of its own afterwards.
    constexpr int kBigSlots = kSearchThreads;  // 64
    BigScratch big{};
    const int retryCap = (int)(n < 65536 ? n : 65536);
    int32_t *retryNodes = nullptr, *retryIdx = nullptr;
    if (ctx->searchVariant != 1) {
        const unsigned capK2 = capK * 8, capP2 = 2 * capK2 + 6 * 1024, capA2 = capA * 8;
        const size_t per2 = (size_t)capK2 * 4 + (size_t)capP2 * 8 + (size_t)capA2 * 8 + (size_t)stackCap * sizeof(StackE);
        const size_t need2 = per2 * kBigSlots + 256 + (size_t)65536 * 8 + 64;
        if (need2 > ctx->retryScratchBytes) {
            cudaFree(ctx->retryScratch);
            ctx->retryScratch = nullptr;
            ctx->retryScratchBytes = 0;
            CK(cudaMalloc(&ctx->retryScratch, need2));
            ctx->retryScratchBytes = need2;
        }
        if (!ctx->retryCounters) CK(cudaMalloc((void**)&ctx->retryCounters, 4 * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(ctx->retryCounters, 0, 4 * sizeof(unsigned long long), (cudaStream_t)stream));
        char* b2 = (char*)ctx->retryScratch;
        retryNodes = (int32_t*)b2;
        retryIdx = retryNodes + 65536;
        char* s2 = b2 + (size_t)65536 * 8 + 64;
        big.pay = (double*)s2;
        big.ais = (double*)(s2 + (size_t)kBigSlots * capP2 * 8);
        big.stack = (StackE*)(s2 + (size_t)kBigSlots * (capP2 + capA2) * 8);
        big.key = (uint32_t*)(s2 + (size_t)kBigSlots * ((size_t)(capP2 + capA2) * 8 + (size_t)stackCap * sizeof(StackE)));
        big.capK = capK2; big.capP = capP2; big.capA = capA2;
        big.nSlots = kBigSlots;
        big.counter = ctx->retryCounters + 2;
    }
    ScanQueue sq{};
    DenseScores ds{};
    EvalScratch es{};
    PhaseTimer timer((cudaStream_t)stream);
    if (ctx->searchVariant != 1 && ctx->evalQueueWarp) {
        // a slice per lane of the launch for the warp-wide evaluation of queued phase-2 entries: four merged lists of an
        // evaluatePlacement (a few hundred entries) fit; an entry that does not is evaluated by the owning lane in its own scratch
        es.capK = 1024; es.capP = 6 * 1024; es.capA = 1024;
        const size_t perLane = (size_t)es.capK * 4 + (size_t)es.capP * 8 + (size_t)es.capA * 8;
        const size_t needE = perLane * ((size_t)threads + (size_t)(ctx->numSMs / 2) * 32) + 256;
        if (needE > ctx->evalBytes) {
            cudaFree(ctx->evalMem);
            ctx->evalMem = nullptr;
            ctx->evalBytes = 0;
            if (cudaMalloc(&ctx->evalMem, needE) == cudaSuccess) ctx->evalBytes = needE;
            else { ctx->evalMem = nullptr; (void)cudaGetLastError(); }
        }
        if (ctx->evalMem) {
            const size_t lanes = (size_t)threads + (size_t)nCrit * 32;
            es.pay = (double*)ctx->evalMem;
            es.ais = es.pay + lanes * es.capP;
            es.key = (uint32_t*)(es.ais + lanes * es.capA);
        }
    }
    if (fsmSMs > 0) {
        unsigned cap = 1024;
        while (cap < 4 * owners) cap <<= 1;
        const size_t bytes = 1024 + (size_t)cap * 8 + owners * sizeof(ScanJob);
        if (bytes > ctx->queueBytes) {
            cudaFree(ctx->queueMem);
            ctx->queueMem = nullptr;
            ctx->queueBytes = 0;
            CK(cudaMalloc(&ctx->queueMem, bytes));
            ctx->queueBytes = bytes;
        }
        CK(cudaMemsetAsync(ctx->queueMem, 0, bytes, (cudaStream_t)stream));
        char* q = (char*)ctx->queueMem;
        sq.head = (unsigned long long*)q;  // every control word on a 256-byte line of its own: the servers poll head / tail
        sq.tail = (unsigned long long*)(q + 256);
        sq.doneSearches = (unsigned long long*)(q + 512);
        sq.ownerCounter = (unsigned long long*)(q + 768);
        sq.ring = (unsigned long long*)(q + 1024);
        sq.jobs = (ScanJob*)(q + 1024 + (size_t)cap * 8);
        sq.cap = cap;
        sq.maxOwners = (int)owners;
    }
    if (scan2) {
        k_scan_build<<<(T.nNodes + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->model, T, sp.effectivelyNon0BLen, ctx->scanUnits, ctx->scanArena,
                                                                              ctx->scanRecs);
        k_scan_nsa<<<(T.nNodes + 255) / 256, 256, 0, (cudaStream_t)stream>>>(T, ctx->scanRecs);
        ctx->launches += 2;
        timer.mark("scan-format copies");
        // ---- dense scoring pass: every scorable node against the removed list of every search that will run
        const bool denseOn = (ctx->denseMode == 1 || (ctx->denseMode < 0 && !sp.strictTopologyStopRules)) && !ctx->treeHasMut &&
                             !sp.deeperSearchForLongBranches && ctx->scanAllStaged;
        if (denseOn) {
            const size_t nN = (size_t)T.nNodes;
            const long long stride = (long long)((ctx->scanNumLists + 31) & ~size_t(31));  // columns <= lists with a copy
            size_t freeB = 0, totalB = 0;
            CK(cudaMemGetInfo(&freeB, &totalB));
            size_t budget = ctx->denseBudget;
            if (ctx->denseBytes == 0 && budget > freeB / 2) budget = freeB / 2;  // first allocation: leave half of what is free
            long long maxRows = stride > 0 ? (long long)(budget / ((size_t)stride * 8)) : 0;
            if (maxRows > n) maxRows = n;
            if (maxRows > 0) {
                const size_t scoresB = ((size_t)maxRows * (size_t)stride * 8 + 255) & ~size_t(255);
                const size_t cB = ((size_t)maxRows * kDenseCUnits * 16 + 255) & ~size_t(255);
                const size_t tabB = (((size_t)n * 4 + (size_t)maxRows * 4 + (size_t)maxRows * 8 + nN * 4) + 255) & ~size_t(255);
                const size_t need = scoresB + cB + tabB + 256;
                if (need > ctx->denseBytes) {
                    cudaFree(ctx->denseMem);
                    ctx->denseMem = nullptr;
                    ctx->denseBytes = 0;
                    if (cudaMalloc(&ctx->denseMem, need) == cudaSuccess) ctx->denseBytes = need;
                    else { ctx->denseMem = nullptr; (void)cudaGetLastError(); }
                }
                if (ctx->denseMem) {
                    char* b = (char*)ctx->denseMem;
                    double* scores = (double*)b;
                    uint4* cArena = (uint4*)(b + scoresB);
                    char* tb = b + scoresB + cB;
                    double* rowBLen = (double*)tb;
                    int32_t* rowOf = (int32_t*)(tb + (size_t)maxRows * 8);
                    int32_t* rowEntry = rowOf + n;
                    int32_t* colPos = rowEntry + maxRows;
                    unsigned long long* counters = (unsigned long long*)(b + scoresB + cB + tabB);  // [0] rows, [1] tasks, [2] (int) columns
                    CK(cudaMemsetAsync(counters, 0, 32, (cudaStream_t)stream));
                    k_dense_cols<<<1, 1024, 0, (cudaStream_t)stream>>>(T.nNodes, ctx->scanRecs, colPos, (int32_t*)(counters + 2));
                    k_dense_prepare<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(ctx->model, T, sp, n, nodes, (int)maxRows, counters, rowOf,
                                                                                                 rowEntry, cArena, rowBLen);
                    timer.mark("dense columns + rows");
                    int densePerSM = 0;
                    const int densePool = 14 * 1024;
                    const size_t denseSmem = (kDenseThreads / 32) * (sizeof(DenseSmem) - sizeof(uint4) + densePool);
                    CK(cudaFuncSetAttribute(k_dense_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)denseSmem));
                    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&densePerSM, k_dense_score, kDenseThreads, denseSmem));
                    if (densePerSM < 1) densePerSM = 1;
                    k_dense_score<<<ctx->numSMs * densePerSM, kDenseThreads, denseSmem, (cudaStream_t)stream>>>(
                        ctx->model, T, densePool, (const int32_t*)(counters + 2), colPos, counters, (int)maxRows, cArena, rowBLen, scores, stride,
                        counters + 1);
                    ctx->launches += 3;
                    timer.mark("dense scoring");
                    ds.scores = scores;
                    ds.rowOf = rowOf;
                    ds.stride = stride;
                }
            }
        }
    } else if ((ctx->searchVariant == 0 || ctx->searchVariant == 3) && T.order && ctx->scanMinSize > 0) {
        k_scan_prepare<<<(T.nNodes + 255) / 256, 256, 0, (cudaStream_t)stream>>>(T, sp.effectivelyNon0BLen, const_cast<ScanNode*>(T.scan));
        ctx->launches++;
    }
    const int scanMin = (T.order && (ctx->searchVariant == 0 || ctx->searchVariant == 3)) ? ctx->scanMinSize : 0;
    const int scanFlags = ctx->searchVariant == 3 ? 3 : ((ctx->scanAppendSitewise ? 0 : 1) | (ctx->scanReplaySequential ? 2 : 0));
    SearchResult* outMain = (SearchResult*)out + nCrit;
    long long* cyclesMain = out_cycles ? (long long*)out_cycles + nCrit : nullptr;
    if (nCrit) {
        // The critical launch goes first, on a stream of its own; a one-thread gate holds this stream until its CTAs are resident.
        const size_t hogSmem = ctx->criticalHogSmem;  // a whole SM's shared memory: nothing else fits next to this CTA
        const unsigned long long firstC = (unsigned long long)nCrit;  // one search per CTA, handed out statically; the counter has nothing more
        CK(cudaMemcpyAsync(ctx->criticalCounter, &firstC, sizeof firstC, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        char* base2 = base + mainBytes;
        const size_t o2 = (size_t)nCrit;
        double* pay2 = (double*)base2;
        double* ais2 = (double*)(base2 + o2 * capP * 8);
        StackE* stack2 = (StackE*)(base2 + o2 * (capP + capA) * 8);
        uint32_t* key2 = (uint32_t*)(base2 + o2 * ((size_t)(capP + capA) * 8 + (size_t)stackCap * sizeof(StackE)));
        BigScratch big2 = big;
        big2.started = ctx->retryCounters + 3;
        EvalScratch es2 = es;
        if (es.key) { es2.pay += (size_t)threads * es.capP; es2.ais += (size_t)threads * es.capA; es2.key += (size_t)threads * es.capK; }
        CK(cudaEventRecord(ctx->criticalEvA, (cudaStream_t)stream));
        CK(cudaStreamWaitEvent(ctx->criticalStream, ctx->criticalEvA, 0));
        fsmKernel<<<nCrit, 32, hogSmem, ctx->criticalStream>>>(ctx->model, T, sp, (int64_t)nCrit, nodesAll, (SearchResult*)out, key2, pay2, ais2, stack2, capK,
                                                                 capP, capA, stackCap, ctx->criticalCounter, (long long*)out_cycles, scanMin, scanFlags,
                                                                 poolBytes, ctx->statsOn ? ctx->searchStats : nullptr, nullptr, nullptr, 1, big2, sq, 0, ds, es2);
        CK(cudaEventRecord(ctx->criticalEvB, ctx->criticalStream));
        k_wait_started<<<1, 1, 0, (cudaStream_t)stream>>>(ctx->retryCounters + 3, (unsigned long long)nCrit);
        ctx->launches += 2;
    }
    if (ctx->searchVariant == 1)
        k_spr_search<<<blocks, kSearchThreads, 0, (cudaStream_t)stream>>>(ctx->model, T, sp, n, nodes, (SearchResult*)out, scrKey, scrPay,
                                                                         scrAis, scrStack, capK, capP, capA, stackCap, ctx->searchCounter,
                                                                         (long long*)out_cycles);
    else
    {
        BigScratch bigMain = big;
        bigMain.headEnd = headEnd;
        fsmKernel<<<blocks, kSearchThreads, fsmSmem, (cudaStream_t)stream>>>(ctx->model, T, sp, n, nodes, outMain, scrKey, scrPay, scrAis, scrStack, capK,
                                                                             capP, capA, stackCap, ctx->searchCounter, cyclesMain, scanMin, scanFlags,
                                                                             poolBytes, ctx->statsOn ? ctx->searchStats : nullptr, nullptr, nullptr, lpw,
                                                                             bigMain, sq, fsmSMs, ds, es);
    }
    if (nCrit) {
        CK(cudaStreamWaitEvent((cudaStream_t)stream, ctx->criticalEvB, 0));
        nodes = nodesAll;
        n = nAll;
    }
    ctx->launches++;
    timer.mark("searches");
    if (ctx->searchVariant != 1 && n < (int64_t(1) << 31)) {
        // safety net: searches that found no large slot free are collected and re-run by one CTA that uses the same slots (free
        // again by then).  With none to re-run the two launches return at once.
        k_collect_overflow<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, (const SearchResult*)out, nodes, retryNodes, retryIdx,
                                                                                         ctx->retryCounters, retryCap);
        BigScratch none{};
        fsmKernel<<<kBigSlots / kSearchThreads, kSearchThreads, fsmSmem, (cudaStream_t)stream>>>(
            ctx->model, T, sp, (int64_t)retryCap, retryNodes, (SearchResult*)out, big.key, big.pay, big.ais, big.stack, big.capK, big.capP, big.capA,
            stackCap, ctx->retryCounters + 1, nullptr, (T.order && (ctx->searchVariant == 0 || ctx->searchVariant == 3)) ? ctx->scanMinSize : 0,
            ctx->searchVariant == 3 ? 3 : ((ctx->scanAppendSitewise ? 0 : 1) | (ctx->scanReplaySequential ? 2 : 0)), poolBytes, nullptr,
            ctx->retryCounters, retryIdx, 32, none, ScanQueue{}, 0, DenseScores{}, EvalScratch{});
        ctx->launches += 2;
        timer.mark("retry net");
    }
    timer.report();
    CK(cudaGetLastError());
    return MAPLE_OK;
}

int maple_place_batch(maple_ctx* ctx, const maple_place_params* p, int64_t n, const int32_t* sampleLists, maple_place_result* out,
                      int32_t scratch_keys_per_sample, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (!ctx->haveTree) { ctx->err = "maple_place_batch: no tree bound (maple_tree_bind)"; return MAPLE_E_STATE; }
    if (n == 0) return MAPLE_OK;
    if (n < 0 || !p || !sampleLists || !out) return MAPLE_E_ARG;
    static_assert(sizeof(maple_place_result) == sizeof(PlaceResult), "placement record layout");
    static_assert(sizeof(maple_place_params) == sizeof(PlaceParams), "placement params layout");
    CK(cudaSetDevice(ctx->device));
    PlaceParams pp;
    memcpy(&pp, p, sizeof pp);
    DevTree T = ctx->tree;
    T.key = ctx->key; T.pay = ctx->pay; T.keyStart = ctx->keyStart; T.payStart = ctx->payStart;
    const unsigned capK = (unsigned)((scratch_keys_per_sample > 0 ? scratch_keys_per_sample : 4096) + 3) & ~3u;
    const unsigned capP = 2 * capK + 6 * 1024, capA = capK / 2 > 2048 ? capK / 2 : 2048;
    const int stackCap = ctx->treeHeight + 8, bestCap = (int)(capK / 4 > 1024 ? capK / 4 : 1024);  // bestNodes entries per sample
    const size_t laneBytes = ((size_t)capP * 8 + (size_t)capA * 8 + (size_t)bestCap * sizeof(PlaceBest) + (size_t)stackCap * sizeof(PlaceStackE) +
                              (size_t)capK * 4 + 15) & ~size_t(15);
    // Variant 1 covers what its scans cover; everything else would run on one lane per warp there, so it goes to variant 0
    const bool warpPlain = ctx->placeVariant == 1 && !ctx->treeHasMut, warpMat = ctx->placeVariant >= 2, warpPar = ctx->placeVariant == 3;
    if ((warpPlain || warpMat) && T.order && !pp.deeperSearchForLongBranches) {
        // one warp per sample: 32 lane slices of scratch (refinement entries run one per lane), the sample list, bestNodes and
        // its refinement results, per-depth states, and the stack of the straight-line fallback
        const unsigned laneK = capK / 8 > 512 ? capK / 8 : 512, laneP = 6 * laneK, laneA = laneK;  // 512 at the default 4096 entries
        const int bestCapW = (int)(capK / 4 > 1024 ? capK / 4 : 1024);
        const size_t warpBytes = ((size_t)32 * laneP * 8 + (size_t)32 * laneA * 8 + (size_t)6 * laneK * 8 + (size_t)bestCapW * sizeof(PlaceEval) +
                                  (size_t)stackCap * sizeof(PlacePath) + (size_t)bestCapW * sizeof(PlaceBest) +
                                  (size_t)stackCap * sizeof(PlaceStackE) + (size_t)32 * laneK * 4 + (size_t)laneK * 4 + (size_t)bestCapW * 4 + 63) & ~size_t(63);
        int perSM = 0;
        if (warpPar) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_place_samples_warp_mat<true>, kPlaceWarpThreads, 0));
        else if (warpMat) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_place_samples_warp_mat<false>, kPlaceWarpThreads, 0));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_place_samples_warp, kPlaceWarpThreads, 0));
        if (perSM < 1) perSM = 1;
        const int warpsPerBlock = kPlaceWarpThreads / 32;
        int64_t nBlocks = (int64_t)ctx->numSMs * perSM;
        if (nBlocks * warpsPerBlock > n) nBlocks = (n + warpsPerBlock - 1) / warpsPerBlock;
        const size_t needW = warpBytes * (size_t)nBlocks * warpsPerBlock + 256;
        if (needW > ctx->placeScratchBytes) {
            cudaFree(ctx->placeScratch);
            ctx->placeScratch = nullptr;
            ctx->placeScratchBytes = 0;
            CK(cudaMalloc(&ctx->placeScratch, needW));
            ctx->placeScratchBytes = needW;
        }
        if (!ctx->searchCounter) CK(cudaMalloc((void**)&ctx->searchCounter, sizeof(unsigned long long)));
        CK(cudaMemsetAsync(ctx->searchCounter, 0, sizeof(unsigned long long), (cudaStream_t)stream));
        k_scan_prepare<<<(T.nNodes + 255) / 256, 256, 0, (cudaStream_t)stream>>>(T, pp.effectivelyNon0BLen, const_cast<ScanNode*>(T.scan));
        ctx->launches++;
        if (warpPar)
            k_place_samples_warp_mat<true><<<(int)nBlocks, kPlaceWarpThreads, 0, (cudaStream_t)stream>>>(
                ctx->model, T, pp, n, sampleLists, (PlaceResult*)out, (char*)ctx->placeScratch, warpBytes, laneK, laneP, laneA, stackCap, bestCapW,
                ctx->searchCounter);
        else if (warpMat)
            k_place_samples_warp_mat<false><<<(int)nBlocks, kPlaceWarpThreads, 0, (cudaStream_t)stream>>>(
                ctx->model, T, pp, n, sampleLists, (PlaceResult*)out, (char*)ctx->placeScratch, warpBytes, laneK, laneP, laneA, stackCap, bestCapW,
                ctx->searchCounter);
        else
            k_place_samples_warp<<<(int)nBlocks, kPlaceWarpThreads, 0, (cudaStream_t)stream>>>(ctx->model, T, pp, n, sampleLists, (PlaceResult*)out,
                                                                                              (char*)ctx->placeScratch, warpBytes, laneK, laneP, laneA,
                                                                                              stackCap, bestCapW, ctx->searchCounter);
        ctx->launches++;
        CK(cudaGetLastError());
        return MAPLE_OK;
    }
    int blocksPerSM = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, k_place_samples, kSearchThreads, 0));
    if (blocksPerSM < 1) blocksPerSM = 1;
    int64_t threads = (int64_t)ctx->numSMs * blocksPerSM * kSearchThreads;
    if (threads > n) threads = n;
    const int blocks = (int)((threads + kSearchThreads - 1) / kSearchThreads);
    const size_t need = laneBytes * (size_t)blocks * kSearchThreads + 256;
    if (need > ctx->placeScratchBytes) {
        cudaFree(ctx->placeScratch);
        ctx->placeScratch = nullptr;
        ctx->placeScratchBytes = 0;
        CK(cudaMalloc(&ctx->placeScratch, need));
        ctx->placeScratchBytes = need;
    }
    if (!ctx->searchCounter) CK(cudaMalloc((void**)&ctx->searchCounter, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(ctx->searchCounter, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    k_place_samples<<<blocks, kSearchThreads, 0, (cudaStream_t)stream>>>(ctx->model, T, pp, n, sampleLists, (PlaceResult*)out,
                                                                        (char*)ctx->placeScratch, laneBytes, capK, capP, capA, stackCap, bestCap,
                                                                        ctx->searchCounter);
    ctx->launches++;
    CK(cudaGetLastError());
    return MAPLE_OK;
}

int maple_ctx_set_place_variant(maple_ctx* ctx, int32_t variant) {
    if (!ctx || variant < 0 || variant > 3) return MAPLE_E_ARG;
    ctx->placeVariant = variant;
    return MAPLE_OK;
}

int maple_ctx_set_search_variant(maple_ctx* ctx, int32_t variant) {
    if (!ctx || variant < 0 || variant > 4) return MAPLE_E_ARG;
    ctx->searchVariant = variant == 4 ? 0 : variant;
    ctx->scanOld = variant == 4 ? true : ctx->scanOldEnv;
    return MAPLE_OK;
}

int maple_ctx_set_lanes_per_warp(maple_ctx* ctx, int32_t lanes) {
    if (!ctx || lanes < 0 || lanes > 32) return MAPLE_E_ARG;
    ctx->lanesPerWarp = lanes;
    return MAPLE_OK;
}

int maple_ctx_set_head_searches(maple_ctx* ctx, int32_t count) {
    if (!ctx || count < 0) return MAPLE_E_ARG;
    ctx->headSearches = count;
    return MAPLE_OK;
}

int maple_ctx_set_critical_searches(maple_ctx* ctx, int32_t count) {
    if (!ctx || count < 0) return MAPLE_E_ARG;
    ctx->criticalSearches = count;
    return MAPLE_OK;
}

int maple_ctx_set_scan_service(maple_ctx* ctx, int32_t fsmSMs) {
    if (!ctx || fsmSMs < -1) return MAPLE_E_ARG;
    ctx->fsmSMs = fsmSMs;
    return MAPLE_OK;
}

static int run_update(maple_ctx* ctx, const maple_tree_rw* rw, int mode, int32_t nEntries, const int32_t* entries, int32_t* out_updates,
                      int32_t* out_status, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (!ctx->haveTree) { ctx->err = "maple_update_partials: no tree bound (maple_tree_bind)"; return MAPLE_E_STATE; }
    if (!rw || !rw->key || !rw->pay || !rw->key_start || !rw->pay_start || !rw->nkeys || !rw->npay || !rw->tails || !rw->dist || !rw->dirty ||
        !out_status || nEntries < 0 || (nEntries > 0 && !entries))
        return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    DevTree T = ctx->tree;
    T.key = rw->key; T.pay = rw->pay; T.keyStart = rw->key_start; T.payStart = rw->pay_start; T.nkeys = rw->nkeys; T.npay = rw->npay;
    T.dist = rw->dist;
    const size_t nN = (size_t)T.nNodes;
    const unsigned capK = 1u << 16, capP = 6u << 16, capA = 1u << 16;
    const int workCap = (int)(4 * nN + 64), walkCap = (int)(nN + 8);
    const size_t need = (size_t)capP * 8 + (size_t)capA * 8 + (size_t)capK * 4 + (size_t)workCap * 8 + (size_t)walkCap * 4 + (size_t)nEntries * 8 + 256;
    if (need > ctx->updateBytes) {
        cudaFree(ctx->updateMem);
        ctx->updateMem = nullptr;
        ctx->updateBytes = 0;
        CK(cudaMalloc(&ctx->updateMem, need));
        ctx->updateBytes = need;
    }
    char* b = (char*)ctx->updateMem;
    double* scrPay = (double*)b;
    double* scrAis = scrPay + capP;
    uint32_t* scrKey = (uint32_t*)(scrAis + capA);
    int32_t* work = (int32_t*)(scrKey + capK);
    int32_t* walk = work + 2 * (size_t)workCap;
    int32_t* dEntries = walk + walkCap;
    int32_t* result = dEntries + 2 * (size_t)nEntries + 2;
    if (nEntries) CK(cudaMemcpyAsync(dEntries, entries, (size_t)nEntries * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    ArenaW a;
    a.key = rw->key; a.pay = rw->pay; a.keyStart = rw->key_start; a.payStart = rw->pay_start; a.nkeys = rw->nkeys; a.npay = rw->npay;
    a.tails = (long long*)rw->tails; a.capK = rw->cap_keys; a.capP = rw->cap_pay;
    k_update_partials<<<1, 32, 0, (cudaStream_t)stream>>>(ctx->model, T, a, rw->dist, rw->dirty, scrKey, scrPay, scrAis, capK, capP, capA, work, workCap,
                                                          walk, walkCap, dEntries, nEntries, mode, result);
    ctx->launches++;
    int32_t h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, result, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    *out_status = h[0];
    if (out_updates) *out_updates = h[1];
    ctx->scan2Ok = false;  // the lists changed: the scan-format copies are sized at maple_tree_bind -- bind again before searching
    return MAPLE_OK;
}

int maple_update_partials(maple_ctx* ctx, const maple_tree_rw* rw, int32_t nEntries, const int32_t* nodeDirection, int32_t* out_status, void* stream) {
    return run_update(ctx, rw, 0, nEntries, nodeDirection, nullptr, out_status, stream);
}

int maple_blen_sweep_sequential(maple_ctx* ctx, const maple_tree_rw* rw, int32_t* out_updates, int32_t* out_status, void* stream) {
    return run_update(ctx, rw, 1, 0, nullptr, out_updates, out_status, stream);
}

int maple_ctx_set_dense_scoring(maple_ctx* ctx, int32_t mode, int64_t maxBytes) {
    if (!ctx || mode < -1 || mode > 1 || maxBytes < 0) return MAPLE_E_ARG;
    ctx->denseMode = mode;
    if (maxBytes > 0) ctx->denseBudget = (size_t)maxBytes;
    return MAPLE_OK;
}

int maple_search_stats(maple_ctx* ctx, int32_t enable, uint64_t* out) {
    if (!ctx) return MAPLE_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->searchStats) {
        CK(cudaMalloc((void**)&ctx->searchStats, kNumSearchStats * 8));
        CK(cudaMemset(ctx->searchStats, 0, kNumSearchStats * 8));
    }
    if (out) {
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(out, ctx->searchStats, kNumSearchStats * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemset(ctx->searchStats, 0, kNumSearchStats * 8));
    }
    ctx->statsOn = enable != 0;
    return MAPLE_OK;
}

int maple_ctx_set_scan_min_size(maple_ctx* ctx, int32_t minNodes) {
    if (!ctx || minNodes < 0) return MAPLE_E_ARG;
    ctx->scanMinSize = minNodes;
    return MAPLE_OK;
}

int64_t maple_launch_count(const maple_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
