/*
 * maple_b200.h -- C ABI of the B200-native SPR-likelihood kernels.
 *
 * The reference (NicolaDM/MAPLE, MAPLEv0.7.5.4.py) has no FFI; its seam for this path is the
 * set of module-level Python functions it already ships across a process boundary
 * (SURVEY.md section 8b).  Each entry point below names the reference function it replaces.
 * A maintainer binds these with ctypes; device buffers are owned by the
 * caller (e.g. torch.cuda tensors) and passed as raw pointers + lengths.
 *
 * Conventions: every function returns 0 on success or a negative MAPLE_E_* code and never
 * throws; maple_last_error() gives a message for the last failure on that context.  A context
 * is thread-compatible, not thread-safe.  `stream` is a cudaStream_t passed as void* (NULL =
 * default stream); launches are asynchronous with respect to the host unless stated.
 *
 * Packed genome lists (maple_b200/genome_list.py; reference tuple spec MAPLEv0.7.5.4.py:378-390):
 *   key stream  uint32 per entry: bits 0-2 type (0-3 ACGT, 4 R, 5 N, 6 O) | 3-4 nLens | 5 flag |
 *               6-7 local-reference nucleotide | 8-31 end position (1-based, inclusive)
 *   pay stream  float64: per entry nLens branch lengths, then the 4-vector of an O entry
 *   list i      starts at key[key_start[i]], pay[pay_start[i]]; key_start[i] < 0 encodes None
 */
#ifndef MAPLE_B200_H
#define MAPLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct maple_ctx maple_ctx;

#define MAPLE_OK 0
#define MAPLE_E_ARG (-1)    /* bad argument */
#define MAPLE_E_CUDA (-2)   /* CUDA runtime error (message in maple_last_error) */
#define MAPLE_E_STATE (-3)  /* model / lists not set */
#define MAPLE_E_NOGPU (-4)  /* no usable CUDA device: there is no CPU fallback */

/* model flags (module globals of the reference: usingErrorRate, errorRateSiteSpecific, useRateVariation) */
#define MAPLE_F_USING_ERROR_RATE 1
#define MAPLE_F_ERROR_SITE_SPECIFIC 2
#define MAPLE_F_RATE_VARIATION 4

/* merge flags (keyword arguments of mergeVectors) */
#define MAPLE_MERGE_UPDOWN 1    /* isUpDown=True */
#define MAPLE_MERGE_RETURN_LK 2 /* returnLK=True */

int maple_version(void);
const char* maple_last_error(const maple_ctx* ctx);

/* Reference-derived constants (MAPLEv0.7.5.4.py:3606-3693): lRef, rootFreqs; flags = MAPLE_F_*.
 * Fails with MAPLE_E_NOGPU when the device is not available. */
int maple_ctx_create(maple_ctx** out, int device, int32_t lRef, const double rootFreqs[4], int32_t flags);
int maple_ctx_destroy(maple_ctx* ctx);

/* The model arrays startTopologyUpdatesParallel receives in its input tuple (:9581):
 * mutMatrixGlobal Q[16] row-major, siteRates[lRef] (mutMatrices[pos] = Q*siteRates[pos], :6367) or
 * NULL, errorRateGlobal, errorRates[lRef] or NULL, cumulativeRate[lRef+1],
 * cumulativeErrorRate[lRef+1] or NULL, totError (:6385/:6390).  HOST pointers; copied. */
int maple_ctx_set_model(maple_ctx* ctx, const double Q[16], const double* siteRates, double errorRate,
                        const double* errorRates, const double* cumulativeRate, const double* cumulativeErrorRate,
                        double totError);

/* thresholdProb (:51), thresholdDiffForUpdate (:61), thresholdFoldChangeUpdate (:62),
 * minBLenSensitivity already multiplied by 1/lRef (:3618). */
int maple_ctx_set_thresholds(maple_ctx* ctx, double thresholdProb, double thresholdDiffForUpdate,
                             double thresholdFoldChangeUpdate, double minBLenSensitivity);

/* Bind the arena the batch calls index into (DEVICE pointers, caller-owned, must outlive the calls). */
int maple_lists_bind(maple_ctx* ctx, const uint32_t* key, const double* pay, const int64_t* key_start,
                     const int64_t* pay_start, int64_t nLists);

/* appendProbNode(probVectP, probVectC, isTipC, bLen) -> float (:6505) for n (P,C) pairs.
 * DEVICE pointers.  out[i] = log-likelihood cost or -inf. */
int maple_append_prob_batch(maple_ctx* ctx, int64_t n, const int32_t* pIdx, const int32_t* cIdx, const uint8_t* isTipC,
                            const double* bLen, double* out, void* stream);

/* Same call with HOST buffers: copies the four argument arrays to the device, runs the kernel,
 * copies the scores back and synchronises.  This is the end-to-end form of the call. */
int maple_append_prob_batch_host(maple_ctx* ctx, int64_t n, const int32_t* pIdx, const int32_t* cIdx,
                                 const uint8_t* isTipC, const double* bLen, double* out);

/* mergeVectors(probVect1,bLen1,fromTip1,probVect2,bLen2,fromTip2,returnLK,isUpDown,numMinor1,numMinor2)
 * (:4446) for n pairs.  flags[i] = MAPLE_MERGE_*.  Result i is written at out_key[out_key_start[i]],
 * out_pay[out_pay_start[i]] (caller sizes the slots: at most nkeys1+nkeys2 keys and 6x that many
 * doubles); out_nkeys/out_npay receive the sizes, out_status 0 = list, 1 = None (:4758, :4812),
 * 2 = likelihood underflow (the reference raises).  numMinor1/2, out_lk may be NULL unless
 * MAPLE_MERGE_RETURN_LK is used.  shorten != 0 applies shorten() (:3721) to each result in place,
 * as the reference's callers do before storing a list (:5542, :6201, :6267). */
int maple_merge_batch(maple_ctx* ctx, int64_t n, const int32_t* idx1, const double* bLen1, const uint8_t* fromTip1,
                      const int32_t* idx2, const double* bLen2, const uint8_t* fromTip2, const uint8_t* flags,
                      const int32_t* numMinor1, const int32_t* numMinor2, uint32_t* out_key, double* out_pay,
                      const int64_t* out_key_start, const int64_t* out_pay_start, int32_t* out_nkeys, int32_t* out_npay,
                      double* out_lk, int32_t* out_status, int32_t shorten, void* stream);

/* estimateBranchLengthWithDerivative(probVectP, probVectC, fromTipC) (:5040) for n pairs.
 * scratch: device doubles, pair i may use scratch[scratch_start[i] ...] with room for
 * nkeys(P)+nkeys(C) values.  out_status 0 = out[i] holds the length, 1 = python False. */
int maple_blen_batch(maple_ctx* ctx, int64_t n, const int32_t* pIdx, const int32_t* cIdx, const uint8_t* fromTipC,
                     double* scratch, const int64_t* scratch_start, double* out, int32_t* out_status, void* stream);

/* areVectorsDifferent(probVect1, probVect2) (:5419); a None second list counts as different. */
int maple_vectors_differ_batch(maple_ctx* ctx, int64_t n, const int32_t* idx1, const int32_t* idx2, uint8_t* out,
                               void* stream);

/* rootVector(probVect, bLen, isFromTip, ...) (:4916) for lists that are expressed relative to the
 * reference genome (no MAT mutations between the node and the root).  Output slots as in
 * maple_merge_batch (at most nkeys keys, 6x doubles); shorten != 0 applies shorten() as :4994 does. */
int maple_root_vector_batch(maple_ctx* ctx, int64_t n, const int32_t* idx, const double* bLen, const uint8_t* isFromTip,
                            uint32_t* out_key, double* out_pay, const int64_t* out_key_start, const int64_t* out_pay_start,
                            int32_t* out_nkeys, int32_t* out_npay, int32_t shorten, void* stream);

/* Tables findProbRoot reads (:4865-4912): cumulativeBases[(lRef+1)*4] (:3640-3647) and, under the error model,
 * rootFreqsLogErrorCumulative[lRef+1] (:6379-6389).  HOST pointers; copied. */
int maple_ctx_set_root_tables(maple_ctx* ctx, const int32_t* cumulativeBases, const double* rootFreqsLogErrorCumulative);

/* findProbRoot(probVect) (:4865) for n lists expressed relative to the reference genome.  DEVICE pointers. */
int maple_prob_root_batch(maple_ctx* ctx, int64_t n, const int32_t* idx, double* out, void* stream);

/* passGenomeListThroughBranch(probVect, mutations[node], dirIsUp) (:3749) for n lists: list idx[i] goes through the MAT
 * mutation list of node mutNode[i] (CSR mutStart / mut triples pos1,upNuc,downNuc as in maple_tree_bind).  Output slots as
 * in maple_merge_batch; a slot needs nkeys + 2*nMutations + 2 keys. */
int maple_pass_branch_batch(maple_ctx* ctx, int64_t n, const int32_t* idx, const int32_t* mutNode, const uint8_t* dirIsUp,
                            const int32_t* mutStart, const int32_t* mut, uint32_t* out_key, double* out_pay, const int64_t* out_key_start,
                            const int64_t* out_pay_start, int32_t* out_nkeys, int32_t* out_npay, void* stream);

/* Gather-copy n whole lists between arenas (DEVICE pointers): list i goes from
 * src_key[src_key_start[i]] / src_pay[src_pay_start[i]] (nkeys[i] keys, npay[i] doubles) to the dst
 * offsets.  Used to append batch results to the resident tree arena (what the reference does by
 * assigning tree.probVect[node] = newList, e.g. :6200, :6275, :6309). */
int maple_lists_copy(maple_ctx* ctx, int64_t n, const uint32_t* src_key, const double* src_pay, const int64_t* src_key_start,
                     const int64_t* src_pay_start, const int32_t* nkeys, const int32_t* npay, uint32_t* dst_key, double* dst_pay,
                     const int64_t* dst_key_start, const int64_t* dst_pay_start, void* stream);

/* ---- device-resident SPR search: the seam the reference ships across its process boundary ----------------
 * startTopologyUpdatesParallel(inputTuple) (:9580-9716): on a frozen tree, for every listed node run
 * appendProbNode(current placement) (:9646) and findBestParentTopology (:6817) and report the proposal. */

/* Stop rules / thresholds of the input tuple and the module globals a forked worker inherits. */
typedef struct {
    int32_t strictTopologyStopRules;    /* inputTuple[3] */
    int32_t allowedFailsTopology;       /* inputTuple[4] */
    int32_t deeperSearchForLongBranches; /* global, --deeperSearchForLongBranches */
    int32_t reserved;
    double thresholdLogLKtopology;      /* inputTuple[5] */
    double thresholdTopologyPlacement;  /* inputTuple[6] */
    double thresholdLogLKoptimizationTopology; /* global; adaptive in the main process (:11770-11773) */
    double thresholdLogLKconsecutivePlacement; /* global (:63) */
    double effectivelyNon0BLen;         /* 1/(10 lRef) (:3615) */
    double BLenThresholdDeeperSearch;   /* (log lRef + 5)/lRef (:3624) */
    double defaultBLen;                 /* --defaultBLen (:88) */
} maple_search_params;

/* One record per searched node. */
typedef struct {
    int32_t placement;   /* proposedMoves entry: re-attachment node, or -1 for no proposal */
    int32_t bestNode;    /* findBestParentTopology's bestNode (-1: search not run) */
    int32_t status;      /* 0 searched; 1 not needed (:9674); 2 aborted where the reference's try/except swallows an
                            exception (:9703); 3 per-search scratch exhausted (re-run with more scratch) */
    int32_t phase1;      /* SPR candidate placements scored by the phase-1 appendProbNode calls (:7011, :7223) */
    double improvement;  /* bestLKdiff - bestCurrentLK (:9701) */
    double bestCurrentLK, bestScore, bLenTop, bLenBottom, bLenAppend; /* bestBranchLengths (:7638) */
} maple_search_result;

/* Tree arrays (DEVICE pointers, caller-owned; node = int index like the reference's Tree, :331-376):
 * up/child0/child1 with -1 for none, dist, isTip = no children and no minorSequences, optional MAT mutation lists
 * as CSR (mutStart[nNodes+1], mut = triples pos1,upNuc,downNuc) or NULL, nkeys[4*nNodes] = entries per list and
 * npay[4*nNodes] = payload doubles per list (npay may be NULL: lists are then never staged in shared memory).
 * The bound arena must hold 4*nNodes lists: id = family*nNodes + node, family 0 probVect, 1 probVectUpRight,
 * 2 probVectUpLeft, 3 probVectTotUp. */
int maple_tree_bind(maple_ctx* ctx, int32_t nNodes, int32_t root, const int32_t* up, const int32_t* child0, const int32_t* child1,
                    const double* dist, const uint8_t* isTip, const int32_t* mutStart, const int32_t* mut, const int32_t* nkeys,
                    const int32_t* npay);

/* Search the n listed nodes (DEVICE int32) on the frozen tree; out = n records (DEVICE).  scratch_keys_per_search:
 * entries of per-search list scratch (0 = default 8192); max_concurrent_searches caps the resident threads (0 = fill
 * the GPU).  A search that exhausts its scratch is re-run on the device by a second small launch with 8x the entries; only if
 * that is not enough either does its record come back with status 3.  Deterministic: a node's record does not depend on which
 * other nodes are in the batch. */
int maple_spr_search_batch(maple_ctx* ctx, const maple_search_params* p, int64_t n, const int32_t* nodes,
                           maple_search_result* out, int32_t scratch_keys_per_search, int32_t max_concurrent_searches,
                           int64_t* out_cycles /* optional DEVICE int64[n]: SM clock cycles each search took; NULL to skip */,
                           void* stream);

/* ---- placement of new samples on a frozen tree ------------------------------------------------------------------------
 * findBestParentForNewSample(tree, root, diffs, sample, computePlacementSupportOnly=False) (:7912-8292) for n samples; the
 * reference's own batch form of this is process_chunk under joblib (:11190-11287).  Globals the function reads: */
typedef struct {
    int32_t strictStopRules;            /* --strictInitialStopRules (:7097-analogue at :8089) */
    int32_t allowedFails;               /* --allowedFails (:54) */
    int32_t deeperSearchForLongBranches;
    int32_t onlyFindIdentical;          /* any error-rate option, --supportFor0Branches or --HnZ: only identical samples are absorbed (:7936) */
    double thresholdLogLK;              /* after the multiplication by log(lRef) (:3609) */
    double thresholdLogLKoptimization;  /* idem (:3611) */
    double thresholdLogLKconsecutivePlacement;
    double effectivelyNon0BLen, BLenThresholdDeeperSearch, oneMutBLen;
} maple_place_params;

typedef struct {
    int32_t bestNode;      /* placement branch (the branch above this node), or the leaf that absorbs the sample */
    int32_t status;        /* 0 placed; 1 absorbed as a minor sequence of leaf bestNode (the reference returns (node, 1.0, None, diffs),
                              :7949/:8002); 2 aborted where the reference would raise; 3 per-sample scratch exhausted */
    int32_t phase1;        /* candidate branches scored in the walk (:8033 / :8050) */
    int32_t missedMinors;  /* leaves strictly less informative than the sample (:7960, :8004) */
    double bestScore, bLenTop, bLenBottom, bLenAppend; /* bestBranchLengths; python False is 0.0 */
} maple_place_result;

/* sampleLists: DEVICE int32[n], ids of the samples' tip genome lists (probVectTerminalNode output, :3882) in the bound arena
 * (ids >= 4*nNodes); the tree must be bound (maple_tree_bind).  out: DEVICE records.  The tree is not modified: a sample that
 * the reference would append to minorSequences is reported with status 1. */
int maple_place_batch(maple_ctx* ctx, const maple_place_params* p, int64_t n, const int32_t* sampleLists, maple_place_result* out,
                      int32_t scratch_keys_per_sample, void* stream);

/* Which kernel maple_place_batch launches (default 3): 0 = one sample per thread, the straight-line walk; 1 = one sample per
 * warp: windows of the pre-order scored one node per lane, leaf comparisons one per lane, refinement entries one per lane
 * (a few rare shapes run the straight-line walk inside that kernel; batches on trees with MAT mutations or with
 * --deeperSearchForLongBranches are launched on variant 0, which is the faster one for them); 2 = variant 1 with MAT trees
 * covered: lane 0 walks the part of the tree above mutation-carrying nodes, every mutation-free subtree is scanned by the warp;
 * 3 = variant 2 with the window bookkeeping in parallel form (prefix maximum + three pointer-jumping passes instead of the
 * one-lane replay; same checks, same status).  Same results; a sample that exhausts its scratch reports status 3 in every variant. */
int maple_ctx_set_place_variant(maple_ctx* ctx, int32_t variant);

/* Which search kernel maple_spr_search_batch launches: 0 (default) = one search per lane as a warp-converged state
 * machine, with subtrees whose lists are all stored ones scanned by the whole warp over scan-format copies of the stored
 * lists (scan2.cuh: bulk-copy staging, precomputed site factors, prefix-form replay); 1 = the straight-line
 * one-search-per-thread kernel; 2 = the state machine without warp scans; 3 = the first form of the warp scans with the
 * queued-site form of appendProbNode and the node-by-node window replay; 4 = the first form of the warp scans (arena lists
 * staged per lane, pointer-jumping replay)  (1-4: kept for A/B measurements and to test the alternative paths; same results). */
int maple_ctx_set_search_variant(maple_ctx* ctx, int32_t variant);

/* Searches a warp of maple_spr_search_batch (variant 0) runs at a time: that many of its lanes own a search each, the subtree scans
 * of all of them are executed by the whole warp one after the other.  0 (default) = chosen per launch from the number of
 * searches (few searches -> few per warp, so that a long search does not share its warp).  Tuning only. */
int maple_ctx_set_lanes_per_warp(maple_ctx* ctx, int32_t lanes);

/* Searches that get an SM each.  A round ends with its longest search, and one search is a dependent chain that runs about 1.6x
 * faster on an SM it does not share (instruction cache, issue slots).  count > 0: the FIRST count entries of the node list of
 * every following maple_spr_search_batch (the caller sorts the list longest-first, e.g. by out_cycles of the previous round) are
 * run by a launch of their own -- one single-warp CTA each that takes a whole SM's shared memory -- next to the usual launch,
 * which gets the other SMs.  Worth it when the longest searches are a large part of the round (a shard of a multi-GPU round);
 * at most half the SMs; ignored when the batch has fewer than 4 * count searches.  0 (default) = off.  Same results. */
int maple_ctx_set_critical_searches(maple_ctx* ctx, int32_t count);

/* Head of the list.  count > 0: the first count entries of the node list of every following maple_spr_search_batch (sorted
 * longest-first by the caller) are handed out one per WARP -- lane 0 of every warp pulls them, the other lanes wait until they are
 * gone -- so that each long search has a warp to itself while it runs and the warps share the long ones out one at a time;
 * afterwards every lane pulls as usual.  (With 28 searches to a warp from the start, the long ones advance at a fraction of the
 * warp's speed until the short ones are gone, and the round ends with whichever warp is left holding the most.)  0 (default) =
 * off.  Same results. */
int maple_ctx_set_head_searches(maple_ctx* ctx, int32_t count);

/* How maple_spr_search_batch (variant 0) divides the GPU: the CTAs on the first fsmSMs SMs own the searches (one per lane, their
 * merges / branch lengths / candidate scores near the pruning point run there) and post every subtree scan in a global-memory
 * slot; the CTAs of all other SMs do nothing but take scans from a ticket ring, run them and hand the results back.  0 (default)
 * = off: every warp scans for its own lanes; -1 = chosen per launch from the stop rules (strict rules: half the SMs own, else a
 * sixth).  Experimental (slower than the default on the measured rounds); results do not depend on it. */
int maple_ctx_set_scan_service(maple_ctx* ctx, int32_t fsmSMs);

/* ---- lists that follow an edit of the tree -----------------------------------------------------------------------------
 * The bound tree and arena, writable: the same DEVICE memory maple_lists_bind / maple_tree_bind were given (key, pay and the four
 * per-list tables, dist), plus tails[2] = entries used in key / doubles used in pay (DEVICE; new lists are appended there and the
 * tables re-pointed, the old lists stay where they are), the capacities of key and pay, and dirty[nNodes] (uint8, DEVICE; the
 * reference's tree.dirty). */
typedef struct {
    uint32_t* key;
    double* pay;
    int64_t* key_start;
    int64_t* pay_start;
    int32_t* nkeys;
    int32_t* npay;
    int64_t* tails;
    int64_t cap_keys, cap_pay;
    double* dist;
    uint8_t* dirty;
} maple_tree_rw;

/* updatePartials(tree, nodeList) (:5479-5815): nodeDirection = HOST int32 pairs (node, direction) in the order of the python
 * list (the LAST pair is taken first, like nodeList.pop()); direction 2 = the change comes from the parent, 0 / 1 = from that
 * child.  Runs the reference's work list as the reference does -- one lane, the same order, updateBLen (:5385) on an
 * inconsistent zero-length branch included -- so the lists, lengths and dirty flags come out as the reference's.
 * out_status (HOST): 0 done; 2 the reference would raise; 3 arena capacity (or scratch) exhausted: restore the tables, dist and
 * tails from a copy taken before the call, grow the arena and call again.  Synchronises.  After it the tree must be bound again
 * (maple_tree_bind) before maple_spr_search_batch / maple_place_batch. */
int maple_update_partials(maple_ctx* ctx, const maple_tree_rw* rw, int32_t nEntries, const int32_t* nodeDirection, int32_t* out_status,
                          void* stream);

/* The loop of traverseTreeToOptimizeBranchLengths(tree, root, fastPass=False) (:8815-8886) below the root's children: every dirty
 * branch re-estimated in the reference's visiting order, each accepted change followed by updatePartials before the next
 * estimate (the reference's default, sequential mode).  The caller does the scan of the root's own two branches first
 * (:8745-8814: maple_merge_batch(returnLK) + maple_prob_root_batch, then maple_update_partials twice).  out_updates (HOST) =
 * lengths changed.  Status and re-binding as for maple_update_partials. */
int maple_blen_sweep_sequential(maple_ctx* ctx, const maple_tree_rw* rw, int32_t* out_updates, int32_t* out_status, void* stream);

/* Dense scoring pass of maple_spr_search_batch (variant 0).  What a candidate branch scores against a pruned subtree
 * (appendProbNode(probVectTotUp[node], removedPartials, ...), :7011/:7223) does not depend on the state of the walk, only whether
 * the walk visits it does; and in a deep round a search visits most of the tree.  So before the searches run, one regular kernel
 * scores every scorable node against the removed list of every search of the batch into a [searches x nodes] matrix of doubles
 * in HBM, and the subtree scans of those searches only do the reference's bookkeeping on scores they read.  Same arithmetic, same
 * bits.  Applies to trees without MAT mutations (the removed list is then the same for a whole search).
 * mode: 0 (default) = never; -1 = when it applies and the stop rules are the non-strict ones of the deep rounds; 1 = whenever it
 * applies.  Experimental: measured slower than scoring in place.  maxBytes: HBM the matrix may take (0 = keep; default 64 GiB, and never more than half of what is free at the first
 * allocation); searches beyond it scan the usual way.  Results do not depend on either. */
int maple_ctx_set_dense_scoring(maple_ctx* ctx, int32_t mode, int64_t maxBytes);

/* Subtrees of at least minNodes nodes are scanned by the whole warp (default 8; 0 = never).  Tuning only: results do not
 * depend on it. */
int maple_ctx_set_scan_min_size(maple_ctx* ctx, int32_t minNodes);

/* Profiling counters of the search kernel (40 uint64, meaning in DESIGN.md / scripts/time_search.py): enable != 0
 * switches collection on for later launches; out != NULL receives and resets the counters (synchronises). */
int maple_search_stats(maple_ctx* ctx, int32_t enable, uint64_t* out);

/* Kernel launches issued by this context so far (bench.py reports it as gpu_launches). */
int64_t maple_launch_count(const maple_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MAPLE_B200_H */
