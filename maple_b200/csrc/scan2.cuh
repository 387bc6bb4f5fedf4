// Subtree scans of the SPR search, second form: scan-format lists, bulk-copy staging, prefix-form replay.
//
// What a scan job is, and why its results equal the reference's walk (MAPLEv0.7.5.4.py:6975-7170), is described above
// warp_scan_job in search_fsm.cuh; this file keeps that contract (same visited set, same counts, same phase-2 queue in
// discovery order, same per-depth hand-down) and changes how the work is laid out:
//
//  * SCAN-FORMAT LISTS.  Once per launch k_scan_build rewrites every stored probVectTotUp list into a walk-friendly
//    copy in a dense arena ordered by pre-order position: 8-byte entries {key, aux} followed by the payload.  aux
//    carries the entry's payload index and a few class bits, and the payload of a nucleotide entry carries
//    mutMatrices[pos][nuc][ref] (Q * siteRate, :6367) so that the commonest site -- a certain nucleotide of the candidate
//    branch against a reference run of the removed subtree, :6729-6742 -- costs one shared-memory load and two multiplies.
//    The removed list (the same for every candidate of a job) gets the same treatment once per job, and there the whole
//    factor min(0.25, Q[ref][c]*rate*(bLen+len)) of a nucleotide against a plain reference run of the candidate (:6657-6663)
//    is precomputed.  One AND of the two aux words classifies a segment: nothing to do / one of the two precomputed cases /
//    general site (the unchanged append_site code, run with the lanes converged).  Same factors, same order of
//    multiplications, same carry-over rule as dev_append: the scores are bit-identical to the lane path's.
//  * STAGING.  The lists of a window's candidates are neighbours in the scan arena, so one cp.async.bulk (UBLKCP) per
//    window brings them to shared memory, completion on an mbarrier; no registers, no per-lane copy loops.
//  * REPLAY.  What a node hands to its children only changes at SCORED nodes, so every node carries the position of its
//    nearest scored proper ancestor (static per tree; k_scan_build).  With it the bookkeeping of a window is: running best =
//    prefix maximum over the (at most 32) scored nodes; failedPasses = one pointer-jumping pass over those <= 32 lanes in
//    registers; stop rule per node; "reached" = no earlier node of the window that does not descend covers me, i.e. an
//    exclusive prefix maximum of (w + size[w]) over non-descending nodes.  The prefix maximum of the scores includes nodes the
//    stop rule prunes; that is exact unless a pruned node holds a new best, which is checked -- such a window is replayed node
//    by node like the reference does.
#pragma once
#include "search.cuh"

namespace maple {

// loads that other SMs' stores must be visible to (scan service): volatile / L1-bypassing on the device
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
__device__ __forceinline__ int ld_volatile_i32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void spin_pause(unsigned ns) { __nanosleep(ns); }
template <class T>
__device__ __forceinline__ T ld_cg(const T* p) {  // ld.global.cg of a 4- or 8-byte object
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "ld_cg: 4 or 8 bytes");
    T v;
    if (sizeof(T) == 4) {
        const unsigned u = __ldcg(reinterpret_cast<const unsigned*>(p));
        memcpy(&v, &u, sizeof v);
    } else {
        const unsigned long long u = __ldcg(reinterpret_cast<const unsigned long long*>(p));
        memcpy(&v, &u, sizeof v);
    }
    return v;
}
#else
inline unsigned long long ld_volatile_u64(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
inline int ld_volatile_i32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
inline void spin_pause(unsigned) { host_yield(); }  // the emulated lane lets the others run
template <class T>
inline T ld_cg(const T* p) { return *p; }
#endif

// ---------------------------------------------------------------------------------------------------------------
// scan-format lists
// entry.x = the arena key (type | nLens<<3 | flag<<5 | nuc<<6 | end<<8); entry.y = aux:
//   bits 0-15  byte offset (from the list's payload base) of the entry's payload [len0][len1](4-vector of an O entry).  An
//              informative entry has one more slot IN FRONT of that, at offset - 8: the candidate side's [g] of a nucleotide
//              entry (mutMatrices[pos][nuc][ref]) or [a] of an O entry; the removed side's [f] of a nucleotide entry or [a] of
//              an O entry -- so every precomputed factor is one load from "payload base - 8 + offset".  A candidate-side O
//              entry below the 0.02 shortcut also carries [q0..q3] = mutMatrices[pos][.][ref] behind its vector (scan_convert_slow_o).
//   bits 16-22 candidate side: one-hot of the entry type; removed side: types of the OTHER list this entry is informative
//              against (append_informative) -- the AND of both is non-zero exactly at the informative segments
//   bit 31     candidate side: plain reference run (type R, no lengths); removed side: nucleotide with at most one length
//              -> the factor is the removed side's precomputed [f]
//   bit 30     candidate side: plain nucleotide (no lengths); removed side: plain reference run
//              -> the factor is min(0.25, [g] * bLen) with the candidate side's [g]
//   bit 29     candidate side: O entry whose probability of the reference nucleotide exceeds 0.02 (the shortcut of :6692);
//              removed side: any reference run -> the factor is that probability, the candidate side's [a]
//   bit 28     candidate side: any reference run; removed side: O entry whose probability of ITS reference nucleotide exceeds
//              0.02 (:6615) -> the factor is that probability, the removed side's [a]
//   bit 27     candidate side: O entry below that shortcut whose [a] slot the job has overwritten with its whole factor against a
//              plain reference run (scan_convert_slow_o, on the job's own staged copy); removed side: plain reference run
//              (type R, no lengths) -> the factor is that slot
//   (under the error model the removed side's [f] carries the error term of :6657 where there is one, a candidate-side
//   nucleotide entry has [e] = 0.33333 * error rate of the site in front of its [g], which the walk adds to the factor of bit 30
//   when the removed node is a tip (:6742), and bit 27 is not set on the removed side then -- those sites take the general code)
constexpr int kMinCarryOverHi = 0x0a711b0e;  // high word of kMinCarryOver = DBL_MIN * 1e50 (checked in scan_walk's host build)
constexpr uint32_t SA_OFF = 0xffffu /* bytes */, SA_TYPES = 0x7f0000u, SA_FAST_C = 0x80000000u, SA_FAST_P = 0x40000000u, SA_FAST_PO = 0x20000000u,
                   SA_FAST_CO = 0x10000000u, SA_FAST_PS = 0x08000000u, SA_FAST = 0xf8000000u;

struct ScanRec {  // one per pre-order position, 32 bytes
    int32_t node;
    int32_t size;     // nodes in the subtree
    int32_t nsa;      // pre-order position of the nearest SCORED proper ancestor, -1 if there is none
    uint32_t depths;  // depth | depth of that ancestor << 16
    uint32_t off;     // scan-format list: offset in the scan arena, 16-byte units (valid with SR_STAGED)
    uint32_t cnt;     // 16-byte units: entries | payload << 16
    uint32_t flags;   // SN_ELIG | SN_TOT | SN_PUSHED | SN_INNER as in ScanNode, plus:
    int32_t col;      // dense scoring pass (k_dense_cols): column of this node's score in a search's row, -1 = none (not scored, or no
                      // copy); otherwise, as k_scan_build leaves it: the list's O entries below the 0.02 shortcut -- their number
                      // (bits 30-31, 3 = three or more) and the entry indices of the first three (10 bits each, 1023 = beyond)
};
constexpr uint32_t SR_STAGED = 16;  // a scan-format copy of probVectTotUp exists
constexpr uint32_t SR_SCORED = 64;  // SN_ELIG && SN_TOT && SN_PUSHED: the walk scores this node when it reaches it

// Candidate-side copy of one stored list.  Returns the payload doubles written.  slowInfo: the list's O entries below the 0.02
// shortcut (without the error model) -- their number in bits 30-31 (3 = more than two) and the entry indices of the first two
// (15 bits each, 0x7fff = beyond).
__device__ inline int scan_build_p(const DevModel& m, const uint32_t* k, const double* p, int nk, uint2* outE, double* outP,
                                   uint32_t* slowInfo = nullptr) {
    int np = 0, ip = 0;
    uint32_t slow = 0, nSlow = 0;
    for (int i = 0; i < nk; i++) {
        const uint32_t key = __ldg(k + i);
        const int type = int(key & 7u), nl = int((key >> 3) & 3u), nuc = int((key >> 6) & 3u), end = int(key >> 8);
        uint32_t aux = 1u << (16 + type);
        bool slowO = false;
        if (type == T_R) aux |= SA_FAST_CO;
        if (type == T_R && nl == 0) aux |= SA_FAST_C;
        if (type < 4 && nl == 0) aux |= SA_FAST_P;
        if (type < 4) {
            const SiteQ q(m, end - 1);
            // under the error model, in front of [g]: the term a removed TIP adds to the factor min(0.25, [g] * bLen) (:6742)
            if (m.U) outP[np++] = (double)(1) * 0.33333 * site_eps(m, end - 1);
            outP[np++] = q.at(type, nuc);  // [g] = mutMatrices[pos][nuc of the entry][reference nuc]
        } else if (type == T_O) {  // [a]: the shortcut's probability, or room for the job's own factor (bit 27)
            const double a = __ldg(p + ip + nl + nuc);
            if (a > 0.02) aux |= SA_FAST_PO;
            else {
                slowO = true;
                if (nSlow < 2) slow |= uint32_t(i < 0x7fff ? i : 0x7fff) << (15 * nSlow);
                nSlow++;
            }
            outP[np++] = a;
        }
        aux |= uint32_t(np) << 3;
        for (int q = 0; q < nl; q++) outP[np++] = __ldg(p + ip + q);
        ip += nl;
        if (type == T_O) {
            // the vector only where the walk itself may need it: an entry above the shortcut is [a] against every reference run,
            // and the rare site where it meets something else reads the vector from the stored list (scan_orig_payload)
            if (!(aux & SA_FAST_PO))
                for (int q = 0; q < 4; q++) outP[np++] = __ldg(p + ip + q);
            ip += 4;
            if (slowO) {
                const SiteQ q(m, end - 1);
                for (int j = 0; j < 4; j++) outP[np++] = q.at(j, nuc);  // what getPartialVec reads for the reference nucleotide (:4110-4141)
            }
        }
        outE[i] = make_uint2(key, aux);
    }
    if (nk & 1) outE[nk] = make_uint2(0u, 0u);
    if (slowInfo) *slowInfo = slow | ((nSlow < 3 ? nSlow : 3u) << 30);
    return np;
}

// Payload of entry idx of a stored (arena) list: [len0][len1][4-vector].
struct ScanOrig {
    const uint32_t* k;
    const double* p;
};
__device__ __noinline__ const double* scan_orig_payload(ScanOrig o, int idx) {
    int ip = 0;
    for (int i = 0; i < idx; i++) {
        const uint32_t key = __ldg(o.k + i);
        ip += int((key >> 3) & 3u) + ((key & 7u) == uint32_t(T_O) ? 4 : 0);
    }
    return o.p + ip;
}

// The whole factor of a candidate-side O entry below the 0.02 shortcut against a plain reference run of the removed list
// (:6692-6703; the same getPartialVec and the same sum as append_site): it depends on the job only through bLen.
// pay = the entry's payload in the scan format: [len0][len1] [vector] [q0..q3].
struct SlotQ {
    const double* g;
    __device__ __forceinline__ double at(int i, int) const { return g[i]; }
};
__device__ __forceinline__ double scan_slow_o_factor(uint32_t key, const double* pay, double bLen) {
    const int nl = int((key >> 3) & 3u), x = int((key >> 6) & 3u);
    double contrib = bLen;
    if (nl == 1) contrib += pay[0];
    const double* a = pay + nl;
    const SlotQ q{pay + nl + 4};
    double t3[4];
    gv_nuc(q, 0.0, x, contrib, false, false, t3);
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) tot += a[j] * t3[j];
    return tot;
}

// Removed-side copy (one per job, shared memory).  Returns the payload doubles written, or -1 if it does not fit.
__device__ inline int scan_build_c(const DevModel& m, const uint32_t* k, const double* p, double bLen, bool isTipC, uint2* outE, int capE, double* outP,
                                   int capP) {
    constexpr unsigned long long INF = append_informative_mask();
    int np = 0, ip = 0;
    for (int i = 0;; i++) {
        if (i >= capE) return -1;
        const uint32_t key = ld_cg(k + i);
        const int type = int(key & 7u), nl = int((key >> 3) & 3u), nuc = int((key >> 6) & 3u), end = int(key >> 8);
        if (np + 8 > capP || (np + 8) * 8 > int(SA_OFF)) return -1;
        uint32_t row = 0;
        for (int t1 = 0; t1 < 7; t1++) row |= uint32_t((INF >> (t1 * 8 + type)) & 1ull) << t1;
        const bool fastNuc = type < 4 && nl <= 1;
        uint32_t aux = row << 16;
        if (fastNuc) aux |= SA_FAST_C;
        if (type == T_R) aux |= SA_FAST_PO;
        // (bLen == 0 included: the factor min(0.25, g * 0) = 0 makes the walk return -inf there, as the reference does, :6663.
        //  Under the error model the candidate side's factors against a plain reference run pick up an error term when the
        //  removed node is a tip (:6729-6742, :6696): the walk adds it for a nucleotide; an O entry takes the general code then.)
        if (type == T_R && nl == 0) aux |= (m.U && isTipC) ? SA_FAST_P : (SA_FAST_P | SA_FAST_PS);
        if (type == T_O) {
            const double a = ld_cg(p + ip + nl + nuc);
            if (a > 0.02) {
                aux |= SA_FAST_CO;
                outP[np++] = a;
            }
        }
        aux |= uint32_t(np + (type < 4 ? 1 : 0)) << 3;
        if (type < 4) {
            const SiteQ q(m, end - 1);
            const double g = q.at(nuc, type);  // mutMatrices[pos][reference nuc][nuc of the entry]
            double contrib = bLen;
            if (nl == 1) contrib += ld_cg(p + ip);
            // the whole factor against a plain reference run of the candidate (:6640-6663); -1 marks "the reference returns -inf"
            const bool flag2 = m.U && (isTipC || (nl > 0 && ((key >> 5) & 1u)));
            double f;
            if (flag2) f = fmin(0.25, g * contrib) + site_eps(m, end - 1) * 0.33333;
            else f = (contrib == 0.0) ? -1.0 : fmin(0.25, g * contrib);
            outP[np++] = f;
            for (int q2 = 0; q2 < nl; q2++) outP[np++] = ld_cg(p + ip + q2);
            ip += nl;
            outP[np++] = g;
        } else {
            for (int q2 = 0; q2 < nl; q2++) outP[np++] = ld_cg(p + ip + q2);
            ip += nl;
            if (type == T_O) {
                for (int q2 = 0; q2 < 4; q2++) outP[np++] = ld_cg(p + ip + q2);
                ip += 4;
            }
        }
        outE[i] = make_uint2(key, aux);
        if (end == m.lRef) return np;
    }
}

// Host builds with -DMAPLE_HOST_STATS (ad-hoc analysis, tests/hostsim): which sites take the general code.
#if defined(MAPLE_HOST_STATS) && !defined(__CUDACC__)
extern "C" { unsigned long long g_scan_hist[16]; }
#define SCAN_HIST(i) (g_scan_hist[i]++)
#else
#define SCAN_HIST(i) ((void)0)
#endif

__device__ __noinline__ double scan_site_general(const DevModel& m, uint32_t k1, const double* pay1, uint32_t k2, const double* pay2, int pos,
                                                 double bLen, bool isTipC, double F) {
#if defined(MAPLE_HOST_STATS) && !defined(__CUDACC__)
    {
        const int t1 = int(k1 & 7u), t2 = int(k2 & 7u), nl1 = int((k1 >> 3) & 3u), nl2 = int((k2 >> 3) & 3u);
        SCAN_HIST(nl1 == 2 ? 6 : (t1 < 4 && t2 < 4) ? 0 : (t1 < 4 && t2 == T_O) ? 1 : (t1 == T_O && t2 < 4) ? 2 : (t1 == T_O && t2 == T_O) ? 3
                  : (t1 == T_R && t2 == T_O) ? 4 : (t1 == T_O && t2 == T_R) ? (nl2 ? 10 : 5) : (t1 == T_R && t2 < 4) ? (nl1 ? 11 : 12)
                  : (t1 < 4 && t2 == T_R) ? (nl2 ? 13 : nl1 ? 14 : 15) : 7);
    }
#endif
    return append_site_ref(m, k1, pay1, k2, pay2, pos, bLen, isTipC, F);
}

// appendProbNode(candidate list, removed list, isTipC, bLen) over the scan-format copies: the arithmetic and its order are
// dev_append's (:6505-6785).  Called by the lanes of a warp together, one candidate each.
// The segment loop is written without branches around its loads: a precomputed factor is fetched (or 1.0 taken) and multiplied
// in every iteration -- x * 1.0 == x exactly -- and a cursor that does not advance re-reads its entry.  A site without a
// precomputed factor takes the general code then and there (a call inside the iteration, after which the lanes go on together:
// a lane that stopped to wait for the others would have to run the rest of its lists on its own afterwards).
// :6772-6783, out of line (it is rare): F has fallen to minimumCarryOver or below.  Returns (F, Lk) after the carry-over; F = -1:
// the reference returns -inf.  (By value: arguments by reference would pin the walk's accumulators in local memory.)
struct ScanCarry {
    double F, Lk;
};
__device__ __noinline__ ScanCarry scan_carry_over(double F, double Lk) {
    if (F <= kMinCarryOver) {
        if (F < DBL_MIN) return ScanCarry{-1.0, Lk};
        Lk += log(F);
        F = 1.0;
    }
    return ScanCarry{F, Lk};
}

template <class GetOrig>
__device__ __forceinline__ double scan_walk(const DevModel& m, const uint2* eP, const double* pP, const uint2* eC, const double* pC, bool isTipC,
                                            double bLen, const double* one /* a 1.0 next to the lists (same memory space) */,
                                            GetOrig getOrig /* () -> ScanOrig: the stored list the candidate side was copied from */) {
    const int lRef = m.lRef;
    const char* const fC = reinterpret_cast<const char*>(pC - 1);  // factor slots: payload base - 8 + offset
    const char* const fP = reinterpret_cast<const char*>(pP - 1);
    int iP = 0, iC = 0;
    uint2 a = eP[0], b = eC[0];
    double F = 1.0;
    double Lk = bLen * (-(double)lRef);
    const bool uTip = m.U && isTipC;
    if (uTip) Lk += m.totError;
    SCAN_HIST(9);
    for (;;) {
        SCAN_HIST(8);
        const uint32_t mm = a.y & b.y;
        const int e1 = int(a.x >> 8), e2 = int(b.x >> 8);
        const int np = min(e1, e2);
        if ((mm & (SA_FAST | SA_TYPES)) && !(mm & SA_FAST)) {  // an informative site without a precomputed factor
            const double* pay1 = pP + ((a.y & SA_OFF) >> 3);
            if (a.y & SA_FAST_PO) pay1 = scan_orig_payload(getOrig(), iP);  // an O entry copied without its vector
            F = scan_site_general(m, a.x, pay1, b.x, pC + ((b.y & SA_OFF) >> 3), np - 1, bLen, isTipC, F);
        } else {
            // the factor, if there is one: the removed side's slot (bits 31, 28) or the candidate side's (30, 29, 27)
            const bool fromC = (mm & (SA_FAST_C | SA_FAST_CO)) != 0;
            const char* src = (fromC ? fC : fP) + ((fromC ? b.y : a.y) & SA_OFF);
            if (!(mm & SA_FAST)) src = reinterpret_cast<const char*>(one);
            double f = *reinterpret_cast<const double*>(src);
            if ((mm & SA_FAST) == SA_FAST_P) {  // min(0.25, [g] * bLen): neither is ever a NaN, so the cap applies from 0.25 up
                f *= bLen;
                if (__double2hiint(f) >= 0x3fd00000) f = 0.25;
                if (uTip) f += *reinterpret_cast<const double*>(src - 8);  // error model, removed tip: + [e]
            }
            F *= f;
        }
        if (np == lRef) break;
        // F <= minimumCarryOver (also catches the -1 marker of an impossible site): screened by the high word
        if (__double2hiint(F) <= kMinCarryOverHi) {
            const ScanCarry c = scan_carry_over(F, Lk);
            if (c.F < 0.0) return -INFINITY;
            F = c.F;
            Lk = c.Lk;
        }
        iP += (e1 == np);
        iC += (e2 == np);
        a = eP[iP];
        b = eC[iC];
    }
    if (!(F > 0.0)) return -INFINITY;
    return Lk + log(F);
}

// The record of pre-order position i (k_scan_build, one thread per position).  nsaOf[] = nearest scored ancestor-or-self per
// position, filled top-down by the caller (see scan_build_all).
__device__ inline uint32_t scan_static_flags(const DevTree& T, double eff, int node) {
    const int up = T.up[node];
    const int64_t nN = T.nNodes;
    uint32_t fl = 0;
    if (up >= 0 && (T.dist[node] > eff || T.up[up] < 0)) fl |= SN_ELIG;
    if (up >= 0 && T.keyStart[(T.child0[up] == node ? 1 : 2) * nN + up] >= 0) fl |= SN_PUSHED;
    if (T.child0[node] >= 0) fl |= SN_INNER;
    if (T.keyStart[3 * nN + node] >= 0) fl |= SN_TOT;
    if ((fl & (SN_ELIG | SN_TOT | SN_PUSHED)) == (SN_ELIG | SN_TOT | SN_PUSHED)) fl |= SR_SCORED;
    return fl;
}

// 16-byte units of the scan-format copy of the probVectTotUp list at pre-order position i: entries | payload << 16
// (0 = no list, or too large to stage).  maple_tree_bind turns these into offsets.
__device__ inline uint32_t scan_count_units(const DevTree& T, int i, bool U /* error model: [e] slots */) {
    const int node = T.order[i];
    if (node < 0 || T.pre[node] != i) return 0;
    const int64_t id = 3 * (int64_t)T.nNodes + node, ks = T.keyStart[id];
    if (ks < 0) return 0;
    const int nk = T.nkeys[id];
    int np = 0, ip = 0;
    for (int q = 0; q < nk; q++) {
        const uint32_t key = __ldg(T.key + ks + q);
        const int type = int(key & 7u);
        const int nl = int((key >> 3) & 3u);
        np += nl + (type < 4 ? (U ? 2 : 1) : 0);
        if (type == T_O) {  // [a] and, below the 0.02 shortcut, the vector and [q0..q3] (sized as without the error model)
            np += 1;
            if (!(__ldg(T.pay + T.payStart[id] + ip + nl + int((key >> 6) & 3u)) > 0.02)) np += 8;
            ip += 4;
        }
        ip += nl;
    }
    const uint32_t ue = uint32_t(nk + 1) >> 1, up = uint32_t(np + 1) >> 1;
    if (nk <= 0) return 0u;
    return (ue < 65536u && up < 65536u && np < 8192) ? (ue | (up << 16)) : ~0u;  // ~0u: too large for a scan-format copy (16-bit byte offsets)
}

// The record of pre-order position i and, where there is one, the scan-format copy of its list (nsa is filled by scan_fill_nsa
// once every record exists).
__device__ inline ScanRec scan_build_rec(const DevModel& m, const DevTree& T, double eff, int i, uint32_t units, uint4* arena) {
    ScanRec r;
    const int node = T.order[i];
    r.node = node; r.size = 1; r.nsa = -1; r.depths = 0; r.off = 0; r.cnt = 0; r.flags = 0; r.col = 0;
    if (node < 0 || T.pre[node] != i) {  // positions past the reachable nodes
        r.node = -1;
        return r;
    }
    r.size = T.size[node];
    r.depths = uint32_t(T.depth[node]) & 0xffffu;
    r.flags = scan_static_flags(T, eff, node);
    const uint32_t off = T.scanOff[i];
    if (off != ~0u && units) {
        const int64_t id = 3 * (int64_t)T.nNodes + node;
        uint4* dst = arena + off;
        uint32_t slow = 0;
        scan_build_p(m, T.key + T.keyStart[id], T.pay + T.payStart[id], T.nkeys[id], reinterpret_cast<uint2*>(dst),
                     reinterpret_cast<double*>(dst + (units & 0xffffu)), &slow);
        r.col = int32_t(slow);
        r.off = off;
        r.cnt = units;
        r.flags |= SR_STAGED;
    }
    return r;
}

__device__ inline void scan_fill_nsa(const DevTree& T, ScanRec* recs, int i) {
    const int node = recs[i].node;
    if (node < 0) return;
    int a = T.up[node];
    while (a >= 0) {
        const int pa = T.pre[a];
        if (recs[pa].flags & SR_SCORED) {
            recs[i].nsa = pa;
            recs[i].depths |= (uint32_t(T.depth[a]) & 0xffffu) << 16;
            return;
        }
        a = T.up[a];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the job
constexpr int kWin2 = 96;    // pre-order positions per window
constexpr int kPath2 = 40;   // per-depth states kept in shared memory (deeper ones in global scratch)

struct PathE2 {  // what a node hands to its children: midProb and failedPasses (:7090-7106)
    double lk;
    int failed, pad;
};

struct ScanJob {  // filled by the lane that owns the search
    int R, pruned, sibling, failed0;
    double best, lastLK0, removedBLen;
    int isRemovedTip, pathCap, qCap, pad;
    const uint32_t* remK;
    const double* remP;
    PathE2* gpath;
    uint32_t* qTop;
    const double* scoreRow;  // dense scoring pass: the search's row of precomputed candidate scores, or nullptr
    // results
    double bestOut;
    int phase1, qN, newBest, err;
    int state, pad2;  // scan service (below): 0 none, 1 posted, 2 served, 3 declined -- the owner's warp runs it itself
};

// ---- scan service: jobs handed from the warps that own searches to warps (on other SMs) that do nothing but scans.
// One slot per owning lane in global memory; a ring of tickets says which slots are waiting.  Producer: fill the slot,
// __threadfence, take a ticket (atomicAdd on tail), publish (ticket+1)<<32 | owner in ring[ticket % cap].  Consumer: take a
// ticket below tail (CAS on head), wait for its ring entry, read the slot with L1-bypassing loads, run the job, write the
// results, __threadfence, state = 2.  Every owner has at most one job outstanding and cap >= 4 * owners, so a ring entry is
// never overwritten before its consumer has read it.
struct ScanQueue {
    unsigned long long* ring;
    unsigned long long* head;
    unsigned long long* tail;
    unsigned long long* doneSearches;  // searches completed by the warps that own them; servers leave when it reaches n
    unsigned long long* ownerCounter;  // owner ids handed to the owning warps
    ScanJob* jobs;
    unsigned cap;       // power of two; 0 = no service: every warp scans for its own lanes
    int maxOwners;
};



struct Scan2Smem {
    double scoreS[32];  // by rank k: score of the k-th scored node of the window
    double bbS[32];     // running best after the k-th scored node
    PathE2 path[kPath2];
    uint32_t info[kWin2];  // by window position: record flags | number of scored nodes before it << 8 | depth below the job's root << 16
    int size[kWin2];
    int failS[32];    // by rank: failedPasses handed down
    uint32_t offS[32], cntS[32];
    int colS[32];     // by rank: column of the node's precomputed score (dense scoring pass), -1 = none
    short nsa[kWin2];  // window position of the nearest scored proper ancestor, or -(path index)-1
    unsigned char slotS[32];  // by rank: window position
    unsigned long long mbar;
    double one;     // 1.0: what scan_walk multiplies by where there is no factor
    ScanJob job;
    uint4 pool[1];  // the removed list's scan-format copy (for the whole job), then the window's lists
};

#ifdef __CUDA_ARCH__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(b)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one lane: arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* b) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
#endif

// ---- dense scoring pass -------------------------------------------------------------------------------------------------
// In a deep round a search visits most of the tree (61 % of all (search, node) pairs at 100 000 sequences), and what a candidate
// scores does not depend on the state of the walk -- only whether it is visited does.  So before the searches run, one regular
// kernel scores EVERY scorable node against the removed list of EVERY search that will run (k_dense_score: each warp keeps 32
// candidate lists in shared memory and sweeps a block of removed lists over them; same scan_walk, so the same bits), into a
// [searches x nodes] matrix of doubles in HBM (43 GB at 100 000 sequences; searches beyond the memory budget simply do not get
// a row).  The subtree scans of a search with a row then only do the bookkeeping: a window's scores are one coalesced read.
// Applies when no list is re-referenced on the way (no MAT mutations), i.e. the removed list is the same for the whole search.
struct DenseScores {
    const double* scores;   // [maxRows][stride]
    const int32_t* rowOf;   // per entry of the node list: its row, -1 = none
    long long stride;       // doubles per row (columns rounded up to 32)
};

constexpr int kDenseCUnits = 64;  // 16-byte units reserved per removed-list copy (1 KB): entry units | total units in the first word pair

struct DenseRowHeader {  // first 16 bytes of a removed-list copy
    int32_t entUnits, units, isTip, pad;
};

// Per entry i of the node list: would startTopologyUpdatesParallel search it (:9646-9674)?  If so, and its removed list has a copy
// that fits, it gets a row: the copy is built at cArena + row * kDenseCUnits.  One thread per entry (k_dense_prepare).
__device__ inline void dense_prepare_entry(const DevModel& m, const DevTree& T, const SearchParams& sp, int64_t i, const int32_t* nodes, int maxRows,
                                           unsigned long long* rowCounter, int32_t* rowOf, int32_t* rowEntry, uint4* cArena, double* rowBLen) {
    rowOf[i] = -1;
    const int node = nodes[i];
    if (T.up[node] < 0) return;
    const int parent = T.up[node];
    const LRef vectUp = (T.child0[parent] == node) ? tree_list(T, 1, parent) : tree_list(T, 2, parent);
    const LRef own = tree_list(T, 0, node);
    if (!vectUp.k || !own.k) return;
    const double bestCurrentLK = dev_append<true>(m, vectUp.k, vectUp.p, own.k, own.p, T.isTip[node] != 0, T.dist[node]);
    if (!(bestCurrentLK < sp.thresholdTopologyPlacement || T.dist[node] != 0.0)) return;  // :9674
    if (own.nk > 2 * (kDenseCUnits - 1) - 8) return;  // too long for the 1 KB slot: this search scans the usual way
    const unsigned long long row = atomicAdd(rowCounter, 1ULL);
    if (row >= (unsigned long long)maxRows) return;
    uint4* slot = cArena + row * (size_t)kDenseCUnits;
    const int entUnits = (own.nk + 1) >> 1;
    const int capP = (kDenseCUnits - 1 - entUnits) * 2;
    const int npC = scan_build_c(m, own.k, own.p, T.dist[node], T.isTip[node] != 0, reinterpret_cast<uint2*>(slot + 1), own.nk,
                                 reinterpret_cast<double*>(slot + 1 + entUnits), capP);
    DenseRowHeader h;
    h.entUnits = entUnits; h.units = npC < 0 ? 0 : entUnits + ((npC + 1) >> 1); h.isTip = T.isTip[node] != 0; h.pad = 0;
    *reinterpret_cast<DenseRowHeader*>(slot) = h;
    rowBLen[row] = T.dist[node];
    rowEntry[row] = int32_t(i);
    if (npC >= 0) rowOf[i] = int32_t(row);  // (a copy that does not fit leaves its row unused)
}

constexpr int kDenseCBlock = 128;  // removed lists swept over a tile of candidates per task

struct DenseSmem {          // per warp
    uint4 cBuf[2][kDenseCUnits];
    unsigned long long mbar;
    double one;             // 1.0 (scan_walk)
    uint4 pool[1];          // the tile's candidate lists
};

// One task of k_dense_score: tile `tile` (32 consecutive columns) against rows [row0, row1).  colPos[c] = pre-order position of
// column c.  Whole warp.
__device__ inline void dense_score_task(const DevModel& m, const DevTree& t, DenseSmem& W, int poolBytes, uint32_t& mbarParity, int tile, int nCols,
                                        const int32_t* colPos, int row0, int row1, const uint4* cArena, const double* rowBLen, double* scores,
                                        long long stride) {
    const unsigned FULL = 0xffffffffu;
    const int lane = int(threadIdx.x & 31);
    const int col = tile * 32 + lane;
    const bool have = col < nCols;
    uint32_t off = 0, cnt = 0;
    if (have) {
        const ScanRec* r = t.scan2 + colPos[col];
        off = r->off;
        cnt = r->cnt;
    }
    // stage the tile's lists: they are neighbours in the arena; what does not fit the pool is read where it lies
    const uint32_t off0 = __shfl_sync(FULL, off, 0);
    const uint32_t myEnd = off + (cnt & 0xffffu) + (cnt >> 16) - off0;
    const unsigned fits = __ballot_sync(FULL, have && myEnd <= uint32_t(poolBytes >> 4));
    const int nFit = fits == FULL ? 32 : __ffs(~fits) - 1;
    const uint4* base = t.scanArena + off;
    __syncwarp();
#ifdef __CUDA_ARCH__
    if (nFit > 0) {
        const uint32_t endLast = __shfl_sync(FULL, off + (cnt & 0xffffu) + (cnt >> 16), nFit - 1);
        if (lane == 0) bulk_load(W.pool, t.scanArena + off0, (endLast - off0) << 4, &W.mbar);
        mbar_wait(&W.mbar, mbarParity);
        mbarParity ^= 1u;
        if (lane < nFit) base = W.pool + (off - off0);
    }
#endif
    const uint2* eP = reinterpret_cast<const uint2*>(base);
    const double* pP = reinterpret_cast<const double*>(base + (cnt & 0xffffu));
    if (lane == 0) W.one = 1.0;
    // sweep the block of removed lists: copy k+1 lands in the other buffer while copy k is walked
    auto load_c = [&](int row, int b) {
        const uint4* src = cArena + (size_t)row * kDenseCUnits;
        for (int u = lane; u < kDenseCUnits; u += 32) W.cBuf[b][u] = src[u];  // (a 1 KB slot: two 16-byte loads per lane)
    };
    load_c(row0, 0);
    __syncwarp();
    for (int row = row0; row < row1; row++) {
        const int b = (row - row0) & 1;
        if (row + 1 < row1) load_c(row + 1, b ^ 1);
        const DenseRowHeader h = *reinterpret_cast<const DenseRowHeader*>(&W.cBuf[b][0]);
        if (h.units > 0 && have) {
            const uint2* eC = reinterpret_cast<const uint2*>(&W.cBuf[b][1]);
            const double* pC = reinterpret_cast<const double*>(&W.cBuf[b][1 + h.entUnits]);
            const double sc = scan_walk(m, eP, pP, eC, pC, h.isTip != 0, rowBLen[row], &W.one, [&]() {
                const int64_t id = 3 * (int64_t)t.nNodes + __ldg(&t.scan2[colPos[col]].node);
                return ScanOrig{t.key + t.keyStart[id], t.pay + t.payStart[id]};
            });
            scores[(size_t)row * stride + col] = sc;
        }
        __syncwarp();
    }
}

__device__ __forceinline__ double warp_max_incl_d(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double x = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = fmax(v, x);
    }
    return v;
}
__device__ __forceinline__ int warp_max_incl_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = max(v, x);
    }
    return v;
}

__device__ __noinline__ double scan_append_generic(const DevModel& m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC,
                                                   bool isTipC, double bLen) {
    return dev_append_sitewise<true>(m, kP, pP, kC, pC, isTipC, bLen);
}

// W.job holds the request (written by the owning lane, visible to the warp); the results are left in W.job.
// mbarParity: phase parity of W.mbar, kept by the caller across jobs.  st: optional profiling counters (lane 0 adds).
// remote: the job belongs to a lane of another warp (scan service): if the removed list's copy does not fit the pool the job is
// declined (W.job.err = 4) instead of being scored from the arena lists, whose stale copies this SM's L1 might hold.
// EXTRAS = false compiles the scan service and the dense scoring pass out (the default kernel: they are off by default, and their
// code costs the hot paths registers).
template <bool EXTRAS>
__device__ void warp_scan_job2(const DevModel& m, const DevTree& t, const SearchParams& sp, Scan2Smem& W, int poolBytes, int scanFlags /* 2: every window replayed node by node */,
                               uint32_t& mbarParity, unsigned long long* st, bool remoteArg) {
    const bool remote = EXTRAS && remoteArg;
    const unsigned FULL = 0xffffffffu;
    const int lane = int(threadIdx.x & 31);
    const unsigned ltMask = (1u << lane) - 1u;
    const ScanJob J = W.job;
    const int R = J.R;
    double best = J.best;
    const bool isRemovedTip = J.isRemovedTip != 0;
    const double removedBLen = J.removedBLen;
    PathE2* const gpath = J.gpath;
    const int pathCap = J.pathCap, qCap = J.qCap;
    uint32_t* const qTop = J.qTop;
    int phase1 = 0, qN = 0, newBest = 0, err = 0;
    int pos = t.pre[R];
    const int end = pos + t.size[R], d0 = t.depth[R];
    long long tk = st ? clock64() : 0;
    if (st && lane == 0) { st[17] += 1; st[18] += (unsigned long long)(end - pos); }
    // ---- the removed list in scan format, at the front of the pool: the same for every candidate of the job
    const double* const scoreRow = EXTRAS ? J.scoreRow : nullptr;  // dense scoring pass: the scores are there already, this job only keeps the books
    int cEntUnits = 0, cUnits = 0;  // 16-byte units of its entries / of the whole copy; cUnits == 0: it does not fit (at most half the pool)
    if (lane == 0 && !scoreRow) {
        int nkC = 0;
        const int capE = poolBytes >> 4;  // half the pool at 8 bytes per entry
        while (nkC < capE && int(ld_cg(J.remK + nkC) >> 8) != m.lRef) nkC++;
        nkC++;
        cEntUnits = (nkC + 1) >> 1;
        const int capP = ((poolBytes >> 1) - 16 * cEntUnits) >> 3;
        if (nkC <= capE && capP >= 8) {
            const int npC = scan_build_c(m, J.remK, J.remP, removedBLen, isRemovedTip, reinterpret_cast<uint2*>(W.pool), nkC,
                                         reinterpret_cast<double*>(W.pool + cEntUnits), capP);
            if (npC >= 0) cUnits = cEntUnits + ((npC + 1) >> 1);
        }
    }
    if (lane == 0) {
        W.one = 1.0;
        if (pathCap < 2) err = 3;
        else W.path[0] = PathE2{J.lastLK0, J.failed0, 0};
    }
    cEntUnits = __shfl_sync(FULL, cEntUnits, 0);
    cUnits = __shfl_sync(FULL, cUnits, 0);
    const bool cOk = cUnits > 0;
    err = __shfl_sync(FULL, err, 0);
    if (remote && !cOk && !scoreRow && !err) err = 4;
    __syncwarp();
    const uint4* const arena = t.scanArena;
    const int poolUnits = (poolBytes >> 4) - cUnits;
    uint4* const winPool = W.pool + cUnits;
    const uint2* const cEnt = reinterpret_cast<const uint2*>(W.pool);
    const double* const cPay = reinterpret_cast<const double*>(W.pool + cEntUnits);
    while (pos < end && !err) {
        // ---- window: positions pos .. pos+nWin-1, at most 32 of them scored, their lists within the pool
        int nWin = 0, nScore = 0;
        bool generic = !cOk;  // some list has no scan-format copy: the whole window is scored from the arena lists
        const bool dense = scoreRow != nullptr;  // scores read from the search's row; a node without a column is scored from the arena lists
        bool dyn = false;
        for (int sweep = 0; sweep < kWin2 / 32 && pos + nWin < end && nScore < 32; sweep++) {
            const int w = nWin + lane, idx = pos + w;
            uint32_t flags = 0, off = 0, cnt = 0;
            int size = 1, nsaCode = -1, rel = 0, col = -1;
            if (idx < end) {
                const uint4* src4 = reinterpret_cast<const uint4*>(t.scan2 + idx);
                const uint4 a4 = __ldg(src4), b4 = __ldg(src4 + 1);
                const int node = int(a4.x), nsa = int(a4.z);
                size = int(a4.y);
                const int depth = int(a4.w & 0xffffu), nsaDepth = int(a4.w >> 16);
                off = b4.x; cnt = b4.y; flags = b4.z; col = int(b4.w);
                // Two per-search exceptions to the static flags.  Neither occurs on the walks the state machine hands over (the
                // children of the pruned node's parent are never inside a scanned subtree, and a job's root was pushed through an
                // existing upper list), but if one did, the static ancestor links would be off: such a window is replayed node by node.
                if (node == J.pruned || node == J.sibling) {  // children of the pruned node's parent are not scored (:6978)
                    if (flags & SR_SCORED) dyn = true;
                    flags &= ~(SN_ELIG | SR_SCORED);
                }
                if (node == R && !(flags & SN_PUSHED)) {
                    flags |= SN_PUSHED;
                    if ((flags & (SN_ELIG | SN_TOT)) == (SN_ELIG | SN_TOT)) { flags |= SR_SCORED; dyn = true; }
                }
                rel = depth - d0;
                if (nsa >= pos) nsaCode = nsa - pos;
                else {
                    const int pi = (nsa >= 0 && nsaDepth >= d0) ? nsaDepth - d0 + 1 : 0;
                    nsaCode = -pi - 1;
                }
            }
            const unsigned need = __ballot_sync(FULL, (flags & SR_SCORED) != 0);
            const int room = 32 - nScore;
            int take = min(32, end - pos - nWin);
            if (__popc(need) > room) take = __fns(need, 0, room) + 1;  // cut right after the node that fills the last lane
            const unsigned mine = need & (take >= 32 ? FULL : ((1u << take) - 1u));
            const int rank = nScore + __popc(mine & ltMask);
            if (lane < take) {
                W.info[w] = (flags & 0xffu) | (uint32_t(rank) << 8) | (uint32_t(rel) << 16);
                W.size[w] = size;
                W.nsa[w] = short(nsaCode);
                if (flags & SR_SCORED) {
                    W.slotS[rank] = (unsigned char)w;
                    W.offS[rank] = off;
                    W.cntS[rank] = cnt;
                    W.colS[rank] = col;
                }
            }
            if (__any_sync(FULL, lane < take && (flags & SR_SCORED) && !(flags & SR_STAGED))) generic = true;
            nScore += __popc(mine);
            nWin += take;
        }
        __syncwarp();
        // cut the window where its lists stop fitting the pool
        uint32_t myOff = 0, myCnt = 0, off0 = 0;
        if (!dense && !generic && nScore > 0) {
            if (lane < nScore) { myOff = W.offS[lane]; myCnt = W.cntS[lane]; }
            off0 = __shfl_sync(FULL, myOff, 0);
            const uint32_t myEnd = myOff + (myCnt & 0xffffu) + (myCnt >> 16) - off0;
            const unsigned fits = __ballot_sync(FULL, lane < nScore && myEnd <= uint32_t(poolUnits));
            const int nFit = fits == FULL ? 32 : __ffs(~fits) - 1;  // leading run of lists that fit (offsets grow with the position)
            if (nFit == 0) generic = true;
            else if (nFit < nScore) {
                nWin = W.slotS[nFit];
                nScore = nFit;
            }
        }
        // ---- stage the lists: one bulk copy, completion on the mbarrier
        double sc = -INFINITY;
        if (!dense && !generic && nScore > 0) {
            const uint32_t endLast = __shfl_sync(FULL, myOff + (myCnt & 0xffffu) + (myCnt >> 16), nScore - 1);
#ifdef __CUDA_ARCH__
            if (lane == 0) bulk_load(winPool, arena + off0, (endLast - off0) << 4, &W.mbar);
            mbar_wait(&W.mbar, mbarParity);
            mbarParity ^= 1u;
#else
            if (lane == 0)
                for (uint32_t u = 0; u < endLast - off0; u++) winPool[u] = arena[off0 + u];
            __syncwarp();
#endif
            // The staged copies are this job's own: the O entries of the candidates that are below the 0.02 shortcut (the first
            // two of a list) get their whole factor against a plain reference run of the removed list written into their [a]
            // slot (bit 27), so that the walk below finds a precomputed factor there as well -- most of the sites that would
            // take the general code are of this kind.  One entry per lane, whichever candidate it belongs to.
            if (!EXTRAS && !(m.U && isRemovedTip)) {
                const uint32_t slow = lane < nScore ? uint32_t(W.colS[lane]) : 0u;
                const int nSlow = min(int(slow >> 30), 2);
                int before = nSlow;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int x = __shfl_up_sync(FULL, before, o);
                    if (lane >= o) before += x;
                }
                const int total = __shfl_sync(FULL, before, 31);
                if (total > 0) {
                    before -= nSlow;
                    uint32_t* const work = reinterpret_cast<uint32_t*>(W.scoreS);  // (free until this window's scores are in) 64 items
                    const uint32_t listOff = (myOff - off0) << 4, payOff = listOff + ((myCnt & 0xffffu) << 4);  // bytes from winPool
                    for (int r = 0; r < nSlow; r++) {
                        const uint32_t idx = (slow >> (15 * r)) & 0x7fffu;
                        work[before + r] = idx == 0x7fffu ? ~0u : ((listOff + idx * 8u) | (payOff << 16));
                    }
                    __syncwarp();
                    char* const wp = reinterpret_cast<char*>(winPool);
                    for (int t = lane; t < total; t += 32) {
                        const uint32_t it = work[t];
                        if (it == ~0u) continue;
                        uint2* const e = reinterpret_cast<uint2*>(wp + (it & 0xffffu));
                        const uint2 ev = *e;
                        double* const pay = reinterpret_cast<double*>(wp + (it >> 16) + (ev.y & SA_OFF));
                        pay[-1] = scan_slow_o_factor(ev.x, pay, removedBLen);
                        e->y = ev.y | SA_FAST_PS;
                    }
                    __syncwarp();
                }
            }
        }
        if (st) {
            const long long now = clock64();
            if (lane == 0) { st[23] += (unsigned long long)(now - tk); st[19] += 1; st[20] += nScore; st[24] += nWin; }
            tk = now;
        }
        if (lane < nScore && dense && W.colS[lane] >= 0) sc = __ldg(scoreRow + W.colS[lane]);
        else if (lane < nScore) {
            if (!generic && !dense) {
                const uint4* base = winPool + (myOff - off0);
                const uint2* eP = reinterpret_cast<const uint2*>(base);
                const double* pP = reinterpret_cast<const double*>(base + (myCnt & 0xffffu));
                sc = scan_walk(m, eP, pP, cEnt, cPay, isRemovedTip, removedBLen, &W.one, [&]() {
                    const int64_t id = 3 * (int64_t)t.nNodes + __ldg(&t.scan2[pos + W.slotS[lane]].node);
                    return ScanOrig{t.key + t.keyStart[id], t.pay + t.payStart[id]};
                });
            } else {
                const int64_t id = 3 * (int64_t)t.nNodes + __ldg(&t.scan2[pos + W.slotS[lane]].node);
                sc = scan_append_generic(m, t.key + t.keyStart[id], t.pay + t.payStart[id], J.remK, J.remP, isRemovedTip, removedBLen);
            }
        }
        __syncwarp();
        if (st) {
            const long long now = clock64();
            if (lane == 0) st[6] += (unsigned long long)(now - tk);
            tk = now;
        }
        // ---- replay of the reference's bookkeeping over the window, prefix form
        // running best after each scored node (lane k = k-th scored node)
        const double bbIn = fmax(best, warp_max_incl_d(sc, lane));
        double bbEx = __shfl_up_sync(FULL, bbIn, 1);
        if (lane == 0) bbEx = best;
        // failedPasses handed down by each scored node: SET 0 on a new best, else inherited + (1 on a consecutive worsening)
        int fval = 0, ptr = -1;
        {
            double lkIn = 0.0;
            int code = -1;
            if (lane < nScore) code = W.nsa[W.slotS[lane]];
            const int anc = code >= 0 ? (W.info[code] >> 8) & 0xff : 0;  // rank of the scored ancestor inside the window
            const double ancScore = __shfl_sync(FULL, sc, anc);
            if (lane < nScore) {
                int failedIn = 0;
                if (code >= 0) lkIn = ancScore;
                else {
                    const int pi = -code - 1;
                    const PathE2 pe = pi < kPath2 ? W.path[pi] : gpath[pi];
                    lkIn = pe.lk;
                    failedIn = pe.failed;
                }
                const bool nb = sc > bbEx;
                if (!nb) {
                    fval = (sc < (lkIn - sp.thresholdLogLKconsecutivePlacement)) ? 1 : 0;
                    if (code >= 0) ptr = anc;
                    else fval += failedIn;
                }
            }
            while (__any_sync(FULL, ptr >= 0)) {
                const int q = ptr >= 0 ? ptr : 0;
                const int pv = __shfl_sync(FULL, fval, q), pp = __shfl_sync(FULL, ptr, q);
                if (ptr >= 0) { fval += pv; ptr = pp; }
            }
        }
        if (lane < nScore) { W.scoreS[lane] = sc; W.bbS[lane] = bbIn; W.failS[lane] = fval; }
        __syncwarp();
        // stop rule of every node as if it were reached; a node that does not descend hides the rest of its subtree
        constexpr int NC = kWin2 / 32;
        double midc[NC];
        int failc[NC], skipc[NC];
        uint32_t infc[NC];
        bool scoredc[NC], nbc[NC], quec[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const int w = c * 32 + lane;
            midc[c] = 0.0; failc[c] = 0; infc[c] = 0; skipc[c] = 0;
            scoredc[c] = nbc[c] = quec[c] = false;
            if (w < nWin) {
                const uint32_t inf = W.info[w];
                infc[c] = inf;
                const int r = int((inf >> 8) & 0xffu);
                const double before = r == 0 ? best : W.bbS[r - 1];
                const bool scored = (inf & SR_SCORED) != 0;
                double bestAfter = before;
                if (scored) {
                    midc[c] = W.scoreS[r];
                    failc[c] = W.failS[r];
                    bestAfter = W.bbS[r];
                    nbc[c] = midc[c] > before;
                    quec[c] = midc[c] > before - sp.thresholdLogLKoptimizationTopology;  // :7071
                } else {
                    const int code = W.nsa[w];
                    if (code >= 0) {
                        const int q = (W.info[code] >> 8) & 0xff;
                        midc[c] = W.scoreS[q];
                        failc[c] = W.failS[q];
                    } else {
                        const int pi = -code - 1;
                        const PathE2 pe = pi < kPath2 ? W.path[pi] : gpath[pi];
                        midc[c] = pe.lk;
                        failc[c] = pe.failed;
                    }
                }
                scoredc[c] = scored;
                const bool within = midc[c] > (bestAfter - sp.thresholdLogLKtopology);
                const bool rule = sp.strictTopologyStopRules ? (failc[c] <= sp.allowedFailsTopology && within)
                                                             : (failc[c] <= sp.allowedFailsTopology || within);
                const bool dead = (inf & (SN_ELIG | SN_TOT)) == SN_ELIG;  // eligible but no probVectTotUp: the walk moves on (:6999)
                const bool descend = (inf & SN_PUSHED) && !dead && (inf & SN_INNER) && rule;
                skipc[c] = descend ? 0 : w + W.size[w];
            }
        }
        // reached = no earlier node of the window that does not descend covers me
        int maxTarget = 0, nCounted = 0;
        bool bad = false, anyNb = false;
        {
            int carry = 0;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int w = c * 32 + lane;
                const int incl = max(carry, warp_max_incl_i(skipc[c], lane));
                int excl = __shfl_up_sync(FULL, incl, 1);
                if (lane == 0) excl = carry;
                carry = __shfl_sync(FULL, incl, 31);
                const bool reached = w < nWin && excl <= w;
                if (w < nWin) {
                    if (reached) {
                        maxTarget = max(maxTarget, skipc[c] ? skipc[c] : w + 1);
                        if (scoredc[c]) { nCounted++; anyNb |= nbc[c]; }
                    } else {
                        if (scoredc[c] && nbc[c]) bad = true;  // a pruned node would have raised the prefix maximum
                        quec[c] = false;
                        skipc[c] = -1;  // not reached
                    }
                }
            }
        }
        int j = 0;
        if (!__any_sync(FULL, bad || dyn) && !(scanFlags & 2)) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                maxTarget = max(maxTarget, __shfl_xor_sync(FULL, maxTarget, o));
                nCounted += __shfl_xor_sync(FULL, nCounted, o);
            }
            j = maxTarget;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int w = c * 32 + lane;
                const unsigned qm = __ballot_sync(FULL, quec[c]);
                if (quec[c]) {
                    const int at = qN + __popc(qm & ltMask);
                    if (at < qCap) qTop[-1 - at] = uint32_t(__ldg(&t.scan2[pos + w].node));
                }
                qN += __popc(qm);
                // a reached node that descends and whose subtree goes on past this window leaves its hand-down for the windows
                // that follow (one such node per depth: they are the ancestors of the next window's first node)
                if (w < nWin && skipc[c] == 0 && w + W.size[w] > j) {
                    const int rel1 = int(infc[c] >> 16) + 1;
                    if (rel1 >= pathCap) err = 3;
                    else if (rel1 < kPath2) W.path[rel1] = PathE2{midc[c], failc[c], 0};
                    else gpath[rel1] = PathE2{midc[c], failc[c], 0};
                }
            }
            if (qN > qCap) err = 3;
            err = __any_sync(FULL, err == 3) ? 3 : err;
            if (__any_sync(FULL, anyNb)) newBest = 1;
            if (nScore > 0) best = __shfl_sync(FULL, bbIn, nScore - 1);
            phase1 += nCounted;
        } else {
            // node-by-node replay, as the reference walks (every lane runs it redundantly on the shared arrays)
            if (st && lane == 0) st[25] += 1;
            while (j < nWin) {
                const uint32_t inf = W.info[j];
                const int rel = int(inf >> 16), sz = W.size[j];
                const bool scored = (inf & SR_SCORED) != 0;
                PathE2 pe;
                if (rel < kPath2) pe = W.path[rel];
                else pe = gpath[rel];
                double midProb = pe.lk;
                int failed = pe.failed;
                bool alive = true, descend = false;
                if (inf & SN_PUSHED) {
                    if (inf & SN_ELIG) {
                        if (!(inf & SN_TOT)) alive = false;
                        else if (scored) {
                            midProb = W.scoreS[(inf >> 8) & 0xff];
                            phase1++;
                            if (midProb > best - sp.thresholdLogLKoptimizationTopology) {  // :7071
                                if (qN >= qCap) err = 3;
                                else if (lane == 0) qTop[-1 - qN] = uint32_t(__ldg(&t.scan2[pos + j].node));
                                qN++;
                            }
                            if (midProb > best) { best = midProb; failed = 0; newBest = 1; }
                            else if (midProb < (pe.lk - sp.thresholdLogLKconsecutivePlacement)) failed++;
                        }
                    }
                    if (alive && (inf & SN_INNER)) {
                        if (sp.strictTopologyStopRules) descend = failed <= sp.allowedFailsTopology && midProb > (best - sp.thresholdLogLKtopology);
                        else descend = failed <= sp.allowedFailsTopology || midProb > (best - sp.thresholdLogLKtopology);
                        if (descend) {
                            if (rel + 1 >= pathCap) { err = 3; descend = false; }
                            else if (rel + 1 < kPath2) { if (lane == 0) W.path[rel + 1] = PathE2{midProb, failed, 0}; }
                            else if (lane == 0) gpath[rel + 1] = PathE2{midProb, failed, 0};
                        }
                    }
                }
                __syncwarp();
                j += descend ? 1 : sz;
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
    if (lane == 0) {
        W.job.bestOut = best;
        W.job.phase1 = phase1;
        W.job.qN = qN;
        W.job.newBest = newBest;
        W.job.err = err;
    }
    __syncwarp();
}

// The loop of a serving warp: take a posted job, run it, hand the results back; leave when every search of the launch has been
// completed by the warp that owns it (then nothing can be posted any more) and the ring is empty.
__device__ void scan_server_loop(const DevModel& m, const DevTree& t, const SearchParams& sp, Scan2Smem& W, int poolBytes, int scanFlags,
                                 uint32_t& mbarParity, unsigned long long* st, const ScanQueue& sq, int64_t n) {
    const unsigned FULL = 0xffffffffu;
    const int lane = int(threadIdx.x & 31);
    unsigned idleNs = 500;  // an idle server backs off (up to 8 us between looks at the ring): thousands of warps poll two words
    long long tk = st ? clock64() : 0;
    for (;;) {
        long long owner = -1;
        if (lane == 0) {
            const unsigned long long h = ld_volatile_u64(sq.head), tl = ld_volatile_u64(sq.tail);
            if (h < tl) {
                if (atomicCAS(sq.head, h, h + 1ULL) == h) {
                    const unsigned long long* slot = sq.ring + (h & (sq.cap - 1));
                    unsigned long long v;
                    while (((v = ld_volatile_u64(slot)) >> 32) != h + 1ULL) spin_pause(20);
                    owner = (long long)(v & 0xffffffffULL);
                }
            } else if (ld_volatile_u64(sq.doneSearches) >= (unsigned long long)n) owner = -2;
        }
        owner = __shfl_sync(FULL, owner, 0);
        if (owner == -2) break;
        if (owner < 0) {
            spin_pause(idleNs);
            if (idleNs < 8000) idleNs <<= 1;
            continue;
        }
        idleNs = 500;
        if (st) {
            const long long now = clock64();
            if (lane == 0) st[29] += (unsigned long long)(now - tk);  // idle
            tk = now;
        }
        ScanJob* J = sq.jobs + owner;
        __threadfence();
        if (lane == 0) {  // the request, read past this SM's L1
            ScanJob& L = W.job;
            L.R = ld_cg(&J->R); L.pruned = ld_cg(&J->pruned); L.sibling = ld_cg(&J->sibling); L.failed0 = ld_cg(&J->failed0);
            L.best = ld_cg(&J->best); L.lastLK0 = ld_cg(&J->lastLK0); L.removedBLen = ld_cg(&J->removedBLen);
            L.isRemovedTip = ld_cg(&J->isRemovedTip); L.pathCap = ld_cg(&J->pathCap); L.qCap = ld_cg(&J->qCap);
            L.remK = ld_cg(&J->remK); L.remP = ld_cg(&J->remP); L.gpath = ld_cg(&J->gpath); L.qTop = ld_cg(&J->qTop);
            L.scoreRow = ld_cg(&J->scoreRow);
        }
        __syncwarp();
        warp_scan_job2<true>(m, t, sp, W, poolBytes, scanFlags, mbarParity, st, true);
        if (lane == 0) {
            const ScanJob& L = W.job;
            const bool declined = L.err == 4;
            J->bestOut = L.bestOut; J->phase1 = L.phase1; J->qN = L.qN; J->newBest = L.newBest; J->err = declined ? 0 : L.err;
            __threadfence();
            *reinterpret_cast<volatile int*>(&J->state) = declined ? 3 : 2;
            if (st) st[27] += 1;
        }
        __syncwarp();
        if (st) {
            const long long now = clock64();
            if (lane == 0) st[28] += (unsigned long long)(now - tk);  // serving
            tk = now;
        }
    }
}

}  // namespace maple
