// Per-thread likelihood primitives over packed genome lists (sm_100a, fp64).
//
// Each function is one co-walk of two sorted run-length streams by ONE thread, in the same
// order of floating-point operations as the reference so that merged lists and branch lengths
// are bit-identical to it (compile with -fmad=false: no FMA contraction).  Reference lines are
// into MAPLEv0.7.5.4.py:
//   getPartialVec 4073-4141 | simplify 3697-3717 | shorten 3721-3745 | mergeVectors 4446-4859
//   appendProbNode 6505-6785 | estimateBranchLengthWithDerivative 5040-5358
//   areVectorsDifferent 5419-5472 | passGenomeListThroughBranch 3749-3877 | rootVector 4916-4996
//
// Because every packed entry stores its end position, the next segment boundary of a co-walk is
// always min(end1,end2); the reference's "pos+1 / min(entry1[1],entry2[1])" case analysis and
// its per-case index bumps collapse into "advance every cursor whose end equals the boundary".
#pragma once
#include <cfloat>
#include <cmath>
#include "glist.cuh"

namespace maple {

struct DevModel {
    int lRef, U, errSS, rateVar;
    double Q[16];
    double pi[4];
    double errorRate, totError;
    double thresholdProb, thresholdDiffForUpdate, thresholdFoldChangeUpdate, minBLenSensitivity;
    const double* siteRates;   // [lRef] (rateVar)
    const double* errorRates;  // [lRef] (U && errSS)
    const double* cumRate;     // [lRef+1]
    const double* cumErr;      // [lRef+1] (U && errSS)
    const int32_t* cumBases;   // cumulativeBases [(lRef+1)*4] (findProbRoot only; maple_ctx_set_root_tables)
    const double* piLogErrCum; // rootFreqsLogErrorCumulative [lRef+1] (findProbRoot under the error model)
};

constexpr double kMinCarryOver = DBL_MIN * 1e50;  // :3623

// mutMatrices[pos][i][j] = Q[i][j]*siteRates[pos] (:6367), formed on access instead of stored
struct SiteQ {
    const double* Q;
    double r;
    bool rv;
    __device__ __forceinline__ SiteQ(const DevModel& m, int pos) : Q(m.Q), r(1.0), rv(m.rateVar != 0) {
        if (rv) r = __ldg(m.siteRates + pos);
    }
    __device__ __forceinline__ double at(int i, int j) const { return rv ? Q[i * 4 + j] * r : Q[i * 4 + j]; }
};

__device__ __forceinline__ double site_eps(const DevModel& m, int pos) {
    return (m.U && m.errSS) ? __ldg(m.errorRates + pos) : m.errorRate;
}

__device__ __forceinline__ void uniform4(double* o) { o[0] = o[1] = o[2] = o[3] = 0.25; }

// python's sum() of a 4-list: CPython >= 3.12 compensates (Neumaier); see oracle/maple_oracle.c
__device__ __forceinline__ double py_sum4(const double* v) {
#ifdef MAPLE_NAIVE_SUM
    return ((v[0] + v[1]) + v[2]) + v[3];
#else
    double f = v[0], c = 0.0;
#pragma unroll
    for (int i = 1; i < 4; i++) {
        double x = v[i], t = f + x;
        if (fabs(f) >= fabs(x)) c += (f - t) + x;
        else c += (x - t) + f;
        f = t;
    }
    if (c != 0.0 && isfinite(c)) f += c;
    return f;
#endif
}

// (A/B: -DMAPLE_LEAN compiles the getPartialVec helpers and tree_list out of line -- one copy each instead of one per use)
#ifdef MAPLE_LEAN
#define MAPLE_HELPER_INLINE __noinline__
#else
#define MAPLE_HELPER_INLINE __forceinline__
#endif

// getPartialVec for an O vector (:4085-4109)
__device__ MAPLE_HELPER_INLINE void gv_vec(const SiteQ& q, double t, const double* v, bool up, double* o) {
    if (t == 0.0) {
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
        return;
    }
    double r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double tot = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) tot += (up ? q.at(j, i) : q.at(i, j)) * v[j];
        tot *= t;
        tot += v[i];
        r[i] = tot;
    }
    // the reference bails out at the first negative component; later components do not matter
    if (r[0] < 0 || r[1] < 0 || r[2] < 0 || r[3] < 0) { uniform4(o); return; }
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
}

// getPartialVec for a single nucleotide x (:4110-4141); `flag` already includes usingErrorRate
// (QT: SiteQ, or anything with the same at(i, j))
template <class QT>
__device__ MAPLE_HELPER_INLINE void gv_nuc(const QT& q, double eps, int x, double t, bool up, bool flag, double* o) {
    if (flag) {
        double nv[4];
        const double e3 = eps * 0.33333;
#pragma unroll
        for (int i = 0; i < 4; i++) nv[i] = (i == x) ? 1.0 - eps : e3;
        if (t == 0.0) { o[0] = nv[0]; o[1] = nv[1]; o[2] = nv[2]; o[3] = nv[3]; return; }
        double r[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double tot = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) tot += q.at(j, i) * nv[i];
            tot *= t;
            tot += nv[j];
            r[j] = tot;
        }
        if (r[0] < 0 || r[1] < 0 || r[2] < 0 || r[3] < 0) { uniform4(o); return; }
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
        return;
    }
    if (t == 0.0) {
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = (i == x) ? 1.0 : 0.0;
        return;
    }
    double d = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double val = (up ? q.at(x, i) : q.at(i, x)) * t;
        if (i == x) { val += 1.0; d = val; }
        o[i] = val;
    }
    if (d < 0) uniform4(o);
}

__device__ __forceinline__ double sel4(const double* v, int i) {
    // register-friendly dynamic index into a 4-vector
    double r = v[0];
    r = (i == 1) ? v[1] : r;
    r = (i == 2) ? v[2] : r;
    r = (i == 3) ? v[3] : r;
    return r;
}

// One informative site of appendProbNode (:6586-6761): multiplies F by the site's factor; false = the reference returns -inf.
template <class Cur>
__device__ __forceinline__ bool append_site(const DevModel& m, const Cur& e1, const Cur& e2, int pos, double bLen, bool isTipC, bool U,
                                            double& F) {
    // an informative site: one side at least is a single-site entry
    double contrib = bLen;  // :6586-6599
    if (e1.type < 5) {
        if (e1.nl == 1) contrib += e1.l0();
        else if (e1.nl == 2) contrib += e1.l1();
    } else if (e1.nl == 1) contrib += e1.l0();
    if (e2.nl == 1) contrib += e2.l0();
    const SiteQ q(m, pos);
    double t2[4], t3[4], a[4];
    if (e1.type == T_R) {
        if (e2.type == T_O) {  // :6611-6638
            const int i1 = e2.nuc;
            e2.vec(a);
            const double ai = sel4(a, i1);
            if (ai > 0.02) F *= ai;
            else {
                double tot;
                if (e1.nl == 2) {
                    const bool flag1 = U && e1.flag;
                    const double eps = site_eps(m, pos);
                    tot = 0.0;
                    gv_vec(q, contrib, a, false, t3);
                    gv_nuc(q, eps, i1, e1.l0(), false, flag1, t2);
#pragma unroll
                    for (int i = 0; i < 4; i++) tot += t3[i] * t2[i] * m.pi[i];
                    tot /= m.pi[i1];
                } else if (contrib != 0.0) {
                    gv_vec(q, contrib, a, false, t3);
                    tot = sel4(t3, i1);
                } else tot = ai;
                F *= tot;
            }
        } else {  // R / different nucleotide :6640-6663
            const bool flag2 = U && (isTipC || (e2.nl > 0 && e2.flag));
            if (e1.nl == 2) {
                const bool flag1 = U && e1.flag;
                const int i1 = e2.nuc, i2 = e2.type;
                const double eps = site_eps(m, pos);
                gv_nuc(q, eps, i2, contrib, false, flag2, t3);
                gv_nuc(q, eps, i1, e1.l0(), false, flag1, t2);
                double tot = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++) tot += t3[i] * t2[i] * m.pi[i];
                F *= tot / m.pi[i1];
            } else if (flag2) {
                const double eps = site_eps(m, pos);
                F *= fmin(0.25, q.at(e2.nuc, e2.type) * contrib) + eps * 0.33333;
            } else if (contrib != 0.0) {
                F *= fmin(0.25, q.at(e2.nuc, e2.type) * contrib);
            } else return false;
        }
    } else if (e1.type == T_O) {  // :6674-6703
        e1.vec(a);
        if (e2.type == T_O) {
            double b[4];
            e2.vec(b);
            double tot = 0.0;
            if (contrib != 0.0) {
                gv_vec(q, contrib, b, false, t3);
#pragma unroll
                for (int j = 0; j < 4; j++) tot += a[j] * t3[j];
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) tot += a[j] * b[j];
            }
            F *= tot;
        } else {
            const int i2 = (e2.type == T_R) ? e1.nuc : e2.type;
            const double ai = sel4(a, i2);
            if (ai > 0.02) F *= ai;
            else {
                const bool fl = U && (isTipC || (e2.nl > 0 && e2.flag));
                const double eps = fl ? site_eps(m, pos) : 0.0;
                gv_nuc(q, eps, i2, contrib, false, fl, t3);
                double tot = 0.0;
#pragma unroll
                for (int j = 0; j < 4; j++) tot += a[j] * t3[j];
                F *= tot;
            }
        }
    } else {  // e1 is a non-reference nucleotide, e2 differs :6713-6761
        const bool flag1 = U && e1.nl > 0 && e1.flag;
        const int i1 = e1.type;
        if (e2.type < 5) {
            const int i2 = (e2.type == T_R) ? e1.nuc : e2.type;
            const bool flag2 = U && (isTipC || (e2.nl > 0 && e2.flag));
            if (e1.nl == 2) {
                const double eps = site_eps(m, pos);
                gv_nuc(q, eps, i2, contrib, false, flag2, t3);
                gv_nuc(q, eps, i1, e1.l0(), false, flag1, t2);
                double tot = 0.0;
#pragma unroll
                for (int j = 0; j < 4; j++) tot += m.pi[j] * t3[j] * t2[j];
                F *= tot / m.pi[i1];
            } else if (flag1 || flag2) {
                const double eps = site_eps(m, pos);
                F *= (fmin(0.25, q.at(i1, i2) * contrib) + (double)(int(flag1) + int(flag2)) * 0.33333 * eps);
            } else if (contrib != 0.0) {
                F *= fmin(0.25, q.at(i1, i2) * contrib);
            } else return false;
        } else {  // nucleotide / O
            e2.vec(a);
            const double ai = sel4(a, i1);
            if (ai > 0.02) F *= ai;
            else if (e1.nl == 2) {
                const double eps = site_eps(m, pos);
                gv_nuc(q, eps, i1, e1.l0(), false, flag1, t2);
                gv_vec(q, contrib, a, false, t3);
                double tot = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++) tot += t2[i] * t3[i] * m.pi[i];
                F *= (tot / m.pi[i1]);
            } else if (contrib != 0.0) {
                gv_vec(q, contrib, a, false, t3);
                F *= sel4(t3, i1);
            } else F *= ai;
        }
    }
    return true;
}

__device__ __forceinline__ bool append_informative(int t1, int t2) {
    return t1 != T_N && t2 != T_N && !(t1 == T_R && t2 == T_R) && !(t1 < 4 && t1 == t2);
}

// ------------------------------------------------------------------------------------------------
// appendProbNode (:6505-6785)
template <bool LD>
__device__ double dev_append(const DevModel& m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC,
                             bool isTipC, double bLen) {
    const int lRef = m.lRef;
    const bool U = m.U != 0;
    Cursor<LD> e1, e2;
    e1.init(kP, pP);
    e2.init(kC, pC);
    int pos = 0;
    double F = 1.0;
    double Lk = bLen * (-(double)lRef);
    if (U && isTipC) Lk += m.totError;
    for (;;) {
        const int newPos = min(e1.end, e2.end);
        if (append_informative(e1.type, e2.type)) {
            if (!append_site(m, e1, e2, pos, bLen, isTipC, U, F)) return -INFINITY;
        }
        pos = newPos;
        if (pos == lRef) break;
        if (e1.end == pos) e1.next();
        if (e2.end == pos) e2.next();
        if (F <= kMinCarryOver) {  // :6772-6783
            if (F < DBL_MIN) return -INFINITY;
            Lk += log(F);
            F = 1.0;
        }
    }
    if (!(F > 0.0)) return -INFINITY;
    return Lk + log(F);
}

// The cheap outcomes of an informative site, straight from the raw keys (same arithmetic as append_site): two certain
// states without an across-the-root length (:6657-6663, :6729-6742), and an O entry whose probability for the other side's
// state is above the 0.02 shortcut (:6615, :6692, :6746).  Returns false when the site needs the general code.
__device__ __forceinline__ bool append_site_fast(const DevModel& m, uint32_t k1, const double* pay1, uint32_t k2, const double* pay2, int pos,
                                                 double bLen, double& F, bool& minusInf) {
    const int t1 = int(k1 & 7u), t2 = int(k2 & 7u), nl1 = int((k1 >> 3) & 3u), nl2 = int((k2 >> 3) & 3u);
    const int nuc1 = int((k1 >> 6) & 3u), nuc2 = int((k2 >> 6) & 3u);
    if (t1 < 5 && t2 < 5) {
        if (nl1 == 2 || m.U) return false;
        double contrib = bLen;
        if (nl1 == 1) contrib += pay1[0];
        if (nl2 == 1) contrib += pay2[0];
        if (contrib == 0.0) { minusInf = true; return true; }
        const int from = t1 == T_R ? nuc2 : t1, to = t2 == T_R ? nuc1 : t2;
        const SiteQ q(m, pos);
        F *= fmin(0.25, q.at(from, to) * contrib);
        return true;
    }
    if (t1 < 5 && t2 == T_O) {
        const double ai = pay2[nl2 + (t1 == T_R ? nuc2 : t1)];
        if (ai > 0.02) { F *= ai; return true; }
        return false;
    }
    if (t1 == T_O && t2 < 5) {
        const double ai = pay1[nl1 + (t2 == T_R ? nuc1 : t2)];
        if (ai > 0.02) { F *= ai; return true; }
        return false;
    }
    return false;
}

// One queued site of dev_append_q4: returns F times the site's factor, or -1 when the reference returns -inf here.
__device__ __noinline__ double append_site_general(const DevModel& m, uint32_t k1, const double* pay1, uint32_t k2, const double* pay2, int pos,
                                                   double bLen, bool isTipC, double F) {
    Cursor<false> e1, e2;
    e1.key = nullptr; e1.pay = pay1; e1.decode(k1);
    e2.key = nullptr; e2.pay = pay2; e2.decode(k2);
    if (!append_site(m, e1, e2, pos, bLen, isTipC, m.U != 0, F)) return -1.0;
    return F;
}
__device__ __forceinline__ double append_site_ref(const DevModel& m, uint32_t k1, const double* pay1, uint32_t k2, const double* pay2, int pos,
                                                  double bLen, bool isTipC, double F) {
    bool minusInf = false;
    if (append_site_fast(m, k1, pay1, k2, pay2, pos, bLen, F, minusInf)) return minusInf ? -1.0 : F;
    return append_site_general(m, k1, pay1, k2, pay2, pos, bLen, isTipC, F);
}

// bit (t1*8+t2) set <=> append_informative(t1, t2), for entry types 0..6
__host__ __device__ constexpr unsigned long long append_informative_mask() {
    unsigned long long mk = 0;
    for (int t1 = 0; t1 < 7; t1++)
        for (int t2 = 0; t2 < 7; t2++)
            if (t1 != T_N && t2 != T_N && !(t1 == T_R && t2 == T_R) && !(t1 < 4 && t1 == t2)) mk |= 1ull << (t1 * 8 + t2);
    return mk;
}

// appendProbNode again, same arithmetic in the same order, arranged for a warp whose lanes score different pairs at
// once: every lane first walks to its next informative site -- a loop over the raw keys with a handful of integer
// operations per segment, payload offsets advanced without touching the payload -- then the lanes evaluate one site
// each together, so the site code runs once per "k-th site of every lane" instead of once per segment of any lane.
template <bool LD>
__device__ double dev_append_sitewise(const DevModel& m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC,
                                      bool isTipC, double bLen) {
    constexpr unsigned long long INF = append_informative_mask();
    const int lRef = m.lRef;
    const uint32_t *q1 = kP, *q2 = kC;
    uint32_t k1 = LD ? __ldg(q1) : *q1, k2 = LD ? __ldg(q2) : *q2;
    const double *y1 = pP, *y2 = pC;  // payload of the current entries
    int pos = 0;
    double F = 1.0;
    double Lk = bLen * (-(double)lRef);
    if (m.U && isTipC) Lk += m.totError;
    for (;;) {
        bool site = false;
        for (;;) {  // segments that contribute nothing leave F untouched, so the carry-over test has nothing to do
            if ((INF >> ((k1 & 7u) * 8u + (k2 & 7u))) & 1ull) { site = true; break; }
            const int e1 = int(k1 >> 8), e2 = int(k2 >> 8);
            pos = min(e1, e2);
            if (pos == lRef) break;
            if (e1 == pos) { y1 += ((k1 >> 3) & 3u) + ((k1 & 7u) == 6u ? 4u : 0u); ++q1; k1 = LD ? __ldg(q1) : *q1; }
            if (e2 == pos) { y2 += ((k2 >> 3) & 3u) + ((k2 & 7u) == 6u ? 4u : 0u); ++q2; k2 = LD ? __ldg(q2) : *q2; }
        }
        if (!site) break;
        F = append_site_ref(m, k1, y1, k2, y2, pos, bLen, isTipC, F);
        if (F < 0.0) return -INFINITY;
        const int e1 = int(k1 >> 8), e2 = int(k2 >> 8);
        pos = min(e1, e2);
        if (pos == lRef) break;
        if (e1 == pos) { y1 += ((k1 >> 3) & 3u) + ((k1 & 7u) == 6u ? 4u : 0u); ++q1; k1 = LD ? __ldg(q1) : *q1; }
        if (e2 == pos) { y2 += ((k2 >> 3) & 3u) + ((k2 & 7u) == 6u ? 4u : 0u); ++q2; k2 = LD ? __ldg(q2) : *q2; }
        if (F <= kMinCarryOver) {  // :6772-6783
            if (F < DBL_MIN) return -INFINITY;
            Lk += log(F);
            F = 1.0;
        }
    }
    if (!(F > 0.0)) return -INFINITY;
    return Lk + log(F);
}

// appendProbNode for a warp whose lanes score different pairs at once (subtree scans): same arithmetic in the same order
// as dev_append, but the walk and the site arithmetic are separated.  Each lane walks its two key streams with a few
// integer operations per segment and parks up to four informative sites (raw keys, payload offsets, position) in
// registers; then the lanes evaluate their parked sites together.  The long site code therefore runs about
// (sites per pair / 4) times per batch with most lanes active, instead of once per segment with a few.
template <bool LD>
__device__ double dev_append_q4(const DevModel& m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC, bool isTipC,
                                double bLen) {
    constexpr unsigned long long INF = append_informative_mask();
    const unsigned act = __activemask();  // the lanes scoring in this batch: they walk and evaluate in lock step
    const int lRef = m.lRef;
    const uint32_t *q1 = kP, *q2 = kC;
    uint32_t k1 = LD ? __ldg(q1) : *q1, k2 = LD ? __ldg(q2) : *q2;
    int o1 = 0, o2 = 0;  // payload offsets of the current entries
    int pos = 0;
    double F = 1.0;
    double Lk = bLen * (-(double)lRef);
    if (m.U && isTipC) Lk += m.totError;
    bool done = false, dead = false;  // dead: the reference returned -inf at some site
    for (;;) {
        uint32_t a1 = 0, a2 = 0, b1 = 0, b2 = 0, c1 = 0, c2 = 0, d1 = 0, d2 = 0;
        int ao1 = 0, ao2 = 0, bo1 = 0, bo2 = 0, co1 = 0, co2 = 0, do1 = 0, do2 = 0, ap = 0, bp = 0, cp = 0, dp = 0;
        int ns = 0;
        while (!done) {
            const int t1 = int(k1 & 7u), t2 = int(k2 & 7u);
            if ((INF >> (t1 * 8 + t2)) & 1ull) {
                if (ns == 0) { a1 = k1; a2 = k2; ao1 = o1; ao2 = o2; ap = pos; }
                else if (ns == 1) { b1 = k1; b2 = k2; bo1 = o1; bo2 = o2; bp = pos; }
                else if (ns == 2) { c1 = k1; c2 = k2; co1 = o1; co2 = o2; cp = pos; }
                else { d1 = k1; d2 = k2; do1 = o1; do2 = o2; dp = pos; }
                ns++;
            }
            const int e1 = int(k1 >> 8), e2 = int(k2 >> 8);
            pos = min(e1, e2);
            if (pos == lRef) { done = true; break; }
            if (e1 == pos) { o1 += int((k1 >> 3) & 3u) + (t1 == T_O ? 4 : 0); ++q1; k1 = LD ? __ldg(q1) : *q1; }
            if (e2 == pos) { o2 += int((k2 >> 3) & 3u) + (t2 == T_O ? 4 : 0); ++q2; k2 = LD ? __ldg(q2) : *q2; }
            if (ns == 4) break;
        }
        __syncwarp(act);
        for (int q = 0; q < 4; q++) {
            const bool mine = q < ns && !dead;
            if (!__any_sync(act, mine)) break;
            if (mine) {
                const uint32_t s1 = q == 0 ? a1 : q == 1 ? b1 : q == 2 ? c1 : d1, s2 = q == 0 ? a2 : q == 1 ? b2 : q == 2 ? c2 : d2;
                const int so1 = q == 0 ? ao1 : q == 1 ? bo1 : q == 2 ? co1 : do1, so2 = q == 0 ? ao2 : q == 1 ? bo2 : q == 2 ? co2 : do2;
                const int sp = q == 0 ? ap : q == 1 ? bp : q == 2 ? cp : dp;
                F = append_site_ref(m, s1, pP + so1, s2, pC + so2, sp, bLen, isTipC, F);
                if (F < 0.0) dead = true;
                else if (min(int(s1 >> 8), int(s2 >> 8)) != lRef && F <= kMinCarryOver) {  // :6772-6783
                    if (F < DBL_MIN) dead = true;
                    else { Lk += log(F); F = 1.0; }
                }
            }
        }
        if (dead) done = true;
        if (!__any_sync(act, !done)) break;
    }
    if (dead || !(F > 0.0)) return -INFINITY;
    return Lk + log(F);
}

// simplify (:3697-3717)
__device__ __forceinline__ int simplify4(const double* v, int refA, double thr) {
    double maxP = 0.0;
    int maxI = 0, numA = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (v[i] > maxP) { maxP = v[i]; maxI = i; }
        if (v[i] > thr) numA++;
    }
    if (numA == 1) return maxI == refA ? T_R : maxI;
    return T_O;
}

// ------------------------------------------------------------------------------------------------
// mergeVectors (:4446-4859).  flags bit0 = isUpDown, bit1 = returnLK.
// returns 0 = list written, 1 = None, 2 = likelihood underflow (the reference raises).
template <bool LD>
__device__ int dev_merge(const DevModel& m, const uint32_t* k1, const double* p1, double bLen1, bool fromTip1, const uint32_t* k2,
                         const double* p2, double bLen2, bool fromTip2, int flags, int numMinor1, int numMinor2, Writer& o,
                         double* outLk) {
    const int lRef = m.lRef;
    const bool U = m.U != 0, isUpDown = (flags & 1) != 0, returnLK = (flags & 2) != 0;
    Cursor<LD> e1, e2;
    e1.init(k1, p1);
    e2.init(k2, p2);
    int pos = 0;
    double totalFactor = 1.0, cumulPartLk = 0.0, cumErrorRate = 0.0;
    double nv[4], nv2[4];
    const double *cr = m.cumRate, *ce = m.cumErr;
    if (returnLK) {  // :4487-4494
        cumulPartLk = (bLen1 + bLen2) * (-(double)lRef);
        if (U) {
            if (fromTip1 || numMinor1) cumulPartLk += m.totError * (1 + numMinor1);
            if (fromTip2 || numMinor2) cumulPartLk += m.totError * (1 + numMinor2);
        }
    }
    for (;;) {
        const int newPos = min(e1.end, e2.end);
        if (e1.type == T_N || e2.type == T_N) {
            if (e1.type == T_N && e2.type == T_N) {
                o.put0(T_N, 0, newPos);
            } else if (e1.type == T_N) {
                if (e2.type < 5) {  // copy entry2, adding bLen2 :4501-4548
                    const int t = e2.type, nuc = e2.nuc;
                    if (isUpDown) {
                        if (e2.nl > 0) o.put(t, 2, U ? (e2.nl == 1 ? e2.flag : (e2.l1() != 0.0)) : 0, nuc, newPos, e2.l0() + bLen2, 0.0, nullptr);
                        else if (bLen2 != 0.0 || (U && fromTip2)) o.put(t, 2, U && fromTip2, nuc, newPos, bLen2, 0.0, nullptr);
                        else o.put0(t, nuc, newPos);
                    } else {
                        if (e2.nl > 0) o.put(t, 1, U ? (e2.nl == 1 ? e2.flag : (e2.l1() != 0.0)) : 0, nuc, newPos, e2.l0() + bLen2, 0.0, nullptr);
                        else if (bLen2 != 0.0 || (U && fromTip2)) o.put(t, 1, U && fromTip2, nuc, newPos, bLen2, 0.0, nullptr);
                        else o.put0(t, nuc, newPos);
                    }
                } else {  // N / O :4550-4576
                    double a[4];
                    e2.vec(a);
                    if (isUpDown) {
                        const SiteQ q(m, pos);
                        double totB = bLen2;
                        if (e2.nl == 1) totB += e2.l0();
                        gv_vec(q, totB, a, false, nv);
#pragma unroll
                        for (int i = 0; i < 4; i++) nv[i] *= m.pi[i];
                        const double s = py_sum4(nv);
#pragma unroll
                        for (int i = 0; i < 4; i++) nv[i] /= s;
                        o.put(T_O, 0, 0, e2.nuc, newPos, 0.0, 0.0, nv);
                    } else {
                        if (e2.nl == 1) o.put(T_O, 1, 0, e2.nuc, newPos, e2.l0() + bLen2, 0.0, a);
                        else if (bLen2 != 0.0) o.put(T_O, 1, 0, e2.nuc, newPos, bLen2, 0.0, a);
                        else o.put(T_O, 0, 0, e2.nuc, newPos, 0.0, 0.0, a);
                    }
                }
            } else {  // entry2 is N, entry1 informative :4590-4668
                if (e1.type < 5) {
                    const int t = e1.type, nuc = e1.nuc;
                    if (isUpDown) {
                        if (e1.nl == 0) {
                            if (bLen1 != 0.0) o.put(t, 1, 0, nuc, newPos, bLen1, 0.0, nullptr);
                            else o.put0(t, nuc, newPos);
                        } else if (e1.nl == 1) o.put(t, 1, U ? e1.flag : 0, nuc, newPos, e1.l0() + bLen1, 0.0, nullptr);
                        else o.put(t, 2, U ? e1.flag : 0, nuc, newPos, e1.l0(), e1.l1() + bLen1, nullptr);
                    } else {
                        if (e1.nl > 0) o.put(t, 1, U ? (e1.nl == 1 ? e1.flag : (e1.l1() != 0.0)) : 0, nuc, newPos, e1.l0() + bLen1, 0.0, nullptr);
                        else if (bLen1 != 0.0 || (U && fromTip1)) o.put(t, 1, U && fromTip1, nuc, newPos, bLen1, 0.0, nullptr);
                        else o.put0(t, nuc, newPos);
                    }
                } else {  // O / N :4644-4668
                    double a[4];
                    e1.vec(a);
                    const double l0 = e1.l0();
                    if (isUpDown && ((e1.nl == 1 && l0 > 0) || bLen1 != 0.0)) {
                        const SiteQ q(m, pos);
                        double totB = bLen1;
                        if (e1.nl == 1) totB += l0;
                        gv_vec(q, totB, a, true, nv);
                        const double s = py_sum4(nv);
#pragma unroll
                        for (int i = 0; i < 4; i++) nv[i] /= s;
                        o.put(T_O, 0, 0, e1.nuc, newPos, 0.0, 0.0, nv);
                    } else {
                        if (e1.nl == 1) o.put(T_O, 1, 0, e1.nuc, newPos, l0 + bLen1, 0.0, a);
                        else if (bLen1 != 0.0) o.put(T_O, 1, 0, e1.nuc, newPos, bLen1, 0.0, a);
                        else o.put(T_O, 0, 0, e1.nuc, newPos, 0.0, 0.0, a);
                    }
                }
            }
            if (returnLK) {  // :4578-4587 / :4670-4679
                cumulPartLk += (bLen1 + bLen2) * (__ldg(cr + pos) - __ldg(cr + newPos));
                if (U) {
                    if (fromTip1 || fromTip2) {
                        if (m.errSS) cumErrorRate = __ldg(ce + newPos) - __ldg(ce + pos);
                        else cumErrorRate = m.errorRate * (newPos - pos);
                    }
                    if (fromTip1) cumulPartLk += cumErrorRate;
                    if (fromTip2) cumulPartLk += cumErrorRate;
                }
            }
        } else if (e1.type == T_R && e2.type == T_R && !returnLK) {
            o.put0(T_R, 0, newPos);  // the overwhelmingly common segment
        } else {  // both informative :4682-4826
            double totLen1 = bLen1;
            if (e1.type == T_O) {
                if (e1.nl == 1) totLen1 += e1.l0();
            } else if (e1.nl >= 1) {
                totLen1 += e1.l0();
                if (e1.nl == 2) totLen1 += e1.l1();
            }
            double totLen2 = bLen2;
            if (e2.nl >= 1) totLen2 += e2.l0();
            const bool flag1 = U && e1.type != T_O && ((e1.nl > 0 && e1.flag) || fromTip1);
            const bool flag2 = U && e2.type != T_O && ((e2.nl > 0 && e2.flag) || fromTip2);
            int refNuc = -1;
            if (returnLK) {
                if (e1.type == T_R && e2.type == T_R) {  // :4704-4714
                    if (totLen2 > bLen2 || totLen1 > bLen1) {
                        cumulPartLk += (totLen2 - bLen2 + totLen1 - bLen1) * (__ldg(cr + newPos) - __ldg(cr + pos));
                        if (U) {
                            if (((!fromTip1) && flag1) || ((!fromTip2) && flag2)) {
                                if (m.errSS) cumErrorRate = __ldg(ce + pos) - __ldg(ce + newPos);
                                else cumErrorRate = m.errorRate * (pos - newPos);
                                if ((!fromTip1) && flag1) cumulPartLk += cumErrorRate;
                                if ((!fromTip2) && flag2) cumulPartLk += cumErrorRate;
                            }
                        }
                    }
                } else {  // :4715-4730
                    refNuc = (e1.type != T_R) ? e1.nuc : e2.nuc;
                    const SiteQ q(m, pos);
                    cumulPartLk -= q.at(refNuc, refNuc) * (bLen2 + bLen1);
                    if (U && ((e1.type != e2.type) || e1.type == T_O) && (fromTip1 || fromTip2)) {
                        cumErrorRate = m.errSS ? __ldg(m.errorRates + pos) : m.errorRate;
                        if (fromTip1) cumulPartLk += cumErrorRate;
                        if (fromTip2) cumulPartLk += cumErrorRate;
                    }
                }
            }
            if (e2.type == e1.type && e2.type < 5) {  // identical :4732-4751
                if (e1.type == T_R) o.put0(T_R, 0, newPos);
                else {
                    o.put0(e1.type, e1.nuc, newPos);
                    if (returnLK) {
                        const SiteQ q(m, pos);
                        cumulPartLk += q.at(e1.type, e1.type) * (totLen1 + totLen2);
                        if (U) {
                            if (((!fromTip1) && flag1) || ((!fromTip2) && flag2)) {
                                cumErrorRate = m.errSS ? __ldg(m.errorRates + pos) : m.errorRate;
                                if ((!fromTip1) && flag1) cumulPartLk -= cumErrorRate;
                                if ((!fromTip2) && flag2) cumulPartLk -= cumErrorRate;
                            }
                        }
                    }
                }
            } else if (totLen1 == 0.0 && totLen2 == 0.0 && e1.type < 5 && e2.type < 5 && !flag1 && !flag2) {
                return 1;  // :4753-4758
            } else {  // :4759-4826
                const double eps = site_eps(m, pos);
                const SiteQ q(m, pos);
                int i1, i2;
                if (e1.type == T_R) { refNuc = e2.nuc; i1 = refNuc; }
                else { refNuc = e1.nuc; i1 = e1.type; }
                if (i1 <= 4) {
                    if (totLen1 != 0.0 || flag1) {
                        if (isUpDown && e1.nl == 2) {
                            gv_nuc(q, eps, i1, e1.l0(), false, flag1, nv);
#pragma unroll
                            for (int i = 0; i < 4; i++) nv[i] *= m.pi[i];
                            const double up2 = e1.l1() + bLen1;
                            if (up2 != 0.0) {
                                double tmp[4];
                                gv_vec(q, up2, nv, true, tmp);
                                nv[0] = tmp[0]; nv[1] = tmp[1]; nv[2] = tmp[2]; nv[3] = tmp[3];
                            }
                        } else gv_nuc(q, eps, i1, totLen1, isUpDown, flag1, nv);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++) nv[i] = (i == i1) ? 1.0 : 0.0;
                    }
                } else {
                    double a[4];
                    e1.vec(a);
                    gv_vec(q, totLen1, a, isUpDown, nv);
                }
                i2 = (e2.type == T_R) ? refNuc : e2.type;
                if (i2 == T_O) {
                    double b[4];
                    e2.vec(b);
                    gv_vec(q, totLen2, b, false, nv2);
                } else if (totLen2 != 0.0 || flag2) gv_nuc(q, eps, i2, totLen2, false, flag2, nv2);
                else {
#pragma unroll
                    for (int i = 0; i < 4; i++) nv2[i] = (i == i2) ? 1.0 : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) nv[j] *= nv2[j];
                const double totSum = py_sum4(nv);
                if (totSum == 0.0) return 1;
#pragma unroll
                for (int i = 0; i < 4; i++) nv[i] /= totSum;
                const int state = simplify4(nv, refNuc, m.thresholdProb);
                if (state == T_O) o.put(T_O, 0, 0, refNuc, newPos, 0.0, 0.0, nv);
                else if (state == T_R) o.put0(T_R, 0, newPos);
                else o.put0(state, refNuc, newPos);
                if (returnLK) totalFactor *= totSum;
            }
        }
        pos = newPos;
        if (returnLK && totalFactor <= kMinCarryOver) {  // :4830-4839
            if (totalFactor < DBL_MIN) return 2;
            cumulPartLk += log(totalFactor);
            totalFactor = 1.0;
        }
        if (pos == lRef) break;
        if (e1.end == pos) e1.next();
        if (e2.end == pos) e2.next();
    }
    if (returnLK && outLk) *outLk = cumulPartLk + log(totalFactor);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// shorten (:3721-3745), out of place.  The anchor of a run is its FIRST entry (the reference never
// refreshes entryOld after a pop) while the surviving entry is the LAST one.
template <bool LD>
__device__ void dev_shorten(const DevModel& m, const uint32_t* k, const double* p, Writer& o) {
    const int lRef = m.lRef;
    const double thr = m.thresholdProb;
    Cursor<LD> c;
    c.init(k, p);
    int a_type = c.type, a_nl = c.nl, a_flag = c.flag;
    double a_l0 = c.l0(), a_l1 = c.l1();
    // pending entry (copied out because the cursor moves on)
    int p_type = c.type, p_nl = c.nl, p_flag = c.flag, p_nuc = c.nuc, p_end = c.end;
    double p_l0 = a_l0, p_l1 = a_l1, p_vec[4] = {0, 0, 0, 0};
    if (c.type == T_O) c.vec(p_vec);
    while (p_end != lRef) {
        c.next();
        const double l0 = c.l0(), l1 = c.l1();
        bool mergeable = false;
        if (c.type == T_R && a_type == T_R && c.nl == a_nl) {
            if (c.nl == 0) mergeable = true;
            else if (fabs(l0 - a_l0) > thr) mergeable = false;
            else if (c.nl == 2 && fabs(l1 - a_l1) > thr) mergeable = false;
            else mergeable = (!m.U) || (c.flag == a_flag);
        }
        if (!mergeable) {
            o.put(p_type, p_nl, p_flag, p_nuc, p_end, p_l0, p_l1, p_vec);
            a_type = c.type; a_nl = c.nl; a_flag = c.flag; a_l0 = l0; a_l1 = l1;
        }
        p_type = c.type; p_nl = c.nl; p_flag = c.flag; p_nuc = c.nuc; p_end = c.end; p_l0 = l0; p_l1 = l1;
        if (c.type == T_O) c.vec(p_vec);
    }
    o.put(p_type, p_nl, p_flag, p_nuc, p_end, p_l0, p_l1, p_vec);
}

// ------------------------------------------------------------------------------------------------
// estimateBranchLengthWithDerivative (:5040-5358).  returns 0 = value in *out, 1 = python False.
// ais: per-thread scratch with room for one double per informative site.
template <bool LD>
__device__ int dev_blen(const DevModel& m, const uint32_t* kP, const double* pP, const uint32_t* kC, const double* pC, bool fromTipC,
                        double* ais, double* out) {
    const int lRef = m.lRef;
    const bool U = m.U != 0;
    const double* pi = m.pi;
    Cursor<LD> e1, e2;
    e1.init(kP, pP);
    e2.init(kC, pC);
    int pos = 0, nA = 0, nZeros = 0;
    double c1 = -(double)lRef;
    double minAis = 0.0, maxAis = 0.0;
    const double* cr = m.cumRate;
    for (;;) {
        const int end = min(e1.end, e2.end);
        if (e2.type == T_N || e1.type == T_N) {
            c1 += (__ldg(cr + pos) - __ldg(cr + end));
        } else if (e1.type == T_R && e2.type == T_R) {
        } else {
            const SiteQ q(m, pos);
            if (e1.type == T_R) c1 -= q.at(e2.nuc, e2.nuc);
            else c1 -= q.at(e1.nuc, e1.nuc);
            const bool flag1 = U && e1.type != T_O && e1.nl > 0 && e1.flag;
            const bool flag2 = U && e2.type != T_O && (fromTipC || (e2.nl > 0 && e2.flag));
            const double eps = site_eps(m, pos);
            double contrib = 0.0;
            if (e1.type < 5) {
                if (e1.nl == 1) contrib = e1.l0();
                else if (e1.nl == 2) contrib = e1.l1();
            } else if (e1.nl == 1) contrib = e1.l0();
            if (e2.nl >= 1) contrib += e2.l0();
            double coeff0 = 0.0, coeff1 = 0.0;
            int mode = 0;  // 1: (coeff0, coeff1) pair; 2: a single a-value
            bool valid = true;
            if (e1.type == T_R || (e1.type < 4 && e2.type != e1.type)) {
                const int x = (e1.type == T_R) ? e2.nuc : e1.type;  // parent state
                if (e2.type == T_O) {  // :5128-5155, :5253-5278
                    double v[4];
                    e2.vec(v);
                    mode = 1;
                    if (e1.nl == 2) {
                        const double l0 = e1.l0();
                        coeff0 = pi[x] * sel4(v, x);
                        coeff1 = 0.0;
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            coeff0 += pi[i] * q.at(i, x) * l0 * v[i];
                            coeff1 += q.at(x, i) * v[i];
                        }
                        coeff1 *= pi[x];
                        if (contrib != 0.0) coeff0 += coeff1 * contrib;
                        if (flag1) {
                            coeff0 -= 1.33333 * eps * pi[x] * sel4(v, x);
#pragma unroll
                            for (int i = 0; i < 4; i++) coeff0 += pi[i] * v[i] * 0.33333 * eps;
                        }
                    } else {
                        coeff0 = sel4(v, x);
                        coeff1 = 0.0;
#pragma unroll
                        for (int j = 0; j < 4; j++) coeff1 += q.at(x, j) * v[j];
                        if (contrib != 0.0) coeff0 += coeff1 * contrib;
                    }
                } else {  // child is a different single nucleotide (or R under a nucleotide parent)
                    const int c = (e2.type == T_R) ? e1.nuc : e2.type;
                    mode = 2;
                    if (e1.nl == 2) {  // :5158-5172, :5230-5242
                        coeff0 = pi[c] * q.at(c, x) * e1.l0();
                        if (contrib != 0.0) coeff0 += pi[x] * q.at(x, c) * contrib;
                        if (flag2) coeff0 += pi[x] * 0.33333 * eps;
                        if (flag1) coeff0 += pi[c] * 0.33333 * eps;
                        coeff1 = pi[x] * q.at(x, c);
                        if (coeff1 != 0.0) coeff0 = coeff0 / coeff1;
                        else valid = false;
                    } else {
                        coeff0 = contrib;
                        if (flag2) {
                            const double qxc = q.at(x, c);
                            if (e1.type == T_R && qxc == 0.0) valid = false;  // :5176 guards, :5246 does not
                            else coeff0 += eps * 0.33333 / qxc;
                        }
                    }
                }
            } else if (e1.type == T_O) {  // :5188-5215
                double a[4];
                e1.vec(a);
                mode = 1;
                if (e2.type == T_O) {
                    double b[4];
                    e2.vec(b);
                    coeff0 = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
                    coeff1 = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) coeff1 += a[i] * b[j] * q.at(i, j);
                    if (contrib != 0.0) coeff0 += coeff1 * contrib;
                } else {
                    const int i2 = (e2.type == T_R) ? e1.nuc : e2.type;
                    coeff0 = sel4(a, i2);
                    coeff1 = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; i++) coeff1 += a[i] * q.at(i, i2);
                    if (contrib != 0.0) coeff0 += coeff1 * contrib;
                    if (flag2) coeff0 += eps * 0.33333;
                }
            } else {  // same non-reference nucleotide on both sides :5220-5221
                c1 += q.at(e1.type, e1.type);
            }
            double aval = 0.0;
            bool push = false;
            if (mode == 1) {
                if (coeff1 < 0.0) c1 += coeff1 / coeff0;
                else if (coeff1 != 0.0) { aval = coeff0 / coeff1; push = true; }
            } else if (mode == 2 && valid) {
                if (coeff0 != 0.0) { aval = coeff0; push = true; }
                else nZeros++;
            }
            if (push) {
                if (nA == 0) { minAis = aval; maxAis = aval; }
                else { minAis = fmin(minAis, aval); maxAis = fmax(maxAis, aval); }
                ais[nA++] = aval;
            }
        }
        pos = end;
        if (pos == lRef) break;
        if (e1.end == pos) e1.next();
        if (e2.end == pos) e2.next();
    }
    // :5298-5358
    c1 = -c1;
    const int n = nA + nZeros;
    *out = 0.0;
    if (n == 0) return 1;
    if (nZeros) minAis = fmin(0.0, minAis);
    if (minAis < 0.0) { *out = 0.1; return 0; }
    const double sens = m.minBLenSensitivity;
    double tDown = fmin(0.1, n / c1 - minAis);
    if (tDown <= 0.0) return 1;
    double vDown = nZeros ? nZeros / tDown : 0.0;
    for (int i = 0; i < nA; i++) vDown += 1.0 / (ais[i] + tDown);
    double tUp = fmin(0.1, n / c1 - maxAis);
    if (tUp >= 0.1) { *out = 0.1; return 0; }
    if (tUp <= sens) tUp = (minAis != 0.0) ? 0.0 : sens;
    double vUp = nZeros ? nZeros / tUp : 0.0;
    for (int i = 0; i < nA; i++) vUp += 1.0 / (ais[i] + tUp);
    if (vDown > c1 + sens || vUp < c1 - sens) {
        if (vUp < c1 - sens && tUp == 0.0) return 1;
        if (vDown > c1 + sens && tDown >= 0.1) { *out = 0.1; return 0; }
    }
    while (tDown - tUp > sens) {
        const double tMid = (tUp + tDown) / 2;
        double vMid = nZeros ? nZeros / tMid : 0.0;
        for (int i = 0; i < nA; i++) vMid += 1.0 / (ais[i] + tMid);
        if (vMid > c1) tUp = tMid;
        else tDown = tMid;
    }
    *out = tUp;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// areVectorsDifferent (:5419-5472)
template <bool LD>
__device__ bool dev_differ(const DevModel& m, const uint32_t* k1, const double* p1, const uint32_t* k2, const double* p2) {
    if (!k2) return true;
    const int lRef = m.lRef;
    const bool U = m.U != 0;
    const double thr = m.thresholdProb;
    Cursor<LD> e1, e2;
    e1.init(k1, p1);
    e2.init(k2, p2);
    for (;;) {
        if (e1.type != e2.type) return true;
        if (e1.nl != e2.nl) return true;
        if (e1.type < 5) {
            if (e1.nl >= 1) {
                if (fabs(e1.l0() - e2.l0()) > thr) return true;
                if (e1.nl == 2 && fabs(e1.l1() - e2.l1()) > thr) return true;
                if (U && e1.flag != e2.flag) return true;
            }
        } else if (e1.type == T_O) {
            if (e1.nl == 1 && fabs(e1.l0() - e2.l0()) > thr) return true;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double a = e1.v(i), b = e2.v(i);
                const double d = fabs(a - b);
                if (d != 0.0) {
                    if (a == 0.0 || b == 0.0) return true;
                    if (d > m.thresholdDiffForUpdate ||
                        (d > thr && ((d / a > m.thresholdFoldChangeUpdate) || (d / b > m.thresholdFoldChangeUpdate))))
                        return true;
                }
            }
        }
        const int pos = min(e1.end, e2.end);
        if (pos == lRef) break;
        if (e1.end == pos) e1.next();
        if (e2.end == pos) e2.next();
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// rootVector (:4916-4996) for a list already expressed relative to the reference genome (the
// MAT re-referencing around it, :4928-4940 and :4990-4993, is done by the caller); not shortened.
template <bool LD>
__device__ void dev_root_vector(const DevModel& m, const uint32_t* k, const double* p, double bLen, bool isFromTip, Writer& o) {
    const int lRef = m.lRef;
    const bool U = m.U != 0;
    Cursor<LD> c;
    c.init(k, p);
    int pos = 0;
    for (;;) {
        if (c.type == T_N) o.put0(T_N, 0, c.end);
        else if (c.type == T_O) {
            double a[4], nv[4];
            c.vec(a);
            double totB = bLen;
            if (c.nl == 1) totB += c.l0();
            if (totB != 0.0) {
                const SiteQ q(m, pos);
                gv_vec(q, totB, a, false, nv);
#pragma unroll
                for (int i = 0; i < 4; i++) nv[i] *= m.pi[i];
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) nv[i] = a[i] * m.pi[i];
            }
            const double s = py_sum4(nv);
#pragma unroll
            for (int i = 0; i < 4; i++) nv[i] /= s;
            o.put(T_O, 0, 0, c.nuc, c.end, 0.0, 0.0, nv);
        } else if (U) {
            const bool flag1 = (c.nl > 0 && c.flag) || isFromTip;
            if (c.nl >= 1) o.put(c.type, 2, flag1, c.nuc, c.end, c.l0() + bLen, 0.0, nullptr);
            else if (bLen != 0.0 || flag1) o.put(c.type, 2, flag1, c.nuc, c.end, bLen, 0.0, nullptr);
            else o.put0(c.type, c.nuc, c.end);
        } else {
            if (c.nl == 1) o.put(c.type, 2, 0, c.nuc, c.end, c.l0() + bLen, 0.0, nullptr);
            else if (bLen != 0.0) o.put(c.type, 2, 0, c.nuc, c.end, bLen, 0.0, nullptr);
            else o.put0(c.type, c.nuc, c.end);
        }
        pos = c.end;
        if (pos == lRef) break;
        c.next();
    }
}

// ------------------------------------------------------------------------------------------------
// findProbRoot (:4865-4912) for a list expressed relative to the reference genome
template <bool LD>
__device__ double dev_prob_root(const DevModel& m, const uint32_t* k, const double* p) {
    const int lRef = m.lRef;
    const bool U = m.U != 0;
    Cursor<LD> c;
    c.init(k, p);
    double logLK = 0.0, logFactor = 1.0;
    int pos = 0;
    double piLog[4];
#pragma unroll
    for (int i = 0; i < 4; i++) piLog[i] = log(m.pi[i]);
    for (;;) {
        if (U && c.type < 5 && c.nl > 0 && c.flag) {
            if (c.type == T_R) logLK += __ldg(m.piLogErrCum + c.end) - __ldg(m.piLogErrCum + pos);
            else {
                const double eps = site_eps(m, pos);
                logFactor *= (sel4(m.pi, c.type) * (1.0 - 1.33333 * eps) + 0.33333 * eps);
            }
        } else if (c.type == T_R) {
#pragma unroll
            for (int i = 0; i < 4; i++) logLK += piLog[i] * (double)(__ldg(m.cumBases + c.end * 4 + i) - __ldg(m.cumBases + pos * 4 + i));
        } else if (c.type < 4) logLK += sel4(piLog, c.type);
        else if (c.type == T_O) {
            double a[4];
            c.vec(a);
            double tot = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) tot += m.pi[i] * a[i];
            logFactor *= tot;
        }
        pos = c.end;
        if (logFactor <= kMinCarryOver) {
            if (logFactor < DBL_MIN) return -INFINITY;
            logLK += log(logFactor);
            logFactor = 1.0;
        }
        if (pos == lRef) break;
        c.next();
    }
    logLK += log(logFactor);
    return logLK;
}

}  // namespace maple
