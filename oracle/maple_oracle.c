/*
 * maple_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the SPR-likelihood hot path of the reference MAPLEv0.7.5.4.py,
 * operating on the packed genome-list format of maple_b200/genome_list.py.  It exists only
 * so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * can check and time the CUDA path against an independent CPU implementation.  The product
 * path (maple_b200/) never links, imports or calls anything in this directory.
 *
 * Parity pin: every function here is checked against (inputs -> output) vectors recorded from
 * the unmodified reference running in the build container (tests/golden/NAME.json.gz, generated
 * by tests/golden/make_golden.py; checked by tests/test_oracle_golden.py).
 *
 * Reference lines (all in /root/reference/MAPLEv0.7.5.4.py):
 *   getPartialVec 4073-4141 | simplify 3697-3717 | shorten 3721-3745
 *   mergeVectors 4446-4859 | appendProbNode 6505-6785
 *   estimateBranchLengthWithDerivative 5040-5358 | areVectorsDifferent 5419-5472
 *   passGenomeListThroughBranch 3749-3877 | rootVector 4916-4996 | findProbRoot 4865-4912
 *
 * Floating point: compiled with -ffp-contract=off so that a*b+c rounds twice like CPython.
 * All four co-walks use the fact that the packed format stores an explicit end position for
 * every entry: the next segment boundary is always min(end1,end2) (the reference's
 * "pos+1 or min(entry1[1],entry2[1])" case split collapses to that).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t lRef;
    int32_t U;        /* usingErrorRate */
    int32_t errSS;    /* errorRateSiteSpecific */
    int32_t rateVar;  /* useRateVariation */
    double Q[16];     /* mutMatrixGlobal, row-major Q[i][j] */
    double pi[4];     /* rootFreqs */
    double errorRate; /* errorRateGlobal */
    double totError;
    double thresholdProb;
    double thresholdDiffForUpdate;
    double thresholdFoldChangeUpdate;
    double minBLenSensitivity;
    const double *siteRates;  /* [lRef] or NULL */
    const double *errorRates; /* [lRef] or NULL */
    const double *cumRate;    /* cumulativeRate [lRef+1] */
    const double *cumErr;     /* cumulativeErrorRate [lRef+1] or NULL */
    const int32_t *cumBases;  /* cumulativeBases [(lRef+1)*4] or NULL (findProbRoot only) */
    const double *piLogErrCum; /* rootFreqsLogErrorCumulative [lRef+1] or NULL */
} OrModel;

#define MINIMUM_CARRY_OVER (DBL_MIN * 1e50) /* reference :3623 */

typedef struct {
    const uint32_t *key;
    const double *pay;
    int type, nl, flag, nuc, end;
    double l0, l1;
    const double *vec;
} Cur;

static inline void cur_next(Cur *c) {
    uint32_t k = *c->key++;
    c->type = (int)(k & 7u);
    c->nl = (int)((k >> 3) & 3u);
    c->flag = (int)((k >> 5) & 1u);
    c->nuc = (int)((k >> 6) & 3u);
    c->end = (int)(k >> 8);
    c->l0 = c->l1 = 0.0;
    if (c->nl >= 1) c->l0 = *c->pay++;
    if (c->nl == 2) c->l1 = *c->pay++;
    c->vec = NULL;
    if (c->type == 6) {
        c->vec = c->pay;
        c->pay += 4;
    }
}

static inline void cur_init(Cur *c, const uint32_t *key, const double *pay) {
    c->key = key;
    c->pay = pay;
    cur_next(c);
}

static inline const double *site_Q(const OrModel *m, int pos, double *buf) {
    if (!m->rateVar) return m->Q;
    double r = m->siteRates[pos];
    for (int i = 0; i < 16; i++) buf[i] = m->Q[i] * r; /* mutMatrices[pos][j][k]=Q[j][k]*siteRates[pos], :6367 */
    return buf;
}

static inline double site_eps(const OrModel *m, int pos) {
    return (m->U && m->errSS) ? m->errorRates[pos] : m->errorRate;
}

static inline void uniform4(double *o) { o[0] = o[1] = o[2] = o[3] = 0.25; }

/* python's builtin sum() over a 4-list of floats.  CPython >= 3.12 (the interpreter the golden
 * vectors were recorded with) uses Neumaier compensated summation for floats; pypy3 and older
 * CPython add left to right.  -DMAPLE_NAIVE_SUM selects the latter.  The two differ by at most
 * one ulp of the normalised vectors. */
static inline double py_sum4(const double *v) {
#ifdef MAPLE_NAIVE_SUM
    return ((v[0] + v[1]) + v[2]) + v[3];
#else
    double f = v[0], c = 0.0;
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

/* getPartialVec, i12==6 (:4085-4109) */
static void gv_vec(const double *Q, double t, const double *v, int up, double *o) {
    if (t == 0.0) {
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
        return;
    }
    double r[4];
    for (int i = 0; i < 4; i++) {
        double tot = 0.0;
        for (int j = 0; j < 4; j++) tot += (up ? Q[j * 4 + i] : Q[i * 4 + j]) * v[j];
        tot *= t;
        tot += v[i];
        if (tot < 0) { uniform4(o); return; }
        r[i] = tot;
    }
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
}

/* getPartialVec, i12<4 (:4110-4141); flag must already include "usingErrorRate and" */
static void gv_nuc(const double *Q, double eps, int x, double t, int up, int flag, double *o) {
    if (flag) {
        double nv[4];
        for (int i = 0; i < 4; i++) nv[i] = eps * 0.33333;
        nv[x] = 1.0 - eps;
        if (t == 0.0) { o[0] = nv[0]; o[1] = nv[1]; o[2] = nv[2]; o[3] = nv[3]; return; }
        double r[4];
        for (int j = 0; j < 4; j++) {
            double tot = 0.0;
            for (int i = 0; i < 4; i++) tot += Q[j * 4 + i] * nv[i];
            tot *= t;
            tot += nv[j];
            if (tot < 0) { uniform4(o); return; }
            r[j] = tot;
        }
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
        return;
    }
    if (t == 0.0) {
        o[0] = o[1] = o[2] = o[3] = 0.0;
        o[x] += 1.0;
        return;
    }
    for (int i = 0; i < 4; i++) o[i] = (up ? Q[x * 4 + i] : Q[i * 4 + x]) * t;
    o[x] += 1.0;
    if (o[x] < 0) uniform4(o);
}

/* ------------------------------------------------------------------ appendProbNode */
double or_append(const OrModel *m, const uint32_t *kP, const double *pP, const uint32_t *kC, const double *pC,
                 int isTipC, double bLen) {
    const int lRef = m->lRef, U = m->U;
    Cur e1, e2;
    cur_init(&e1, kP, pP);
    cur_init(&e2, kC, pC);
    int pos = 0;
    double F = 1.0, contrib = bLen;
    double Lk = bLen * (-(double)lRef);
    if (U && isTipC) Lk += m->totError;
    double Qb[16], t2[4], t3[4];
    for (;;) {
        int newPos = e1.end < e2.end ? e1.end : e2.end;
        if (e2.type != 5 && e1.type != 5) {
            if (e1.type != e2.type || e1.type == 6) { /* :6586-6599 */
                contrib = bLen;
                if (e1.type < 5) {
                    if (e1.nl == 1) contrib += e1.l0;
                    else if (e1.nl == 2) contrib += e1.l1;
                } else if (e1.nl == 1) contrib += e1.l0;
                if (e2.nl == 1) contrib += e2.l0;
            }
            if (e1.type == 4) {
                if (e2.type == 4) {
                    /* R / R: nothing */
                } else if (e2.type == 6) { /* :6611-6638 */
                    const double *Q = site_Q(m, pos, Qb);
                    int i1 = e2.nuc;
                    if (e2.vec[i1] > 0.02) F *= e2.vec[i1];
                    else {
                        double tot;
                        if (e1.nl == 2) {
                            int flag1 = U && e1.flag;
                            double eps = site_eps(m, pos);
                            tot = 0.0;
                            gv_vec(Q, contrib, e2.vec, 0, t3);
                            gv_nuc(Q, eps, i1, e1.l0, 0, flag1, t2);
                            for (int i = 0; i < 4; i++) tot += t3[i] * t2[i] * m->pi[i];
                            tot /= m->pi[i1];
                        } else if (contrib != 0.0) {
                            gv_vec(Q, contrib, e2.vec, 0, t3);
                            tot = t3[i1];
                        } else tot = e2.vec[i1];
                        F *= tot;
                    }
                } else { /* R / different nucleotide :6640-6663 */
                    int flag2 = U && (isTipC || (e2.nl > 0 && e2.flag));
                    const double *Q = site_Q(m, pos, Qb);
                    if (e1.nl == 2) {
                        int flag1 = U && e1.flag;
                        int i1 = e2.nuc, i2 = e2.type;
                        double eps = site_eps(m, pos);
                        gv_nuc(Q, eps, i2, contrib, 0, flag2, t3);
                        gv_nuc(Q, eps, i1, e1.l0, 0, flag1, t2);
                        double tot = 0.0;
                        for (int i = 0; i < 4; i++) tot += t3[i] * t2[i] * m->pi[i];
                        F *= tot / m->pi[i1];
                    } else if (flag2) {
                        double eps = site_eps(m, pos);
                        F *= fmin(0.25, Q[e2.nuc * 4 + e2.type] * contrib) + eps * 0.33333;
                    } else if (contrib != 0.0) {
                        F *= fmin(0.25, Q[e2.nuc * 4 + e2.type] * contrib);
                    } else return -INFINITY;
                }
            } else if (e1.type == 6) { /* :6674-6703 */
                const double *Q = site_Q(m, pos, Qb);
                if (e2.type == 6) {
                    double tot = 0.0;
                    if (contrib != 0.0) {
                        gv_vec(Q, contrib, e2.vec, 0, t3);
                        for (int j = 0; j < 4; j++) tot += e1.vec[j] * t3[j];
                    } else
                        for (int j = 0; j < 4; j++) tot += e1.vec[j] * e2.vec[j];
                    F *= tot;
                } else {
                    int i2 = (e2.type == 4) ? e1.nuc : e2.type;
                    if (e1.vec[i2] > 0.02) F *= e1.vec[i2];
                    else {
                        int fl = U && (isTipC || (e2.nl > 0 && e2.flag));
                        double eps = fl ? site_eps(m, pos) : 0.0;
                        gv_nuc(Q, eps, i2, contrib, 0, fl, t3);
                        double tot = 0.0;
                        for (int j = 0; j < 4; j++) tot += e1.vec[j] * t3[j];
                        F *= tot;
                    }
                }
            } else { /* e1 is a non-reference nucleotide :6713-6761 */
                if (e2.type != e1.type) {
                    int flag1 = U && e1.nl > 0 && e1.flag;
                    const double *Q = site_Q(m, pos, Qb);
                    int i1 = e1.type;
                    if (e2.type < 5) {
                        int i2 = (e2.type == 4) ? e1.nuc : e2.type;
                        int flag2 = U && (isTipC || (e2.nl > 0 && e2.flag));
                        if (e1.nl == 2) {
                            double eps = site_eps(m, pos);
                            gv_nuc(Q, eps, i2, contrib, 0, flag2, t3);
                            gv_nuc(Q, eps, i1, e1.l0, 0, flag1, t2);
                            double tot = 0.0;
                            for (int j = 0; j < 4; j++) tot += m->pi[j] * t3[j] * t2[j];
                            F *= tot / m->pi[i1];
                        } else if (flag1 || flag2) {
                            double eps = site_eps(m, pos);
                            F *= (fmin(0.25, Q[i1 * 4 + i2] * contrib) + (double)(flag1 + flag2) * 0.33333 * eps);
                        } else if (contrib != 0.0) {
                            F *= fmin(0.25, Q[i1 * 4 + i2] * contrib);
                        } else return -INFINITY;
                    } else { /* nucleotide / O */
                        double eps = site_eps(m, pos);
                        if (e2.vec[i1] > 0.02) F *= e2.vec[i1];
                        else if (e1.nl == 2) {
                            gv_nuc(Q, eps, i1, e1.l0, 0, flag1, t2);
                            gv_vec(Q, contrib, e2.vec, 0, t3);
                            double tot = 0.0;
                            for (int i = 0; i < 4; i++) tot += t2[i] * t3[i] * m->pi[i];
                            F *= (tot / m->pi[i1]);
                        } else if (contrib != 0.0) {
                            gv_vec(Q, contrib, e2.vec, 0, t3);
                            F *= t3[i1];
                        } else F *= e2.vec[i1];
                    }
                }
            }
        }
        pos = newPos;
        if (pos == lRef) break;
        if (e1.end == pos) cur_next(&e1);
        if (e2.end == pos) cur_next(&e2);
        if (F <= MINIMUM_CARRY_OVER) { /* :6772-6783 */
            if (F < DBL_MIN) return -INFINITY;
            Lk += log(F);
            F = 1.0;
        }
    }
    if (!(F > 0.0)) return -INFINITY; /* python would raise on log(0); the caller treats it as no placement */
    return Lk + log(F);
}

/* ------------------------------------------------------------------ output writer */
typedef struct {
    uint32_t *key;
    double *pay;
    int nk, np;
} Out;

static inline void out_put(Out *o, int type, int nl, int flag, int nuc, int end, double l0, double l1, const double *vec) {
    o->key[o->nk++] = (uint32_t)(type & 7) | ((uint32_t)(nl & 3) << 3) | ((uint32_t)(flag ? 1 : 0) << 5) |
                      ((uint32_t)(nuc & 3) << 6) | ((uint32_t)end << 8);
    if (nl >= 1) o->pay[o->np++] = l0;
    if (nl == 2) o->pay[o->np++] = l1;
    if (type == 6) {
        o->pay[o->np++] = vec[0]; o->pay[o->np++] = vec[1]; o->pay[o->np++] = vec[2]; o->pay[o->np++] = vec[3];
    }
}

/* simplify (:3697-3717): 4 = collapses to the local reference, 0-3 = to that nucleotide, 6 = stays O */
static int simplify4(const double *v, int refA, double thr) {
    double maxP = 0.0;
    int maxI = 0, numA = 0;
    for (int i = 0; i < 4; i++) {
        if (v[i] > maxP) { maxP = v[i]; maxI = i; }
        if (v[i] > thr) numA++;
    }
    if (numA == 1) return maxI == refA ? 4 : maxI;
    return 6;
}

/* ------------------------------------------------------------------ mergeVectors
 * returns 0 = list written, 1 = None (incompatible at zero distance / zero product).
 * flags: bit0 isUpDown, bit1 returnLK.  Output capacity: nk <= n1+n2, np <= 6*(n1+n2). */
int or_merge(const OrModel *m, const uint32_t *k1, const double *p1, double bLen1, int fromTip1, const uint32_t *k2,
             const double *p2, double bLen2, int fromTip2, int flags, int numMinor1, int numMinor2, uint32_t *outKey,
             double *outPay, int32_t *outNk, int32_t *outNp, double *outLk) {
    const int lRef = m->lRef, U = m->U, isUpDown = flags & 1, returnLK = (flags >> 1) & 1;
    Cur e1, e2;
    cur_init(&e1, k1, p1);
    cur_init(&e2, k2, p2);
    Out o = {outKey, outPay, 0, 0};
    int pos = 0;
    double totalFactor = 1.0, cumulPartLk = 0.0, cumErrorRate = 0.0;
    double Qb[16], nv[4], nv2[4];
    const double *cr = m->cumRate, *ce = m->cumErr;
    if (returnLK) { /* :4487-4494 */
        cumulPartLk = (bLen1 + bLen2) * (-(double)lRef);
        if (U) {
            if (fromTip1 || numMinor1) cumulPartLk += m->totError * (1 + numMinor1);
            if (fromTip2 || numMinor2) cumulPartLk += m->totError * (1 + numMinor2);
        }
    }
    for (;;) {
        int newPos = e1.end < e2.end ? e1.end : e2.end;
        if (e1.type == 5 || e2.type == 5) {
            if (e1.type == 5 && e2.type == 5) {
                out_put(&o, 5, 0, 0, 0, newPos, 0, 0, NULL);
            } else if (e1.type == 5) {
                if (e2.type < 5) { /* copy entry2, adding bLen2 :4501-4548 */
                    int t = e2.type, nuc = e2.nuc;
                    if (isUpDown) {
                        if (U) {
                            if (e2.nl == 0) {
                                if (bLen2 != 0.0 || fromTip2) out_put(&o, t, 2, fromTip2, nuc, newPos, bLen2, 0.0, NULL);
                                else out_put(&o, t, 0, 0, nuc, newPos, 0, 0, NULL);
                            } else out_put(&o, t, 2, e2.nl == 1 ? e2.flag : (e2.l1 != 0.0), nuc, newPos, e2.l0 + bLen2, 0.0, NULL);
                        } else {
                            if (e2.nl > 0) out_put(&o, t, 2, 0, nuc, newPos, e2.l0 + bLen2, 0.0, NULL);
                            else if (bLen2 != 0.0) out_put(&o, t, 2, 0, nuc, newPos, bLen2, 0.0, NULL);
                            else out_put(&o, t, 0, 0, nuc, newPos, 0, 0, NULL);
                        }
                    } else {
                        if (U) {
                            if (e2.nl == 0) {
                                if (bLen2 != 0.0 || fromTip2) out_put(&o, t, 1, fromTip2, nuc, newPos, bLen2, 0, NULL);
                                else out_put(&o, t, 0, 0, nuc, newPos, 0, 0, NULL);
                            } else out_put(&o, t, 1, e2.nl == 1 ? e2.flag : (e2.l1 != 0.0), nuc, newPos, e2.l0 + bLen2, 0, NULL);
                        } else {
                            if (e2.nl > 0) out_put(&o, t, 1, 0, nuc, newPos, e2.l0 + bLen2, 0, NULL);
                            else if (bLen2 != 0.0) out_put(&o, t, 1, 0, nuc, newPos, bLen2, 0, NULL);
                            else out_put(&o, t, 0, 0, nuc, newPos, 0, 0, NULL);
                        }
                    }
                } else { /* N / O :4550-4576 */
                    if (isUpDown) {
                        const double *Q = site_Q(m, pos, Qb);
                        double totB = bLen2;
                        if (e2.nl == 1) totB += e2.l0;
                        gv_vec(Q, totB, e2.vec, 0, nv);
                        for (int i = 0; i < 4; i++) nv[i] *= m->pi[i];
                        double s = py_sum4(nv);
                        for (int i = 0; i < 4; i++) nv[i] /= s;
                        out_put(&o, 6, 0, 0, e2.nuc, newPos, 0, 0, nv);
                    } else {
                        if (e2.nl == 1) out_put(&o, 6, 1, 0, e2.nuc, newPos, e2.l0 + bLen2, 0, e2.vec);
                        else if (bLen2 != 0.0) out_put(&o, 6, 1, 0, e2.nuc, newPos, bLen2, 0, e2.vec);
                        else out_put(&o, 6, 0, 0, e2.nuc, newPos, 0, 0, e2.vec);
                    }
                }
            } else { /* entry2 is N, entry1 informative :4590-4668 */
                if (e1.type < 5) {
                    int t = e1.type, nuc = e1.nuc;
                    if (isUpDown) {
                        if (U) {
                            if (e1.nl == 0) {
                                if (bLen1 != 0.0) out_put(&o, t, 1, 0, nuc, newPos, bLen1, 0, NULL);
                                else out_put(&o, t, 0, 0, nuc, newPos, 0, 0, NULL);
                            } else if (e1.nl == 1) out_put(&o, t, 1, e1.flag, nuc, newPos, e1.l0 + bLen1, 0, NULL);
                            else out_put(&o, t, 2, e1.flag, nuc, newPos, e1.l0, e1.l1 + bLen1, NULL);
                        } else {
                            if (e1.nl == 0) {
                                if (bLen1 != 0.0) out_put(&o, t, 1, 0, nuc, newPos, bLen1, 0, NULL);
                                else out_put(&o, t, 0, 0, nuc, newPos, 0, 0, NULL);
                            } else if (e1.nl == 1) out_put(&o, t, 1, 0, nuc, newPos, e1.l0 + bLen1, 0, NULL);
                            else out_put(&o, t, 2, 0, nuc, newPos, e1.l0, e1.l1 + bLen1, NULL);
                        }
                    } else {
                        if (U) {
                            if (e1.nl == 0) {
                                if (bLen1 != 0.0 || fromTip1) out_put(&o, t, 1, fromTip1, nuc, newPos, bLen1, 0, NULL);
                                else out_put(&o, t, 0, 0, nuc, newPos, 0, 0, NULL);
                            } else out_put(&o, t, 1, e1.nl == 1 ? e1.flag : (e1.l1 != 0.0), nuc, newPos, e1.l0 + bLen1, 0, NULL);
                        } else {
                            if (e1.nl > 0) out_put(&o, t, 1, 0, nuc, newPos, e1.l0 + bLen1, 0, NULL);
                            else if (bLen1 != 0.0) out_put(&o, t, 1, 0, nuc, newPos, bLen1, 0, NULL);
                            else out_put(&o, t, 0, 0, nuc, newPos, 0, 0, NULL);
                        }
                    }
                } else { /* O / N :4644-4668 */
                    if (isUpDown && ((e1.nl == 1 && e1.l0 > 0) || bLen1 != 0.0)) {
                        const double *Q = site_Q(m, pos, Qb);
                        double totB = bLen1;
                        if (e1.nl == 1) totB += e1.l0;
                        gv_vec(Q, totB, e1.vec, 1, nv);
                        double s = py_sum4(nv);
                        for (int i = 0; i < 4; i++) nv[i] /= s;
                        out_put(&o, 6, 0, 0, e1.nuc, newPos, 0, 0, nv);
                    } else {
                        if (e1.nl == 1) out_put(&o, 6, 1, 0, e1.nuc, newPos, e1.l0 + bLen1, 0, e1.vec);
                        else if (bLen1 != 0.0) out_put(&o, 6, 1, 0, e1.nuc, newPos, bLen1, 0, e1.vec);
                        else out_put(&o, 6, 0, 0, e1.nuc, newPos, 0, 0, e1.vec);
                    }
                }
            }
            if (returnLK) { /* :4578-4587 / :4670-4679 */
                cumulPartLk += (bLen1 + bLen2) * (cr[pos] - cr[newPos]);
                if (U) {
                    if (fromTip1 || fromTip2) {
                        if (m->errSS) cumErrorRate = ce[newPos] - ce[pos];
                        else cumErrorRate = m->errorRate * (newPos - pos);
                    }
                    if (fromTip1) cumulPartLk += cumErrorRate;
                    if (fromTip2) cumulPartLk += cumErrorRate;
                }
            }
        } else { /* both informative :4682-4826 */
            double totLen1 = bLen1;
            if (e1.type == 6) {
                if (e1.nl == 1) totLen1 += e1.l0;
            } else if (e1.nl >= 1) {
                totLen1 += e1.l0;
                if (e1.nl == 2) totLen1 += e1.l1;
            }
            double totLen2 = bLen2;
            if (e2.nl >= 1) totLen2 += e2.l0;
            int flag1 = U && e1.type != 6 && ((e1.nl > 0 && e1.flag) || fromTip1);
            int flag2 = U && e2.type != 6 && ((e2.nl > 0 && e2.flag) || fromTip2);
            int refNuc = -1;
            if (returnLK) {
                if (e1.type == 4 && e2.type == 4) { /* :4704-4714 */
                    if (totLen2 > bLen2 || totLen1 > bLen1) {
                        cumulPartLk += (totLen2 - bLen2 + totLen1 - bLen1) * (cr[newPos] - cr[pos]);
                        if (U) {
                            if (((!fromTip1) && flag1) || ((!fromTip2) && flag2)) {
                                if (m->errSS) cumErrorRate = ce[pos] - ce[newPos];
                                else cumErrorRate = m->errorRate * (pos - newPos);
                                if ((!fromTip1) && flag1) cumulPartLk += cumErrorRate;
                                if ((!fromTip2) && flag2) cumulPartLk += cumErrorRate;
                            }
                        }
                    }
                } else { /* :4715-4730 */
                    refNuc = (e1.type != 4) ? e1.nuc : e2.nuc;
                    const double *Q = site_Q(m, pos, Qb);
                    cumulPartLk -= Q[refNuc * 4 + refNuc] * (bLen2 + bLen1);
                    if (U && ((e1.type != e2.type) || e1.type == 6) && (fromTip1 || fromTip2)) {
                        cumErrorRate = m->errSS ? m->errorRates[pos] : m->errorRate;
                        if (fromTip1) cumulPartLk += cumErrorRate;
                        if (fromTip2) cumulPartLk += cumErrorRate;
                    }
                }
            }
            if (e2.type == e1.type && e2.type < 5) { /* identical :4732-4751 */
                if (e1.type == 4) out_put(&o, 4, 0, 0, 0, newPos, 0, 0, NULL);
                else {
                    out_put(&o, e1.type, 0, 0, e1.nuc, newPos, 0, 0, NULL);
                    if (returnLK) {
                        const double *Q = site_Q(m, pos, Qb);
                        cumulPartLk += Q[e1.type * 4 + e1.type] * (totLen1 + totLen2);
                        if (U) {
                            if (((!fromTip1) && flag1) || ((!fromTip2) && flag2)) {
                                cumErrorRate = m->errSS ? m->errorRates[pos] : m->errorRate;
                                if ((!fromTip1) && flag1) cumulPartLk -= cumErrorRate;
                                if ((!fromTip2) && flag2) cumulPartLk -= cumErrorRate;
                            }
                        }
                    }
                }
            } else if (totLen1 == 0.0 && totLen2 == 0.0 && e1.type < 5 && e2.type < 5 && !flag1 && !flag2) {
                return 1; /* :4753-4758 */
            } else { /* :4759-4826 */
                double eps = site_eps(m, pos);
                const double *Q = site_Q(m, pos, Qb);
                int i1, i2;
                if (e1.type == 4) { refNuc = e2.nuc; i1 = refNuc; }
                else { refNuc = e1.nuc; i1 = e1.type; }
                if (i1 <= 4) {
                    if (totLen1 != 0.0 || flag1) {
                        if (isUpDown && e1.nl == 2) {
                            gv_nuc(Q, eps, i1, e1.l0, 0, flag1, nv);
                            for (int i = 0; i < 4; i++) nv[i] *= m->pi[i];
                            if (e1.l1 + bLen1 != 0.0) {
                                double tmp[4];
                                gv_vec(Q, e1.l1 + bLen1, nv, 1, tmp);
                                memcpy(nv, tmp, sizeof tmp);
                            }
                        } else gv_nuc(Q, eps, i1, totLen1, isUpDown, flag1, nv);
                    } else {
                        nv[0] = nv[1] = nv[2] = nv[3] = 0.0;
                        nv[i1] = 1.0;
                    }
                } else gv_vec(Q, totLen1, e1.vec, isUpDown, nv);
                i2 = (e2.type == 4) ? refNuc : e2.type;
                if (i2 == 6) gv_vec(Q, totLen2, e2.vec, 0, nv2);
                else if (totLen2 != 0.0 || flag2) gv_nuc(Q, eps, i2, totLen2, 0, flag2, nv2);
                else {
                    nv2[0] = nv2[1] = nv2[2] = nv2[3] = 0.0;
                    nv2[i2] = 1.0;
                }
                for (int j = 0; j < 4; j++) nv[j] *= nv2[j];
                double totSum = py_sum4(nv);
                if (totSum == 0.0) return 1;
                for (int i = 0; i < 4; i++) nv[i] /= totSum;
                int state = simplify4(nv, refNuc, m->thresholdProb);
                if (state == 6) out_put(&o, 6, 0, 0, refNuc, newPos, 0, 0, nv);
                else if (state == 4) out_put(&o, 4, 0, 0, 0, newPos, 0, 0, NULL);
                else out_put(&o, state, 0, 0, refNuc, newPos, 0, 0, NULL);
                if (returnLK) totalFactor *= totSum;
            }
        }
        pos = newPos;
        if (returnLK && totalFactor <= MINIMUM_CARRY_OVER) { /* :4830-4839 */
            if (totalFactor < DBL_MIN) return 2;
            cumulPartLk += log(totalFactor);
            totalFactor = 1.0;
        }
        if (pos == lRef) break;
        if (e1.end == pos) cur_next(&e1);
        if (e2.end == pos) cur_next(&e2);
    }
    *outNk = o.nk;
    *outNp = o.np;
    if (returnLK && outLk) *outLk = cumulPartLk + log(totalFactor);
    return 0;
}

/* ------------------------------------------------------------------ shorten (:3721-3745), out of place */
void or_shorten(const OrModel *m, const uint32_t *k, const double *p, uint32_t *outKey, double *outPay, int32_t *outNk,
                int32_t *outNp) {
    const int lRef = m->lRef;
    const double thr = m->thresholdProb;
    Cur c;
    cur_init(&c, k, p);
    Out o = {outKey, outPay, 0, 0};
    /* The reference keeps comparing against `entryOld`, which is NOT refreshed after a pop: the
     * anchor of a run is its first entry, while the entry that survives is the last one. */
    Cur anchor = c, pending = c;
    while (pending.end != lRef) {
        cur_next(&c);
        int mergeable = 0;
        if (c.type == 4 && anchor.type == 4 && c.nl == anchor.nl) {
            if (c.nl == 0) mergeable = 1;
            else if (fabs(c.l0 - anchor.l0) > thr) mergeable = 0;
            else if (c.nl == 2 && fabs(c.l1 - anchor.l1) > thr) mergeable = 0;
            else mergeable = (!m->U) || (c.flag == anchor.flag);
        }
        if (!mergeable) {
            out_put(&o, pending.type, pending.nl, pending.flag, pending.nuc, pending.end, pending.l0, pending.l1, pending.vec);
            anchor = c;
        }
        pending = c;
    }
    out_put(&o, pending.type, pending.nl, pending.flag, pending.nuc, pending.end, pending.l0, pending.l1, pending.vec);
    *outNk = o.nk;
    *outNp = o.np;
}

/* ------------------------------------------------------------------ estimateBranchLengthWithDerivative
 * returns 0 = value in *out, 1 = python False.  scratch: ais[n1+n2] */
int or_blen(const OrModel *m, const uint32_t *kP, const double *pP, const uint32_t *kC, const double *pC, int fromTipC,
            double *ais, double *out) {
    const int lRef = m->lRef, U = m->U;
    const double *pi = m->pi;
    Cur e1, e2;
    cur_init(&e1, kP, pP);
    cur_init(&e2, kC, pC);
    int pos = 0, nA = 0, nZeros = 0;
    double c1 = -(double)lRef;
    double Qb[16];
    const double *cr = m->cumRate;
    for (;;) {
        int end = e1.end < e2.end ? e1.end : e2.end;
        if (e2.type == 5 || e1.type == 5) {
            c1 += (cr[pos] - cr[end]);
        } else if (e1.type == 4 && e2.type == 4) {
        } else {
            const double *Q = site_Q(m, pos, Qb);
            if (e1.type == 4) c1 -= Q[e2.nuc * 4 + e2.nuc];
            else c1 -= Q[e1.nuc * 4 + e1.nuc];
            int flag1 = U && e1.type != 6 && e1.nl > 0 && e1.flag;
            int flag2 = U && e2.type != 6 && (fromTipC || (e2.nl > 0 && e2.flag));
            double eps = site_eps(m, pos);
            double contrib = 0.0;
            if (e1.type < 5) {
                if (e1.nl == 1) contrib = e1.l0;
                else if (e1.nl == 2) contrib = e1.l1;
            } else if (e1.nl == 1) contrib = e1.l0;
            if (e2.nl >= 1) contrib += e2.l0;
            double coeff0 = 0.0, coeff1 = 0.0;
            int mode = 0; /* 1: coeff0/coeff1 pair; 2: single a-value; 0: nothing */
            int valid = 1;
            if (e1.type == 4 || (e1.type < 4 && e2.type != e1.type)) {
                /* parent state x: for R it is the local reference carried by the child entry */
                int x = (e1.type == 4) ? e2.nuc : e1.type;
                if (e2.type == 6) { /* :5128-5155, :5253-5278 */
                    const double *v = e2.vec;
                    mode = 1;
                    if (e1.nl == 2) {
                        coeff0 = pi[x] * v[x];
                        coeff1 = 0.0;
                        for (int i = 0; i < 4; i++) {
                            coeff0 += pi[i] * Q[i * 4 + x] * e1.l0 * v[i];
                            coeff1 += Q[x * 4 + i] * v[i];
                        }
                        coeff1 *= pi[x];
                        if (contrib != 0.0) coeff0 += coeff1 * contrib;
                        if (flag1) {
                            coeff0 -= 1.33333 * eps * pi[x] * v[x];
                            for (int i = 0; i < 4; i++) coeff0 += pi[i] * v[i] * 0.33333 * eps;
                        }
                    } else {
                        coeff0 = v[x];
                        coeff1 = 0.0;
                        for (int j = 0; j < 4; j++) coeff1 += Q[x * 4 + j] * v[j];
                        if (contrib != 0.0) coeff0 += coeff1 * contrib;
                    }
                } else { /* child is a different single nucleotide (or R under a nucleotide parent) */
                    int c = (e2.type == 4) ? e1.nuc : e2.type;
                    mode = 2;
                    if (e1.nl == 2) { /* :5158-5172, :5230-5242 */
                        coeff0 = pi[c] * Q[c * 4 + x] * e1.l0;
                        if (contrib != 0.0) coeff0 += pi[x] * Q[x * 4 + c] * contrib;
                        if (flag2) coeff0 += pi[x] * 0.33333 * eps;
                        if (flag1) coeff0 += pi[c] * 0.33333 * eps;
                        coeff1 = pi[x] * Q[x * 4 + c];
                        if (coeff1 != 0.0) coeff0 = coeff0 / coeff1;
                        else valid = 0;
                    } else {
                        coeff0 = contrib;
                        if (flag2) {
                            /* the R-parent branch guards against a zero rate (:5176), the nucleotide-parent one does not (:5246) */
                            if (e1.type == 4 && Q[x * 4 + c] == 0.0) valid = 0;
                            else coeff0 += eps * 0.33333 / Q[x * 4 + c];
                        }
                    }
                }
            } else if (e1.type == 6) { /* :5188-5215 */
                const double *a = e1.vec;
                mode = 1;
                if (e2.type == 6) {
                    const double *b = e2.vec;
                    coeff0 = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
                    coeff1 = 0.0;
                    for (int i = 0; i < 4; i++)
                        for (int j = 0; j < 4; j++) coeff1 += a[i] * b[j] * Q[i * 4 + j];
                    if (contrib != 0.0) coeff0 += coeff1 * contrib;
                } else {
                    int i2 = (e2.type == 4) ? e1.nuc : e2.type;
                    coeff0 = a[i2];
                    coeff1 = 0.0;
                    for (int i = 0; i < 4; i++) coeff1 += a[i] * Q[i * 4 + i2];
                    if (contrib != 0.0) coeff0 += coeff1 * contrib;
                    if (flag2) coeff0 += eps * 0.33333;
                }
            } else { /* same non-reference nucleotide on both sides :5220-5221 */
                c1 += Q[e1.type * 4 + e1.type];
            }
            if (mode == 1) {
                if (coeff1 < 0.0) c1 += coeff1 / coeff0;
                else if (coeff1 != 0.0) ais[nA++] = coeff0 / coeff1;
            } else if (mode == 2 && valid) {
                if (coeff0 != 0.0) ais[nA++] = coeff0;
                else nZeros++;
            }
        }
        pos = end;
        if (pos == lRef) break;
        if (e1.end == pos) cur_next(&e1);
        if (e2.end == pos) cur_next(&e2);
    }
    /* :5298-5358 */
    c1 = -c1;
    int n = nA + nZeros;
    *out = 0.0;
    if (n == 0) return 1;
    double minAis = 0.0, maxAis = 0.0;
    if (nA) {
        minAis = maxAis = ais[0];
        for (int i = 1; i < nA; i++) {
            if (ais[i] < minAis) minAis = ais[i];
            if (ais[i] > maxAis) maxAis = ais[i];
        }
    }
    if (nZeros) minAis = fmin(0.0, minAis);
    if (minAis < 0.0) { *out = 0.1; return 0; }
    const double sens = m->minBLenSensitivity;
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
        double tMid = (tUp + tDown) / 2;
        double vMid = nZeros ? nZeros / tMid : 0.0;
        for (int i = 0; i < nA; i++) vMid += 1.0 / (ais[i] + tMid);
        if (vMid > c1) tUp = tMid;
        else tDown = tMid;
    }
    *out = tUp;
    return 0;
}

/* ------------------------------------------------------------------ areVectorsDifferent (:5419-5472) */
int or_differ(const OrModel *m, const uint32_t *k1, const double *p1, const uint32_t *k2, const double *p2) {
    if (!k2) return 1;
    const int lRef = m->lRef, U = m->U;
    const double thr = m->thresholdProb;
    Cur e1, e2;
    cur_init(&e1, k1, p1);
    cur_init(&e2, k2, p2);
    for (;;) {
        if (e1.type != e2.type) return 1;
        if (e1.nl != e2.nl) return 1; /* tuple lengths are a function of (type, nLens) for a fixed U */
        if (e1.type < 5) {
            if (e1.nl >= 1) {
                if (fabs(e1.l0 - e2.l0) > thr) return 1;
                if (e1.nl == 2 && fabs(e1.l1 - e2.l1) > thr) return 1;
                if (U && e1.flag != e2.flag) return 1; /* abs(True-False)=1 > thresholdProb */
            }
        } else if (e1.type == 6) {
            if (e1.nl == 1 && fabs(e1.l0 - e2.l0) > thr) return 1;
            for (int i = 0; i < 4; i++) {
                double a = e1.vec[i], b = e2.vec[i];
                double d = fabs(a - b);
                if (d != 0.0) {
                    if (a == 0.0 || b == 0.0) return 1;
                    if (d > m->thresholdDiffForUpdate ||
                        (d > thr && ((d / a > m->thresholdFoldChangeUpdate) || (d / b > m->thresholdFoldChangeUpdate))))
                        return 1;
                }
            }
        }
        int pos = e1.end < e2.end ? e1.end : e2.end;
        if (pos == lRef) break;
        if (e1.end == pos) cur_next(&e1);
        if (e2.end == pos) cur_next(&e2);
    }
    return 0;
}

/* ------------------------------------------------------------------ passGenomeListThroughBranch (:3749-3877)
 * mut: nMut triples (pos1based, upNuc, downNuc) sorted by position.  Output capacity nk <= n + 2*nMut. */
void or_pass_branch(const OrModel *m, const uint32_t *k, const double *p, const int32_t *mut, int nMut, int dirIsUp,
                    uint32_t *outKey, double *outPay, int32_t *outNk, int32_t *outNp) {
    const int lRef = m->lRef;
    Cur c;
    cur_init(&c, k, p);
    Out o = {outKey, outPay, 0, 0};
    int iM = 0, lastPos = 0;
    for (;;) {
        if (c.type == 5) {
            out_put(&o, 5, 0, 0, 0, c.end, 0, 0, NULL);
            lastPos = c.end;
            while (iM < nMut && mut[3 * iM] <= lastPos) iM++;
        } else if (c.type < 4) {
            lastPos += 1;
            if (iM < nMut && mut[3 * iM] <= lastPos) {
                int target = dirIsUp ? mut[3 * iM + 1] : mut[3 * iM + 2];
                iM++;
                if (c.type == target) out_put(&o, 4, c.nl, c.flag, 0, lastPos, c.l0, c.l1, NULL);
                else out_put(&o, c.type, c.nl, c.flag, target, lastPos, c.l0, c.l1, NULL);
            } else out_put(&o, c.type, c.nl, c.flag, c.nuc, lastPos, c.l0, c.l1, NULL);
        } else if (c.type == 4) {
            while (iM < nMut && mut[3 * iM] <= c.end) {
                if (mut[3 * iM] > lastPos + 1) {
                    lastPos = mut[3 * iM] - 1;
                    out_put(&o, 4, c.nl, c.flag, 0, lastPos, c.l0, c.l1, NULL);
                }
                lastPos += 1;
                int nucToPass = dirIsUp ? mut[3 * iM + 2] : mut[3 * iM + 1];
                int newEl = dirIsUp ? mut[3 * iM + 1] : mut[3 * iM + 2];
                iM++;
                out_put(&o, nucToPass, c.nl, c.flag, newEl, lastPos, c.l0, c.l1, NULL);
            }
            if (lastPos < c.end) {
                lastPos = c.end;
                out_put(&o, 4, c.nl, c.flag, 0, lastPos, c.l0, c.l1, NULL);
            }
        } else {
            lastPos += 1;
            if (iM < nMut && mut[3 * iM] <= lastPos) {
                int newEl = dirIsUp ? mut[3 * iM + 1] : mut[3 * iM + 2];
                iM++;
                out_put(&o, 6, c.nl, 0, newEl, lastPos, c.l0, 0, c.vec);
            } else out_put(&o, 6, c.nl, 0, c.nuc, lastPos, c.l0, 0, c.vec);
        }
        if (lastPos == lRef) break;
        cur_next(&c);
    }
    *outNk = o.nk;
    *outNp = o.np;
}

/* ------------------------------------------------------------------ rootVector (:4916-4996) for a list that is already
 * expressed relative to the reference genome (no MAT mutations between node and root); NOT shortened here. */
void or_root_vector(const OrModel *m, const uint32_t *k, const double *p, double bLen, int isFromTip, uint32_t *outKey,
                    double *outPay, int32_t *outNk, int32_t *outNp) {
    const int lRef = m->lRef, U = m->U;
    Cur c;
    cur_init(&c, k, p);
    Out o = {outKey, outPay, 0, 0};
    int pos = 0;
    double Qb[16], nv[4];
    for (;;) {
        if (c.type == 5) out_put(&o, 5, 0, 0, 0, c.end, 0, 0, NULL);
        else if (c.type == 6) {
            double totB = bLen;
            if (c.nl == 1) totB += c.l0;
            if (totB != 0.0) {
                const double *Q = site_Q(m, pos, Qb);
                gv_vec(Q, totB, c.vec, 0, nv);
                for (int i = 0; i < 4; i++) nv[i] *= m->pi[i];
            } else
                for (int i = 0; i < 4; i++) nv[i] = c.vec[i] * m->pi[i];
            double s = py_sum4(nv);
            for (int i = 0; i < 4; i++) nv[i] /= s;
            out_put(&o, 6, 0, 0, c.nuc, c.end, 0, 0, nv);
        } else if (U) {
            int flag1 = (c.nl > 0 && c.flag) || isFromTip;
            if (c.nl >= 1) out_put(&o, c.type, 2, flag1, c.nuc, c.end, c.l0 + bLen, 0.0, NULL);
            else if (bLen != 0.0 || flag1) out_put(&o, c.type, 2, flag1, c.nuc, c.end, bLen, 0.0, NULL);
            else out_put(&o, c.type, 0, 0, c.nuc, c.end, 0, 0, NULL);
        } else {
            if (c.nl == 1) out_put(&o, c.type, 2, 0, c.nuc, c.end, c.l0 + bLen, 0.0, NULL);
            else if (bLen != 0.0) out_put(&o, c.type, 2, 0, c.nuc, c.end, bLen, 0.0, NULL);
            else out_put(&o, c.type, 0, 0, c.nuc, c.end, 0, 0, NULL);
        }
        pos = c.end;
        if (pos == lRef) break;
        cur_next(&c);
    }
    *outNk = o.nk;
    *outNp = o.np;
}

/* ------------------------------------------------------------------ findProbRoot (:4865-4912), list relative to the reference */
double or_prob_root(const OrModel *m, const uint32_t *k, const double *p) {
    const int lRef = m->lRef, U = m->U;
    Cur c;
    cur_init(&c, k, p);
    double logLK = 0.0, logFactor = 1.0;
    int pos = 0;
    double piLog[4];
    for (int i = 0; i < 4; i++) piLog[i] = log(m->pi[i]);
    for (;;) {
        if (U && c.type < 5 && c.nl > 0 && c.flag) {
            if (c.type == 4) logLK += m->piLogErrCum[c.end] - m->piLogErrCum[pos];
            else {
                double eps = m->errSS ? m->errorRates[pos] : m->errorRate;
                logFactor *= (m->pi[c.type] * (1.0 - 1.33333 * eps) + 0.33333 * eps);
            }
        } else if (c.type == 4) {
            for (int i = 0; i < 4; i++)
                logLK += piLog[i] * (double)(m->cumBases[c.end * 4 + i] - m->cumBases[pos * 4 + i]);
        } else if (c.type < 4) logLK += piLog[c.type];
        else if (c.type == 6) {
            double tot = 0.0;
            for (int i = 0; i < 4; i++) tot += m->pi[i] * c.vec[i];
            logFactor *= tot;
        }
        pos = c.end;
        if (logFactor <= MINIMUM_CARRY_OVER) {
            if (logFactor < DBL_MIN) return -INFINITY;
            logLK += log(logFactor);
            logFactor = 1.0;
        }
        if (pos == lRef) break;
        cur_next(&c);
    }
    logLK += log(logFactor);
    return logLK;
}

/* ------------------------------------------------------------------ batch drivers (OpenMP over host cores) */
void or_append_batch(const OrModel *m, const uint32_t *key, const double *pay, const int64_t *keyStart,
                     const int64_t *payStart, int64_t n, const int32_t *pIdx, const int32_t *cIdx, const uint8_t *isTip,
                     const double *bLen, double *out) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        int p = pIdx[i], c = cIdx[i];
        out[i] = or_append(m, key + keyStart[p], pay + payStart[p], key + keyStart[c], pay + payStart[c], isTip[i], bLen[i]);
    }
}

void or_merge_batch(const OrModel *m, const uint32_t *key, const double *pay, const int64_t *keyStart,
                    const int64_t *payStart, int64_t n, const int32_t *idx1, const double *bLen1, const uint8_t *tip1,
                    const int32_t *idx2, const double *bLen2, const uint8_t *tip2, const uint8_t *flags,
                    const int32_t *numMinor1, const int32_t *numMinor2, uint32_t *outKey, double *outPay,
                    const int64_t *outKeyStart, const int64_t *outPayStart, int32_t *outNk, int32_t *outNp, double *outLk,
                    int32_t *outStatus) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        int a = idx1[i], b = idx2[i];
        double lk = 0.0;
        outNk[i] = outNp[i] = 0;
        outStatus[i] = or_merge(m, key + keyStart[a], pay + payStart[a], bLen1[i], tip1[i], key + keyStart[b],
                                pay + payStart[b], bLen2[i], tip2[i], flags[i], numMinor1 ? numMinor1[i] : 0,
                                numMinor2 ? numMinor2[i] : 0, outKey + outKeyStart[i], outPay + outPayStart[i],
                                &outNk[i], &outNp[i], &lk);
        if (outLk) outLk[i] = lk;
    }
}

void or_blen_batch(const OrModel *m, const uint32_t *key, const double *pay, const int64_t *keyStart,
                   const int64_t *payStart, const int32_t *nkeys, int64_t n, const int32_t *pIdx, const int32_t *cIdx,
                   const uint8_t *fromTip, double *out, int32_t *outStatus) {
#pragma omp parallel
    {
        double *ais = NULL;
        int cap = 0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; i++) {
            int p = pIdx[i], c = cIdx[i];
            int need = nkeys[p] + nkeys[c] + 1;
            if (need > cap) {
                free(ais);
                cap = need * 2;
                ais = (double *)malloc(sizeof(double) * cap);
            }
            outStatus[i] = or_blen(m, key + keyStart[p], pay + payStart[p], key + keyStart[c], pay + payStart[c],
                                   fromTip[i], ais, &out[i]);
        }
        free(ais);
    }
}

void or_differ_batch(const OrModel *m, const uint32_t *key, const double *pay, const int64_t *keyStart,
                     const int64_t *payStart, int64_t n, const int32_t *idx1, const int32_t *idx2, uint8_t *out) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        int a = idx1[i], b = idx2[i];
        out[i] = (uint8_t)or_differ(m, key + keyStart[a], pay + payStart[a], keyStart[b] < 0 ? NULL : key + keyStart[b],
                                    keyStart[b] < 0 ? NULL : pay + payStart[b]);
    }
}

/* shorten() (:3721) applied in place to every result slot of a merge batch (what the callers that STORE a list do, :6201, :6267) */
void or_shorten_slots(const OrModel *m, int64_t n, uint32_t *key, double *pay, const int64_t *keyStart, const int64_t *payStart,
                      int32_t *nk, int32_t *np_, const int32_t *status) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        if (status && status[i] != 0) continue;
        int32_t a = 0, b = 0;
        or_shorten(m, key + keyStart[i], pay + payStart[i], key + keyStart[i], pay + payStart[i], &a, &b);
        nk[i] = a;
        np_[i] = b;
    }
}

/* n > 0: use n OpenMP threads from now on, whatever OMP_NUM_THREADS said at start-up (torchrun exports
 * OMP_NUM_THREADS=1 to its workers, which is not what a CPU baseline "on all host cores" means). */
void or_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int or_num_threads(void) {
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
#endif
    return n;
}

/* ==================================================================================================
 * SPR search: the per-node body of startTopologyUpdatesParallel (:9615-9711) and findBestParentTopology
 * (:6817-7724) with evaluatePlacement (:6790-6806), for the default feature set (no time tree, no HnZ, no SPRTA).
 *
 * Phase 2 of the reference (:7460-7639) re-visits every entry of bestNodes whose score >= originalLK - threshold.
 * Entries are only appended when score > (>=) bestLKdiff - threshold and bestLKdiff never drops below originalLK,
 * so EVERY entry qualifies; and an entry's evaluation depends on nothing but the lists it carries.  The phase-2
 * evaluation is therefore done here at the moment the entry would be appended, keeping the running arg-max in
 * discovery order (ties: the later entry wins, :7635) -- same results, no list has to outlive the walk.
 * ================================================================================================== */
typedef struct {
    int32_t nNodes, root;
    const int32_t *up, *child0, *child1; /* -1 = none */
    const double *dist;
    const uint8_t *isTip;     /* no children and no minor sequences */
    const int32_t *mutStart;  /* [nNodes+1] CSR into mut, or NULL when no node carries MAT mutations */
    const int32_t *mut;       /* triples (pos1, upNuc, downNuc) */
    const uint32_t *key;      /* list id = family*nNodes + node; 0 lower, 1 upRight, 2 upLeft, 3 totUp */
    const double *pay;
    const int64_t *keyStart, *payStart;
    const int32_t *nkeys;
} OrTree;

typedef struct {
    int32_t strictTopologyStopRules, allowedFailsTopology, deeperSearchForLongBranches, reserved;
    double thresholdLogLKtopology, thresholdTopologyPlacement, thresholdLogLKoptimizationTopology;
    double thresholdLogLKconsecutivePlacement, effectivelyNon0BLen, BLenThresholdDeeperSearch, defaultBLen;
} OrSearchParams;

typedef struct {
    int32_t placement; /* proposed re-attachment node or -1 */
    int32_t bestNode;  /* findBestParentTopology's bestNode (-1 when the search did not run) */
    int32_t status;    /* 0 ok, 1 search not needed, 2 aborted (the reference's try/except -> placement None), 3 scratch overflow */
    int32_t phase1;    /* candidate placements scored by the phase-1 appendProbNode calls (:7011 / :7223) */
    double improvement, bestCurrentLK, bestScore, bLenTop, bLenBottom, bLenAppend;
} OrSearchResult;

typedef struct {
    const uint32_t *k;
    const double *p;
    int nk;
} LRef;

/* The reference's tree is not quite frozen during a search round: a child of the root with a zero-length branch
 * has probVectTotUp == None until the first search that reaches it from below with needsUpdating computes and STORES
 * it (:7198-7200).  Before that, searches arriving with converged partials skip the node (:7207); afterwards they
 * score it -- so the reference's proposals depend on the order of searches within a worker.  `LazyTot` holds those
 * (at most two) lists.  mode 0 = lazy like the reference (single-threaded, caller's order), mode 1 = pre-filled
 * before any search (order-independent; what the device path does). */
typedef struct {
    int node[2], filled[2], nk[2];
    uint32_t *key[2];
    double *pay[2];
} LazyTot;

typedef struct {
    uint32_t *key;
    double *pay;
    size_t capK, capP, topK, topP;
    int overflow;
    double *ais;
    size_t capA;
    LazyTot *lazy;
} Scratch;

typedef struct {
    int t1, direction, needsUpdating, failedPasses;
    LRef passed, removed;
    int removedScratch; /* the removed list lives in scratch (may be shortened in place) */
    double distance, lastLK;
    size_t markK, markP;
} StackE;

#define OR_STACK_MAX 4096

static LRef lref_null(void) { LRef r = {NULL, NULL, 0}; return r; }

static LRef tree_list(const OrTree *t, int fam, int node) {
    int64_t id = (int64_t)fam * t->nNodes + node;
    LRef r;
    if (t->keyStart[id] < 0) return lref_null();
    r.k = t->key + t->keyStart[id];
    r.p = t->pay + t->payStart[id];
    r.nk = t->nkeys[id];
    return r;
}

static LRef totup_list(const OrTree *t, const Scratch *s, int node) {
    LRef r = tree_list(t, 3, node);
    if (r.k || !s->lazy) return r;
    for (int i = 0; i < 2; i++)
        if (s->lazy->node[i] == node && s->lazy->filled[i]) {
            r.k = s->lazy->key[i];
            r.p = s->lazy->pay[i];
            r.nk = s->lazy->nk[i];
        }
    return r;
}

static void lazy_fill(const OrModel *m, const OrTree *t, Scratch *s, int node, LRef vectUp) {
    LazyTot *z = s->lazy;
    if (!z) return;
    for (int i = 0; i < 2; i++)
        if (z->node[i] == node && !z->filled[i]) {
            LRef pv = tree_list(t, 0, node);
            size_t cap = (size_t)vectUp.nk + pv.nk + 4;
            z->key[i] = (uint32_t *)malloc(sizeof(uint32_t) * cap);
            z->pay[i] = (double *)malloc(sizeof(double) * 6 * cap);
            int32_t nk = 0, np = 0;
            int st = or_merge(m, vectUp.k, vectUp.p, t->dist[node] / 2, 0, pv.k, pv.p, t->dist[node] / 2, 0, 1, 0, 0, z->key[i], z->pay[i], &nk, &np, NULL);
            if (st == 0) { z->filled[i] = 1; z->nk[i] = nk; }
        }
}

static int sc_reserve(Scratch *s, size_t nk) {
    size_t k = (nk + 3) & ~(size_t)3;
    if (s->topK + k > s->capK || s->topP + 6 * k > s->capP) { s->overflow = 3; return 0; }
    return 1;
}

static LRef sc_commit(Scratch *s, int nk, int np) {
    LRef r = {s->key + s->topK, s->pay + s->topP, nk};
    s->topK += ((size_t)nk + 3) & ~(size_t)3;
    s->topP += ((size_t)np + 1) & ~(size_t)1;
    return r;
}

static int n_mut(const OrTree *t, int node) { return t->mutStart ? t->mutStart[node + 1] - t->mutStart[node] : 0; }

/* passGenomeListThroughBranch into scratch */
static LRef s_pass(const OrModel *m, const OrTree *t, Scratch *s, LRef v, int node, int dirIsUp) {
    int nm = n_mut(t, node);
    if (!sc_reserve(s, (size_t)v.nk + 2 * (size_t)nm + 2)) return lref_null();
    int32_t nk = 0, np = 0;
    or_pass_branch(m, v.k, v.p, t->mut + 3 * (size_t)t->mutStart[node], nm, dirIsUp, s->key + s->topK, s->pay + s->topP, &nk, &np);
    return sc_commit(s, nk, np);
}

static LRef s_merge(const OrModel *m, Scratch *s, LRef a, double b1, int t1, LRef b, double b2, int t2, int upDown) {
    if (!a.k || !b.k) { if (!s->overflow) s->overflow = 2; return lref_null(); } /* the reference would raise -> search aborted */
    if (!sc_reserve(s, (size_t)a.nk + b.nk)) return lref_null();
    int32_t nk = 0, np = 0;
    int st = or_merge(m, a.k, a.p, b1, t1, b.k, b.p, b2, t2, upDown ? 1 : 0, 0, 0, s->key + s->topK, s->pay + s->topP, &nk, &np, NULL);
    if (st != 0) return lref_null();
    return sc_commit(s, nk, np);
}

static void s_shorten_inplace(const OrModel *m, LRef *v) {
    int32_t nk = 0, np = 0;
    or_shorten(m, v->k, v->p, (uint32_t *)v->k, (double *)v->p, &nk, &np);
    v->nk = nk;
}

/* rootVector(probVect, bLen, isFromTip, tree, node) for node == root (:4916-4996) */
static LRef s_root_vector(const OrModel *m, const OrTree *t, Scratch *s, LRef v, double bLen, int isFromTip) {
    int root = t->root;
    if (n_mut(t, root)) v = s_pass(m, t, s, v, root, 1);
    if (!v.k || !sc_reserve(s, (size_t)v.nk)) return lref_null();
    int32_t nk = 0, np = 0;
    or_root_vector(m, v.k, v.p, bLen, isFromTip, s->key + s->topK, s->pay + s->topP, &nk, &np);
    LRef r = sc_commit(s, nk, np);
    if (n_mut(t, root)) r = s_pass(m, t, s, r, root, 0);
    if (r.k) s_shorten_inplace(m, &r);
    return r;
}

static LRef s_copy(Scratch *s, LRef v) {
    if (!sc_reserve(s, (size_t)v.nk)) return lref_null();
    /* payload size is implied by the keys */
    size_t np = 0;
    for (int i = 0; i < v.nk; i++) {
        uint32_t k = v.k[i];
        np += ((k >> 3) & 3u) + (((k & 7u) == 6) ? 4 : 0);
    }
    memcpy(s->key + s->topK, v.k, sizeof(uint32_t) * v.nk);
    memcpy(s->pay + s->topP, v.p, sizeof(double) * np);
    return sc_commit(s, v.nk, (int)np);
}

static double s_blen(const OrModel *m, Scratch *s, LRef P, LRef C, int fromTipC) {
    if (!P.k || !C.k) { if (!s->overflow) s->overflow = 2; return 0.0; }
    size_t need = (size_t)P.nk + C.nk + 1;
    if (need > s->capA) { s->overflow = 3; return 0.0; }
    double out = 0.0;
    or_blen(m, P.k, P.p, C.k, C.p, fromTipC, s->ais, &out); /* python False and 0.0 are both "zero length" to the callers */
    return out;
}

/* evaluatePlacement (:6790-6806); returns 0 ok, 1 when the reference would raise (a None list reaches the next call) */
static int eval_placement(const OrModel *m, const OrSearchParams *sp, Scratch *s, LRef midTot, LRef downVect, LRef upVect,
                          double distance, LRef removed, int isRemovedTip, int fromTip1, double *cost, double *bBottom,
                          double *bTop, double *bAppend) {
    if (!midTot.k || !downVect.k || !upVect.k) return 1;
    size_t mk = s->topK, mp = s->topP;
    double bestAppending = s_blen(m, s, midTot, removed, isRemovedTip);
    LRef midLower = s_merge(m, s, downVect, distance / 2, fromTip1, removed, bestAppending, isRemovedTip, 0);
    if (!midLower.k) return 1;
    double bestTop = s_blen(m, s, upVect, midLower, 0);
    LRef midTop = s_merge(m, s, upVect, bestTop, 0, removed, bestAppending, isRemovedTip, 1);
    if (!midTop.k) {
        if (s->overflow) return 1;
        bestTop = sp->defaultBLen * 0.1;
        midTop = s_merge(m, s, upVect, bestTop, 0, removed, bestAppending, isRemovedTip, 1);
        if (!midTop.k) return 1;
    }
    double bestBottom = s_blen(m, s, midTop, downVect, fromTip1);
    LRef newMid = s_merge(m, s, upVect, bestTop, 0, downVect, bestBottom, fromTip1, 1);
    if (!newMid.k) return 1;
    *cost = or_append(m, newMid.k, newMid.p, removed.k, removed.p, isRemovedTip, bestAppending);
    *bBottom = bestBottom;
    *bTop = bestTop;
    *bAppend = bestAppending;
    s->topK = mk;
    s->topP = mp;
    return s->overflow;
}

typedef struct {
    double bestScore;
    int bestNode;
    double bTop, bBottom, bAppend;
} Phase2;

/* one bestNodes entry, evaluated eagerly (:7460-7639 without HnZ / time / SPRTA) */
static int phase2_entry(const OrModel *m, const OrSearchParams *sp, Scratch *s, int t1, LRef midTot, LRef downVect, LRef upVect,
                        double distance, LRef removed, int isRemovedTip, int fromTip1, Phase2 *ph) {
    double cost, bB, bT, bA;
    if (eval_placement(m, sp, s, midTot, downVect, upVect, distance, removed, isRemovedTip, fromTip1, &cost, &bB, &bT, &bA)) return 1;
    double initialCost = or_append(m, upVect.k, upVect.p, downVect.k, downVect.p, fromTip1, distance);
    double newPartialCost = or_append(m, upVect.k, upVect.p, downVect.k, downVect.p, fromTip1, bB + bT);
    double optimizedScore = cost + newPartialCost - initialCost;
    if (optimizedScore >= ph->bestScore) {
        ph->bestNode = t1;
        ph->bestScore = optimizedScore;
        ph->bTop = bT;
        ph->bBottom = bB;
        ph->bAppend = bA;
    }
    return 0;
}

static int find_best_parent_topology(const OrModel *m, const OrTree *t, const OrSearchParams *sp, Scratch *s, int node, int child,
                                     double bestLKdiff, double removedBLen, Phase2 *ph, int *phase1) {
    const int32_t *up = t->up;
    const double *dist = t->dist;
    const double eff = sp->effectivelyNon0BLen;
#define CH(n, i) ((i) == 0 ? t->child0[n] : t->child1[n])
    StackE *stack = (StackE *)malloc(sizeof(StackE) * OR_STACK_MAX);
    int sp_n = 0, rc = 0;
    const int pruned = CH(node, child), sibling = CH(node, 1 - child);
    int bestNodeInit = sibling;
    /* the removed list is copied to scratch: the reference may shorten it in place (:7087) */
    LRef removedRel = s_copy(s, tree_list(t, 0, pruned));
    if (n_mut(t, pruned)) removedRel = s_pass(m, t, s, removedRel, pruned, 1);
    LRef bestRemoved = removedRel;
    if (n_mut(t, bestNodeInit)) bestRemoved = s_pass(m, t, s, bestRemoved, bestNodeInit, 0);
    const int isRemovedTip = t->isTip[pruned];
    const double originalLK = bestLKdiff;
    ph->bestNode = bestNodeInit;
    ph->bestScore = originalLK;
    if (up[node] >= 0) {
        int childUp;
        LRef vectUpUp;
        if (t->child0[up[node]] == node) { childUp = 1; vectUpUp = tree_list(t, 1, up[node]); }
        else { childUp = 2; vectUpUp = tree_list(t, 2, up[node]); }
        LRef probVect1 = tree_list(t, 0, bestNodeInit);
        if (n_mut(t, bestNodeInit)) probVect1 = s_pass(m, t, s, probVect1, bestNodeInit, 1);
        LRef removedRel1 = removedRel;
        if (n_mut(t, node)) {
            probVect1 = s_pass(m, t, s, probVect1, node, 1);
            removedRel1 = s_pass(m, t, s, removedRel, node, 1);
        }
        StackE e;
        memset(&e, 0, sizeof e);
        e.t1 = up[node]; e.direction = childUp; e.needsUpdating = 1; e.passed = probVect1;
        e.distance = dist[bestNodeInit] + dist[node]; e.lastLK = bestLKdiff; e.failedPasses = 0; e.removed = removedRel1; e.removedScratch = 1;
        e.markK = s->topK; e.markP = s->topP;
        stack[sp_n++] = e;
        if (n_mut(t, node)) vectUpUp = s_pass(m, t, s, vectUpUp, node, 0);
        removedRel1 = removedRel;
        if (n_mut(t, bestNodeInit)) {
            vectUpUp = s_pass(m, t, s, vectUpUp, bestNodeInit, 0);
            removedRel1 = s_pass(m, t, s, removedRel, bestNodeInit, 0);
        }
        e.t1 = bestNodeInit; e.direction = 0; e.passed = vectUpUp; e.removed = removedRel1;
        e.markK = s->topK; e.markP = s->topP;
        stack[sp_n++] = e;
        ph->bTop = dist[node]; ph->bBottom = dist[bestNodeInit]; ph->bAppend = removedBLen;
    } else {
        if (t->child0[bestNodeInit] >= 0) {
            int child1 = t->child0[bestNodeInit], child2 = t->child1[bestNodeInit];
            for (int which = 0; which < 2; which++) {
                int target = which == 0 ? child1 : child2, other = which == 0 ? child2 : child1;
                LRef vectUp1 = tree_list(t, 0, other);
                if (n_mut(t, other)) vectUp1 = s_pass(m, t, s, vectUp1, other, 1);
                vectUp1 = s_root_vector(m, t, s, vectUp1, dist[other], t->isTip[other]);
                LRef removedRel1 = bestRemoved;
                if (n_mut(t, target)) {
                    removedRel1 = s_pass(m, t, s, bestRemoved, target, 0);
                    vectUp1 = s_pass(m, t, s, vectUp1, target, 0);
                }
                StackE e;
                memset(&e, 0, sizeof e);
                e.t1 = target; e.direction = 0; e.needsUpdating = 1; e.passed = vectUp1; e.distance = dist[target];
                e.lastLK = bestLKdiff; e.failedPasses = 0; e.removed = removedRel1; e.removedScratch = 1;
                e.markK = s->topK; e.markP = s->topP;
                stack[sp_n++] = e;
            }
        }
        ph->bTop = 0.0; ph->bBottom = dist[bestNodeInit]; ph->bAppend = removedBLen;
    }
    if (s->overflow) { rc = s->overflow; goto done; }

    while (sp_n > 0) {
        StackE E = stack[--sp_n];
        s->topK = E.markK;
        s->topP = E.markP;
        const int t1 = E.t1, direction = E.direction;
        int needsUpdating = E.needsUpdating, failedPasses = E.failedPasses;
        LRef passed = E.passed, removed = E.removed;
        double distance = E.distance, lastLK = E.lastLK, midProb;
        if (direction == 0) {
            if (!(up[t1] == node || up[t1] < 0) && (dist[t1] > eff || up[up[t1]] < 0)) {
                LRef midTot;
                if (needsUpdating) {
                    midTot = s_merge(m, s, passed, distance / 2, 0, tree_list(t, 0, t1), distance / 2, t->isTip[t1], 1);
                    if (s->overflow) { rc = s->overflow; goto done; }
                    if (!midTot.k) continue;
                    LRef stored = totup_list(t, s, t1);
                    if (!or_differ(m, midTot.k, midTot.p, stored.k, stored.p)) needsUpdating = 0;
                } else {
                    midTot = totup_list(t, s, t1);
                    distance = dist[t1];
                }
                if (!midTot.k) continue;
                LRef vectUp = lref_null();
                if (sp->deeperSearchForLongBranches && distance > sp->BLenThresholdDeeperSearch) {
                    LRef midBottom = tree_list(t, 0, t1);
                    vectUp = (t1 == t->child0[up[t1]]) ? tree_list(t, 1, up[t1]) : tree_list(t, 2, up[t1]);
                    if (n_mut(t, t1)) vectUp = s_pass(m, t, s, vectUp, t1, 0);
                    double bB, bT, bA;
                    if (eval_placement(m, sp, s, midTot, midBottom, vectUp, distance, removed, isRemovedTip, t->isTip[t1], &midProb, &bB, &bT, &bA)) {
                        rc = s->overflow ? s->overflow : 2;
                        goto done;
                    }
                } else {
                    midProb = or_append(m, midTot.k, midTot.p, removed.k, removed.p, isRemovedTip, removedBLen);
                    (*phase1)++;
                }
                if (getenv("MAPLE_ORACLE_TRACE")) fprintf(stderr, "cand %d dir %d nu %d midProb %.17g fails %d lastLK %.17g best %.17g\n", t1, direction, needsUpdating, midProb, failedPasses, lastLK, bestLKdiff);
                if (midProb > bestLKdiff - sp->thresholdLogLKoptimizationTopology) { /* :7071 */
                    LRef upV, downV, mt;
                    double dd;
                    if (needsUpdating) { upV = passed; downV = tree_list(t, 0, t1); dd = distance; mt = midTot; }
                    else {
                        upV = (t1 == t->child0[up[t1]]) ? tree_list(t, 1, up[t1]) : tree_list(t, 2, up[t1]);
                        if (n_mut(t, t1)) upV = s_pass(m, t, s, upV, t1, 0);
                        downV = tree_list(t, 0, t1); dd = dist[t1]; mt = totup_list(t, s, t1);
                    }
                    if (phase2_entry(m, sp, s, t1, mt, downV, upV, dd, removed, isRemovedTip, t->isTip[t1], ph)) {
                        rc = s->overflow ? s->overflow : 2;
                        goto done;
                    }
                }
                if (midProb > bestLKdiff) {
                    bestLKdiff = midProb;
                    failedPasses = 0;
                    if (E.removedScratch) s_shorten_inplace(m, &removed); /* :7087 */
                } else if (midProb < (lastLK - sp->thresholdLogLKconsecutivePlacement)) failedPasses++;
            } else midProb = lastLK;

            int traverse = 0;
            if (sp->strictTopologyStopRules) {
                if (failedPasses <= sp->allowedFailsTopology && midProb > (bestLKdiff - sp->thresholdLogLKtopology) && t->child0[t1] >= 0) traverse = 1;
            } else if (failedPasses <= sp->allowedFailsTopology || midProb > (bestLKdiff - sp->thresholdLogLKtopology)) {
                if (t->child0[t1] >= 0) traverse = 1;
            }
            if (traverse) {
                for (int which = 0; which < 2; which++) { /* child 0 is pushed first, so child 1 is explored first */
                    int child1 = CH(t1, which), otherChild = CH(t1, 1 - which);
                    LRef vUp;
                    if (needsUpdating) {
                        LRef otherPV = tree_list(t, 0, otherChild);
                        if (n_mut(t, otherChild)) otherPV = s_pass(m, t, s, otherPV, otherChild, 1);
                        vUp = s_merge(m, s, passed, distance, 0, otherPV, dist[otherChild], t->isTip[otherChild], 1);
                        if (s->overflow) { rc = s->overflow; goto done; }
                    } else vUp = which == 0 ? tree_list(t, 1, t1) : tree_list(t, 2, t1);
                    if (vUp.k) {
                        LRef removed1 = removed;
                        int rs = E.removedScratch;
                        if (n_mut(t, child1)) { removed1 = s_pass(m, t, s, removed, child1, 0); rs = 1; }
                        if (needsUpdating && n_mut(t, child1)) vUp = s_pass(m, t, s, vUp, child1, 0);
                        if (sp_n >= OR_STACK_MAX || s->overflow) { rc = s->overflow ? s->overflow : 3; goto done; }
                        StackE e;
                        memset(&e, 0, sizeof e);
                        e.t1 = child1; e.direction = 0; e.needsUpdating = needsUpdating; e.passed = needsUpdating ? vUp : lref_null();
                        e.distance = dist[child1]; e.lastLK = midProb; e.failedPasses = failedPasses; e.removed = removed1; e.removedScratch = rs;
                        e.markK = s->topK; e.markP = s->topP;
                        stack[sp_n++] = e;
                    }
                }
            }
        } else { /* crawling up from child to parent (:7179-7429) */
            const int otherChild = CH(t1, 2 - direction);
            LRef midBottom = lref_null(), vectUp = lref_null();
            if (up[t1] >= 0 && (dist[t1] > eff || up[up[t1]] < 0)) {
                LRef midTot;
                if (needsUpdating) {
                    LRef otherPV = tree_list(t, 0, otherChild);
                    if (n_mut(t, otherChild)) otherPV = s_pass(m, t, s, otherPV, otherChild, 1);
                    midBottom = s_merge(m, s, passed, distance, 0, otherPV, dist[otherChild], t->isTip[otherChild], 0);
                    if (s->overflow) { rc = s->overflow; goto done; }
                    if (!midBottom.k) continue;
                    vectUp = (t1 == t->child0[up[t1]]) ? tree_list(t, 1, up[t1]) : tree_list(t, 2, up[t1]);
                    if (n_mut(t, t1)) vectUp = s_pass(m, t, s, vectUp, t1, 0);
                    midTot = s_merge(m, s, vectUp, dist[t1] / 2, 0, midBottom, dist[t1] / 2, 0, 1);
                    if (s->overflow) { rc = s->overflow; goto done; }
                    if (!totup_list(t, s, t1).k) lazy_fill(m, t, s, t1, vectUp); /* :7198-7200 */
                    if (!midTot.k) continue;
                    LRef stored = totup_list(t, s, t1);
                    if (!or_differ(m, midTot.k, midTot.p, stored.k, stored.p)) needsUpdating = 0;
                } else midTot = totup_list(t, s, t1);
                if (!midTot.k) continue;
                if (sp->deeperSearchForLongBranches && dist[t1] > sp->BLenThresholdDeeperSearch) {
                    if (!needsUpdating) {
                        midBottom = tree_list(t, 0, t1);
                        vectUp = (t1 == t->child0[up[t1]]) ? tree_list(t, 1, up[t1]) : tree_list(t, 2, up[t1]);
                        if (n_mut(t, t1)) vectUp = s_pass(m, t, s, vectUp, t1, 0);
                    }
                    double bB, bT, bA;
                    if (eval_placement(m, sp, s, midTot, midBottom, vectUp, dist[t1], removed, isRemovedTip, 0, &midProb, &bB, &bT, &bA)) {
                        rc = s->overflow ? s->overflow : 2;
                        goto done;
                    }
                } else {
                    midProb = or_append(m, midTot.k, midTot.p, removed.k, removed.p, isRemovedTip, removedBLen);
                    (*phase1)++;
                }
                if (getenv("MAPLE_ORACLE_TRACE")) fprintf(stderr, "cand %d dir %d nu %d midProb %.17g fails %d lastLK %.17g best %.17g\n", t1, direction, needsUpdating, midProb, failedPasses, lastLK, bestLKdiff);
                if (midProb >= (bestLKdiff - sp->thresholdLogLKoptimizationTopology)) { /* :7293 */
                    LRef upV, downV, mt;
                    if (needsUpdating) { upV = vectUp; downV = midBottom; mt = midTot; }
                    else {
                        upV = (t1 == t->child0[up[t1]]) ? tree_list(t, 1, up[t1]) : tree_list(t, 2, up[t1]);
                        if (n_mut(t, t1)) upV = s_pass(m, t, s, upV, t1, 0);
                        downV = tree_list(t, 0, t1); mt = totup_list(t, s, t1);
                    }
                    if (phase2_entry(m, sp, s, t1, mt, downV, upV, dist[t1], removed, isRemovedTip, t->isTip[t1], ph)) {
                        rc = s->overflow ? s->overflow : 2;
                        goto done;
                    }
                }
                if (midProb > bestLKdiff) { bestLKdiff = midProb; failedPasses = 0; }
                else if (midProb < (lastLK - sp->thresholdLogLKconsecutivePlacement)) failedPasses++;
            } else midProb = lastLK;

            int keep = 0;
            if (sp->strictTopologyStopRules) {
                if (failedPasses <= sp->allowedFailsTopology && midProb > (bestLKdiff - sp->thresholdLogLKtopology)) keep = 1;
            } else if (failedPasses <= sp->allowedFailsTopology || midProb > (bestLKdiff - sp->thresholdLogLKtopology)) keep = 1;
            if (keep) {
                if (up[t1] >= 0) {
                    int upChild;
                    LRef vectUpUp = lref_null(), vUp;
                    if (t1 == t->child0[up[t1]]) { upChild = 0; if (needsUpdating) vectUpUp = tree_list(t, 1, up[t1]); }
                    else { upChild = 1; if (needsUpdating) vectUpUp = tree_list(t, 2, up[t1]); }
                    if (needsUpdating) {
                        if (n_mut(t, t1)) vectUpUp = s_pass(m, t, s, vectUpUp, t1, 0);
                        vUp = s_merge(m, s, vectUpUp, dist[t1], 0, passed, distance, 0, 1);
                        if (s->overflow) { rc = s->overflow; goto done; }
                    } else vUp = direction == 1 ? tree_list(t, 2, t1) : tree_list(t, 1, t1);
                    if (!vUp.k) continue;
                    {
                        LRef removed1 = removed;
                        int rs = E.removedScratch;
                        if (n_mut(t, otherChild)) { removed1 = s_pass(m, t, s, removed, otherChild, 0); rs = 1; }
                        if (needsUpdating && n_mut(t, otherChild)) vUp = s_pass(m, t, s, vUp, otherChild, 0);
                        if (sp_n >= OR_STACK_MAX || s->overflow) { rc = s->overflow ? s->overflow : 3; goto done; }
                        StackE e;
                        memset(&e, 0, sizeof e);
                        e.t1 = otherChild; e.direction = 0; e.needsUpdating = needsUpdating; e.passed = needsUpdating ? vUp : lref_null();
                        e.distance = dist[otherChild]; e.lastLK = midProb; e.failedPasses = failedPasses; e.removed = removed1; e.removedScratch = rs;
                        e.markK = s->topK; e.markP = s->topP;
                        stack[sp_n++] = e;
                    }
                    if (needsUpdating && !midBottom.k) {
                        LRef otherPV = tree_list(t, 0, otherChild);
                        if (n_mut(t, otherChild)) otherPV = s_pass(m, t, s, otherPV, otherChild, 1);
                        midBottom = s_merge(m, s, passed, distance, 0, otherPV, dist[otherChild], t->isTip[otherChild], 0);
                        if (s->overflow) { rc = s->overflow; goto done; }
                        if (!midBottom.k) continue;
                    }
                    {
                        LRef removed1 = removed;
                        int rs = E.removedScratch;
                        if (n_mut(t, t1)) { removed1 = s_pass(m, t, s, removed, t1, 1); rs = 1; }
                        if (needsUpdating && n_mut(t, t1)) midBottom = s_pass(m, t, s, midBottom, t1, 1);
                        if (sp_n >= OR_STACK_MAX || s->overflow) { rc = s->overflow ? s->overflow : 3; goto done; }
                        StackE e;
                        memset(&e, 0, sizeof e);
                        e.t1 = up[t1]; e.direction = upChild + 1; e.needsUpdating = needsUpdating; e.passed = needsUpdating ? midBottom : lref_null();
                        e.distance = dist[t1]; e.lastLK = midProb; e.failedPasses = failedPasses; e.removed = removed1; e.removedScratch = rs;
                        e.markK = s->topK; e.markP = s->topP;
                        stack[sp_n++] = e;
                    }
                } else { /* t1 is the root (:7406-7429) */
                    LRef vUp = lref_null();
                    if (needsUpdating) {
                        vUp = s_root_vector(m, t, s, passed, distance, 0);
                        if (n_mut(t, otherChild)) vUp = s_pass(m, t, s, vUp, otherChild, 0);
                    }
                    LRef removed1 = removed;
                    int rs = E.removedScratch;
                    if (n_mut(t, otherChild)) { removed1 = s_pass(m, t, s, removed, otherChild, 0); rs = 1; }
                    if (sp_n >= OR_STACK_MAX || s->overflow) { rc = s->overflow ? s->overflow : 3; goto done; }
                    StackE e;
                    memset(&e, 0, sizeof e);
                    e.t1 = otherChild; e.direction = 0; e.needsUpdating = needsUpdating; e.passed = vUp;
                    e.distance = dist[otherChild]; e.lastLK = midProb; e.failedPasses = failedPasses; e.removed = removed1; e.removedScratch = rs;
                    e.markK = s->topK; e.markP = s->topP;
                    stack[sp_n++] = e;
                }
            }
        }
    }
done:
    free(stack);
    return rc;
#undef CH
}

/* the per-node body of startTopologyUpdatesParallel (:9619-9711) */
void or_search_node(const OrModel *m, const OrTree *t, const OrSearchParams *sp, int node, Scratch *s, OrSearchResult *r) {
    memset(r, 0, sizeof *r);
    r->placement = -1;
    r->bestNode = -1;
    r->status = 1;
    if (t->up[node] < 0) return;
    s->topK = s->topP = 0;
    s->overflow = 0;
    const int parent = t->up[node];
    const int child = (t->child0[parent] == node) ? 0 : 1;
    LRef vectUp = child == 0 ? tree_list(t, 1, parent) : tree_list(t, 2, parent);
    if (n_mut(t, node)) vectUp = s_pass(m, t, s, vectUp, node, 0);
    const double bestCurrenBLen = t->dist[node];
    LRef own = tree_list(t, 0, node);
    if (!vectUp.k || !own.k) { r->status = 2; return; }
    const double bestCurrentLK = or_append(m, vectUp.k, vectUp.p, own.k, own.p, t->isTip[node], bestCurrenBLen);
    r->bestCurrentLK = bestCurrentLK;
    if (!(bestCurrentLK < sp->thresholdTopologyPlacement || t->dist[node] != 0.0)) return;
    Phase2 ph;
    memset(&ph, 0, sizeof ph);
    int phase1 = 0;
    s->topK = s->topP = 0;
    int rc = find_best_parent_topology(m, t, sp, s, parent, child, bestCurrentLK, bestCurrenBLen, &ph, &phase1);
    r->phase1 = phase1;
    r->status = rc;
    if (rc != 0) return;
    r->bestNode = ph.bestNode;
    r->bestScore = ph.bestScore;
    r->bLenTop = ph.bTop;
    r->bLenBottom = ph.bBottom;
    r->bLenAppend = ph.bAppend;
    if (ph.bestScore + sp->thresholdTopologyPlacement > bestCurrentLK) { /* :9681-9702 */
        int updated = 1, topNode = t->up[node];
        if (ph.bestNode == topNode) updated = 0;
        while (t->dist[topNode] == 0.0 && t->up[topNode] >= 0) topNode = t->up[topNode];
        if (ph.bestNode == topNode && ph.bBottom == 0.0) updated = 0;
        int sibling = child == 0 ? t->child1[parent] : t->child0[parent];
        if (ph.bestNode == sibling) updated = 0;
        if (t->up[ph.bestNode] == sibling && ph.bTop == 0.0) updated = 0;
        if (updated) {
            r->improvement = ph.bestScore - bestCurrentLK;
            r->placement = ph.bestNode;
        }
    }
}

void or_search_batch(const OrModel *m, const OrTree *t, const OrSearchParams *sp, int64_t n, const int32_t *nodes,
                     int64_t scratchKeys, int32_t lazyMode, OrSearchResult *out) {
    LazyTot lazy;
    memset(&lazy, 0, sizeof lazy);
    lazy.node[0] = lazy.node[1] = -1;
    if (t->child0[t->root] >= 0) {
        int ch[2] = {t->child0[t->root], t->child1[t->root]};
        for (int i = 0; i < 2; i++)
            if (t->dist[ch[i]] == 0.0 && t->keyStart[3 * (int64_t)t->nNodes + ch[i]] < 0) lazy.node[i] = ch[i];
    }
    if (lazyMode == 1) { /* pre-fill */
        Scratch s0;
        memset(&s0, 0, sizeof s0);
        s0.capK = (size_t)scratchKeys; s0.capP = 6 * s0.capK;
        s0.key = (uint32_t *)malloc(sizeof(uint32_t) * s0.capK);
        s0.pay = (double *)malloc(sizeof(double) * s0.capP);
        s0.lazy = &lazy;
        for (int i = 0; i < 2; i++)
            if (lazy.node[i] >= 0) {
                int c = lazy.node[i];
                LRef vectUp = (c == t->child0[t->root]) ? tree_list(t, 1, t->root) : tree_list(t, 2, t->root);
                if (n_mut(t, c)) vectUp = s_pass(m, t, &s0, vectUp, c, 0);
                if (vectUp.k) lazy_fill(m, t, &s0, c, vectUp);
            }
        free(s0.key);
        free(s0.pay);
    }
#pragma omp parallel if (lazyMode == 1)
    {
        Scratch s;
        memset(&s, 0, sizeof s);
        s.capK = (size_t)scratchKeys;
        s.capP = 6 * s.capK;
        s.capA = s.capK;
        s.key = (uint32_t *)malloc(sizeof(uint32_t) * s.capK);
        s.pay = (double *)malloc(sizeof(double) * s.capP);
        s.ais = (double *)malloc(sizeof(double) * s.capA);
        s.lazy = &lazy;
#pragma omp for schedule(dynamic, 8)
        for (int64_t i = 0; i < n; i++) or_search_node(m, t, sp, nodes[i], &s, &out[i]);
        free(s.key);
        free(s.pay);
        free(s.ais);
    }
    for (int i = 0; i < 2; i++) {
        free(lazy.key[i]);
        free(lazy.pay[i]);
    }
}

/* ==================================================================================================
 * Placement of a new sample on a frozen tree: findBestParentForNewSample (:7912-8292), default feature set
 * (computePlacementSupportOnly=False, no HnZ, no time tree), with isMinorSequence (:5919-6004).
 * ================================================================================================== */

/* isMinorSequence(probVect1, probVect2, onlyFindIdentical): 0 = neither contains the other, 1 = list 2 is identical to or
 * less informative than list 1, 2 = list 1 is less informative than list 2 */
int or_is_minor(const OrModel *m, const uint32_t *k1, const double *p1, const uint32_t *k2, const double *p2, int onlyFindIdentical) {
    const int lRef = m->lRef;
    Cur e1, e2;
    cur_init(&e1, k1, p1);
    cur_init(&e2, k2, p2);
    int pos = 0, found1bigger = 0, found2bigger = 0;
    for (;;) {
        if (e1.type != e2.type) {
            if (onlyFindIdentical) return 0;
            else if (e1.type == 5) {
                if (e2.type == 4) pos = e1.end < e2.end ? e1.end : e2.end;
                else pos += 1;
                found2bigger = 1;
            } else if (e2.type == 5) {
                if (e1.type == 4) pos = e1.end < e2.end ? e1.end : e2.end;
                else pos += 1;
                found1bigger = 1;
            } else if (e1.type == 6) {
                int i2 = (e2.type == 4) ? e1.nuc : e2.type;
                if (e1.vec[i2] > 0.1) found2bigger = 1;
                else return 0;
                pos += 1;
            } else if (e2.type == 6) {
                int i1 = (e1.type == 4) ? e2.nuc : e1.type;
                if (e2.vec[i1] > 0.1) found1bigger = 1;
                else return 0;
                pos += 1;
            } else return 0;
        } else if (e1.type == 6) {
            for (int j = 0; j < 4; j++) {
                if (onlyFindIdentical) {
                    if (e2.vec[j] != e1.vec[j]) return 0;
                } else if (e2.vec[j] > 0.1 && e1.vec[j] < 0.1) found1bigger = 1;
                else if (e1.vec[j] > 0.1 && e2.vec[j] < 0.1) found2bigger = 1;
            }
            pos += 1;
        } else {
            if (e1.type < 4) pos += 1;
            else pos = e1.end < e2.end ? e1.end : e2.end;
        }
        if (found1bigger && found2bigger) return 0;
        if (pos == lRef) break;
        if (e1.type < 4 || e1.type == 6 || pos == e1.end) cur_next(&e1);
        if (e2.type < 4 || e2.type == 6 || pos == e2.end) cur_next(&e2);
    }
    if (found1bigger) return found2bigger ? 0 : 1;
    return found2bigger ? 2 : 1;
}

typedef struct {
    int32_t strictStopRules, allowedFails, deeperSearchForLongBranches, onlyFindIdentical; /* onlyFindIdentical: any error-rate option, --supportFor0Branches or --HnZ (:7936) */
    double thresholdLogLK, thresholdLogLKoptimization, thresholdLogLKconsecutivePlacement;
    double effectivelyNon0BLen, BLenThresholdDeeperSearch, oneMutBLen;
} OrPlaceParams;

typedef struct {
    int32_t bestNode;
    int32_t status;   /* 0 placed (bestNode, bestScore, lengths); 1 absorbed as a minor sequence of leaf bestNode (:7949, :8002: the
                         reference returns (node, 1.0, None, diffs)); 2 aborted (the reference would raise); 3 scratch overflow */
    int32_t phase1;   /* candidate branches scored in the walk (:8033 / :8050) */
    int32_t missedMinors; /* leaves found strictly less informative than the sample (:7960, :8004) */
    double bestScore, bLenTop, bLenBottom, bLenAppend; /* python False in bestBranchLengths is 0.0 here */
} OrPlaceResult;

typedef struct { int t1, failedPasses; double parentLK; LRef diffs; } PlaceStackE;
typedef struct { int t1; double score; LRef diffs; } PlaceBest;

void or_place_sample(const OrModel *m, const OrTree *t, const OrPlaceParams *pp, const uint32_t *dk, const double *dp, int dnk, Scratch *s,
                     OrPlaceResult *r) {
    memset(r, 0, sizeof *r);
    s->topK = s->topP = 0;
    s->overflow = 0;
    const int root = t->root;
    const double eff = pp->effectivelyNon0BLen, one = pp->oneMutBLen;
    LRef in = {dk, dp, dnk};
    LRef diffs = s_copy(s, in); /* shorten() works in place on it (:8065) */
    if (!diffs.k) { r->status = 3; return; }
    if (n_mut(t, root)) diffs = s_pass(m, t, s, diffs, root, 0);
    int bestNode = root;
    double bTop = 0.0, bBottom = 0.0, bAppend = one; /* (False, False, oneMutBLen) */
    PlaceStackE *stack = (PlaceStackE *)malloc(sizeof(PlaceStackE) * (size_t)(t->nNodes + 4));
    PlaceBest *best = NULL;
    size_t nBest = 0, capBest = 0;
    int sp = 0;
#define PLACE_FAIL(code) do { r->status = (code); goto done; } while (0)
    if (t->child0[root] < 0) {
        LRef pv = tree_list(t, 0, root);
        int cmp = or_is_minor(m, pv.k, pv.p, diffs.k, diffs.p, pp->onlyFindIdentical);
        if (cmp == 1) { r->bestNode = root; r->bestScore = 1.0; PLACE_FAIL(1); }
        else if (cmp == 2) r->missedMinors++;
    }
    {
        LRef rootVect = s_root_vector(m, t, s, tree_list(t, 0, root), 0.0, 0);
        if (!rootVect.k) PLACE_FAIL(s->overflow ? s->overflow : 2);
        double bestLKdiff = or_append(m, rootVect.k, rootVect.p, diffs.k, diffs.p, 1, one);
        const double originalLKdiff = bestLKdiff;
        if (t->child0[root] >= 0) {
            int ch[2] = {t->child0[root], t->child1[root]};
            for (int i = 0; i < 2; i++) {
                LRef dc = diffs;
                if (n_mut(t, ch[i])) dc = s_pass(m, t, s, diffs, ch[i], 0);
                if (!dc.k) PLACE_FAIL(3);
                stack[sp].t1 = ch[i]; stack[sp].parentLK = bestLKdiff; stack[sp].failedPasses = 0; stack[sp].diffs = dc; sp++;
            }
        }
        while (sp > 0) {
            PlaceStackE E = stack[--sp];
            const int t1 = E.t1;
            int failedPasses = E.failedPasses;
            LRef d = E.diffs;
            double LKdiff;
            if (t->child0[t1] < 0) { /* a leaf: is the new sample identical to / contained in it? (:7975-8005) */
                LRef pv = tree_list(t, 0, t1);
                int cmp = or_is_minor(m, pv.k, pv.p, d.k, d.p, pp->onlyFindIdentical);
                if (cmp == 1) { r->bestNode = t1; r->bestScore = 1.0; PLACE_FAIL(1); }
                else if (cmp == 2) r->missedMinors++;
            }
            if (t->dist[t1] > eff && t->up[t1] >= 0) {
                double bestTopLength, bestBottomLength, bestAppendingLength;
                if (pp->deeperSearchForLongBranches && t->dist[t1] > pp->BLenThresholdDeeperSearch) {
                    const int par = t->up[t1];
                    LRef upVect = (t1 == t->child0[par]) ? tree_list(t, 1, par) : tree_list(t, 2, par);
                    if (n_mut(t, t1)) upVect = s_pass(m, t, s, upVect, t1, 0);
                    const int isTip = t->isTip[t1];
                    LRef pv = tree_list(t, 0, t1);
                    const size_t mk = s->topK, mp = s->topP;
                    bestAppendingLength = one;
                    LRef midLower = s_merge(m, s, pv, t->dist[t1] / 2, isTip, d, bestAppendingLength, 1, 0);
                    if (!midLower.k) PLACE_FAIL(s->overflow ? s->overflow : 2);
                    bestTopLength = s_blen(m, s, upVect, midLower, 0);
                    LRef midTop = s_merge(m, s, upVect, bestTopLength, 0, d, bestAppendingLength, 1, 1);
                    if (!midTop.k) PLACE_FAIL(s->overflow ? s->overflow : 2);
                    bestBottomLength = s_blen(m, s, midTop, pv, isTip);
                    LRef newMid = s_merge(m, s, upVect, bestTopLength, 0, pv, bestBottomLength, isTip, 1);
                    if (!newMid.k) PLACE_FAIL(s->overflow ? s->overflow : 2);
                    LKdiff = or_append(m, newMid.k, newMid.p, d.k, d.p, 1, bestAppendingLength);
                    if (!n_mut(t, t1)) { s->topK = mk; s->topP = mp; }
                } else {
                    LRef tot = tree_list(t, 3, t1);
                    if (!tot.k) PLACE_FAIL(2);
                    LKdiff = or_append(m, tot.k, tot.p, d.k, d.p, 1, one);
                    bestBottomLength = t->dist[t1] / 2;
                    bestTopLength = t->dist[t1] / 2;
                    bestAppendingLength = one;
                }
                r->phase1++;
                if (LKdiff >= bestLKdiff) {
                    s_shorten_inplace(m, &d); /* :8065 */
                    bestLKdiff = LKdiff;
                    bestNode = t1;
                    failedPasses = 0;
                    if (nBest == capBest) { capBest = capBest ? 2 * capBest : 64; best = (PlaceBest *)realloc(best, sizeof(PlaceBest) * capBest); }
                    best[nBest].t1 = t1; best[nBest].score = LKdiff; best[nBest].diffs = d; nBest++;
                    bTop = bestTopLength; bBottom = bestBottomLength / 2; bAppend = bestAppendingLength;
                } else if (LKdiff > bestLKdiff - pp->thresholdLogLKoptimization) {
                    if (nBest == capBest) { capBest = capBest ? 2 * capBest : 64; best = (PlaceBest *)realloc(best, sizeof(PlaceBest) * capBest); }
                    best[nBest].t1 = t1; best[nBest].score = LKdiff; best[nBest].diffs = d; nBest++;
                }
                if (LKdiff < (E.parentLK - pp->thresholdLogLKconsecutivePlacement)) failedPasses++;
            } else LKdiff = E.parentLK;
            int go;
            if (pp->strictStopRules) go = failedPasses <= pp->allowedFails && LKdiff > (bestLKdiff - pp->thresholdLogLK);
            else go = failedPasses <= pp->allowedFails || LKdiff > (bestLKdiff - pp->thresholdLogLK);
            if (go && t->child0[t1] >= 0) {
                int ch[2] = {t->child0[t1], t->child1[t1]};
                for (int i = 0; i < 2; i++) {
                    LRef dc = d;
                    if (n_mut(t, ch[i])) dc = s_pass(m, t, s, d, ch[i], 0);
                    if (!dc.k) PLACE_FAIL(3);
                    stack[sp].t1 = ch[i]; stack[sp].parentLK = LKdiff; stack[sp].failedPasses = failedPasses; stack[sp].diffs = dc; sp++;
                }
            }
        }
        /* refinement of every branch within thresholdLogLKoptimization of the best (:8109-8187) */
        double bestScore = bestLKdiff;
        for (size_t i = 0; i < nBest; i++) {
            if (!(best[i].score >= bestLKdiff - pp->thresholdLogLKoptimization)) continue;
            const int node = best[i].t1, par = t->up[node];
            LRef d = best[i].diffs;
            LRef upVect = (node == t->child0[par]) ? tree_list(t, 1, par) : tree_list(t, 2, par);
            const size_t mk = s->topK, mp = s->topP;
            if (n_mut(t, node)) upVect = s_pass(m, t, s, upVect, node, 0);
            const int isTip = t->isTip[node];
            LRef pv = tree_list(t, 0, node), tot = tree_list(t, 3, node);
            if (!upVect.k || !pv.k || !tot.k) PLACE_FAIL(s->overflow ? s->overflow : 2);
            const double bestAppendingLength = s_blen(m, s, tot, d, 1);
            LRef midLower = s_merge(m, s, pv, t->dist[node] / 2, isTip, d, bestAppendingLength, 1, 0);
            if (!midLower.k) PLACE_FAIL(s->overflow ? s->overflow : 2);
            double bestTopLength = s_blen(m, s, upVect, midLower, 0);
            LRef midTop = s_merge(m, s, upVect, bestTopLength, 0, d, bestAppendingLength, 1, 1);
            if (!midTop.k) PLACE_FAIL(s->overflow ? s->overflow : 2);
            double bestBottomLength = s_blen(m, s, midTop, pv, isTip);
            LRef newMid = s_merge(m, s, upVect, bestTopLength, 0, pv, bestBottomLength, isTip, 1);
            if (!newMid.k) PLACE_FAIL(s->overflow ? s->overflow : 2);
            const double appendingCost = or_append(m, newMid.k, newMid.p, d.k, d.p, 1, bestAppendingLength);
            const double initialCost = or_append(m, upVect.k, upVect.p, pv.k, pv.p, isTip, t->dist[node]);
            const double newPartialCost = or_append(m, upVect.k, upVect.p, pv.k, pv.p, isTip, bestBottomLength + bestTopLength);
            const double optimizedScore = appendingCost + newPartialCost - initialCost;
            if (optimizedScore >= bestScore) {
                bestNode = node;
                bestScore = optimizedScore;
                bTop = bestTopLength; bBottom = bestBottomLength; bAppend = bestAppendingLength;
            }
            s->topK = mk;
            s->topP = mp;
            if (s->overflow) PLACE_FAIL(s->overflow);
        }
        if (bestScore == -INFINITY) bestScore = originalLKdiff;
        r->bestNode = bestNode;
        r->bestScore = bestScore;
        r->bLenTop = bTop; r->bLenBottom = bBottom; r->bLenAppend = bAppend;
        r->status = s->overflow ? s->overflow : 0;
    }
done:
#undef PLACE_FAIL
    free(stack);
    free(best);
}

/* Samples are independent on a frozen tree (the reference's process_chunk / joblib seam, :11190-11287). */
void or_place_batch(const OrModel *m, const OrTree *t, const OrPlaceParams *pp, int64_t n, const uint32_t *key, const double *pay,
                    const int64_t *keyStart, const int64_t *payStart, const int32_t *nkeys, int64_t scratchKeys, OrPlaceResult *out) {
#pragma omp parallel
    {
        Scratch s;
        memset(&s, 0, sizeof s);
        s.capK = (size_t)scratchKeys;
        s.capP = 6 * s.capK;
        s.capA = s.capK;
        s.key = (uint32_t *)malloc(sizeof(uint32_t) * s.capK);
        s.pay = (double *)malloc(sizeof(double) * s.capP);
        s.ais = (double *)malloc(sizeof(double) * s.capA);
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = 0; i < n; i++)
            or_place_sample(m, t, pp, key + keyStart[i], pay + payStart[i], nkeys[i], &s, &out[i]);
        free(s.key);
        free(s.pay);
        free(s.ais);
    }
}
