/* TEST INFRASTRUCTURE ONLY -- the oracle. Never linked into, imported by, or called from the product
 * path (genarchbench_b200/, libbsw_gpu.so). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * Plain-C restatement of the reference's banded Smith-Waterman seed extension for ONE pair, written
 * from the semantics of /root/reference/benchmarks/bsw/src/bandedSWA.cpp:
 *   - recurrence, first row/column, band clamp, row-max / last-argmax, gscore, m==0 exit, beg/end
 *     shrink, outputs:                           scalarBandedSWA, bandedSWA.cpp:132-253
 *   - z-drop test WITHOUT the gap-extend factor  ZSCORE16, bandedSWA.cpp:1889-1902 (the vector kernel
 *     computes `insdel` and never uses it); this is what getScores16 -- the function the driver calls
 *     at main_banded.cpp:345 -- actually does.  zdrop_rule=1 selects the scalar rule (:226-231).
 *   - per-pair band                              smithWatermanBatchWrapper16, bandedSWA.cpp:2898-2919
 *     (uint16 arithmetic, integer division before the +1.0)
 *   - row budget                                 smithWaterman512_16, bandedSWA.cpp:3035-3036,3130-3144
 *     (a lane stops once i+1 > min(len2 + band, len1), or once its column range is empty)
 *   - the vector path scores an ambiguous base with the hard-coded DEFAULT_AMBIG = -1 whatever the
 *     matrix says (ctor, bandedSWA.cpp:65; bandedSWA.h:61) and applies the z-drop test even when
 *     zdrop == 0 (ZSCORE16 has no `zdrop > 0` guard, bandedSWA.cpp:3239 vs the scalar :226)
 *   rules=0 selects these vector rules (== getScores16), rules=1 the scalar kernel's (:132-253).
 * Pinned against the compiled reference itself (oracle/_ref, built by oracle/Makefile) by
 * tests/test_oracle.py and against the committed fixtures in tests/golden/.
 *
 * Arithmetic is int32 here; the reference's vector kernel is wrapping int16, so the two agree on the
 * reference's valid domain h0 + len2*match < 32768, len1,len2 < 32768 (SURVEY.md 8a note 4).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t idr, idq, id;
    int32_t len1, len2, h0, seqid, regid;
    int32_t score, tle, gtle, qle, gscore, max_off;
} oracle_seqpair; /* == SeqPair, bandedSWA.h:104-113 (72 bytes with tail padding) */

typedef struct {
    int32_t o_del, e_del, o_ins, e_ins, zdrop, end_bonus, match, mismatch, ambig;
} oracle_params;

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

static inline int sub_score(const oracle_params *P, int t, int q, int rules) {
    if (t >= 4 || q >= 4) return rules == 0 ? -1 : P->ambig; /* N vs anything, including N vs N */
    return t == q ? P->match : -P->mismatch;
}

/* band per pair: the vector wrapper's rule (bandedSWA.cpp:2898-2919; uint16 arithmetic), or, rules == 1, the scalar
 * kernel's (bandedSWA.cpp:164-172; int arithmetic through a double) */
static int pair_band(const oracle_params *P, int qlen, int w, int rules) {
    int mx = 0;
    if (mx < P->match) mx = P->match;
    if (mx < -P->mismatch) mx = -P->mismatch;
    if (mx < P->ambig) mx = P->ambig;
    if (rules == 1) {
        int max_ins = (int)((double)(qlen * mx + P->end_bonus - P->o_ins) / P->e_ins + 1.);
        int max_del = (int)((double)(qlen * mx + P->end_bonus - P->o_del) / P->e_del + 1.);
        return imin(imin(w, imax(max_ins, 1)), imax(max_del, 1));
    }
    uint16_t q = (uint16_t)(qlen * mx);
    uint16_t a = (uint16_t)(q + (uint16_t)(int16_t)(P->end_bonus - P->o_ins));
    int band = imin(w, imax((int)(a / P->e_ins + 1.0), 1));
    uint16_t b = (uint16_t)(q + (uint16_t)(int16_t)(P->end_bonus - P->o_del));
    band = imin(band, imax((int)(b / P->e_del + 1.0), 1));
    return band;
}

/* One pair. Hd[j] holds H(i-1, j-1) ("diagonal" for column j), Ev[j] holds E(i, j). */
static void extend_one(const oracle_params *P, const uint8_t *tgt, int tlen, const uint8_t *qry,
                       int qlen, int h0, int w, int zdrop_rule, int32_t *Hd, int32_t *Ev,
                       oracle_seqpair *out, int64_t *cells) {
    const int oe_del = P->o_del + P->e_del, oe_ins = P->o_ins + P->e_ins;
    int64_t ncell = 0;
    memset(Hd, 0, sizeof(int32_t) * (size_t)(qlen + 2));
    memset(Ev, 0, sizeof(int32_t) * (size_t)(qlen + 2));

    /* row "-1": H(-1,-1)=h0, then one gap open, then extensions while positive (:159-161) */
    Hd[0] = h0;
    if (qlen >= 1) Hd[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (int j = 2; j <= qlen && Hd[j - 1] > P->e_ins; ++j) Hd[j] = Hd[j - 1] - P->e_ins;

    const int band = pair_band(P, qlen, w, zdrop_rule);
    /* rows the vector kernel grants this lane (:3035-3036, :3130-3144) */
    const int row_budget = imin(qlen + band, tlen);

    int best = h0, best_i = -1, best_j = -1, g_i = -1, g = -1, off = 0;
    int beg = 0, end = qlen;
    for (int i = 0; i < tlen; ++i) {
        if (zdrop_rule == 0 && i + 1 > row_budget) break;
        if (beg < i - band) beg = i - band;
        if (end > i + band + 1) end = i + band + 1;
        if (end > qlen) end = qlen;
        if (zdrop_rule == 0 && beg >= end) break;       /* vector: tail <= head retires the lane */
        int hleft = 0;                                  /* H(i, beg-1) */
        if (beg == 0) hleft = imax(h0 - (P->o_del + P->e_del * (i + 1)), 0);
        int f = 0, rowmax = 0, rowarg = -1, j;
        for (j = beg; j < end; ++j) {
            int d = Hd[j], e = Ev[j];
            Hd[j] = hleft;
            int M = d ? d + sub_score(P, tgt[i], qry[j], zdrop_rule) : 0;
            int h = imax(imax(M, e), f);
            hleft = h;
            if (h >= rowmax) { rowmax = h; rowarg = j; } /* LAST column reaching the row max */
            int t = imax(M - oe_del, 0);
            Ev[j] = imax(e - P->e_del, t);
            t = imax(M - oe_ins, 0);
            f = imax(f - P->e_ins, t);
            ++ncell;
        }
        Hd[end] = hleft; Ev[end] = 0;
        if (j == qlen) {                                 /* row reached the query end (:218-221) */
            if (!(g > hleft)) g_i = i;
            g = imax(g, hleft);
        }
        if (rowmax == 0) break;
        if (rowmax > best) {
            best = rowmax; best_i = i; best_j = rowarg;
            off = imax(off, abs(rowarg - i));
        } else if (P->zdrop > 0 || zdrop_rule == 0) {
            int di = i - best_i, dj = rowarg - best_j;
            int pen;
            if (zdrop_rule == 0) pen = di > dj ? di - dj : dj - di;              /* vector: no factor */
            else pen = di > dj ? (di - dj) * P->e_del : (dj - di) * P->e_ins;   /* scalar */
            if (best - rowmax - pen > P->zdrop) break;
        }
        /* shrink [beg,end) past all-zero (H,E) cells (:233-237) */
        for (j = beg; j < end && Hd[j] == 0 && Ev[j] == 0; ++j) {}
        beg = j;
        for (j = end; j >= beg && Hd[j] == 0 && Ev[j] == 0; --j) {}
        end = imin(j + 2, qlen);
    }
    out->score = best; out->qle = best_j + 1; out->tle = best_i + 1;
    out->gtle = g_i + 1; out->gscore = g; out->max_off = off;
    if (cells) *cells += ncell;
}

/* Fills score/qle/tle/gtle/gscore/max_off of pairs[0..n). Sequences: ref+idr (len1, target),
 * qer+idq (len2, query), codes 0..4. Returns 0, or 1 on allocation failure.
 * cells_visited (nullable): total inner-loop iterations == the commented SW_cells++ (:215). */
int bsw_oracle_batch(const oracle_params *P, oracle_seqpair *pairs, const uint8_t *ref,
                     const uint8_t *qer, int64_t n, int32_t w, int32_t nthreads,
                     int64_t *cells_visited, int32_t zdrop_rule) {
    int64_t total = 0;
    int fail = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(+ : total) reduction(| : fail)
    {
        int cap = 0;
        int32_t *Hd = NULL, *Ev = NULL;
#pragma omp for schedule(dynamic, 256)
        for (int64_t k = 0; k < n; ++k) {
            oracle_seqpair *p = pairs + k;
            if (p->len2 + 2 > cap) {
                cap = p->len2 + 2 + 256;
                free(Hd); free(Ev);
                Hd = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
                Ev = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
                if (!Hd || !Ev) { fail = 1; cap = 0; continue; }
            }
            int64_t c = 0;
            extend_one(P, ref + p->idr, p->len1, qer + p->idq, p->len2, p->h0, w, zdrop_rule, Hd, Ev,
                       p, &c);
            total += c;
        }
        free(Hd); free(Ev);
    }
    if (cells_visited) *cells_visited = total;
    return fail;
}

int bsw_oracle_sizeof_seqpair(void) { return (int)sizeof(oracle_seqpair); }
