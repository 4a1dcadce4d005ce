/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
 *
 * Plain-C restatement of bwa-mem2's batched mate-rescue Smith-Waterman (class kswv), one pair at a time:
 * /root/reference/benchmarks/fmi/bwa-mem2/x86_64/src/kswv.cpp, the AVX512BW bodies (the only ones the class
 * has): kswv512_u8 (:371-716) behind kswvBatchWrapper8 (:177-369) and kswv512_16 (:933-1215) behind
 * kswvBatchWrapper16 (:733-930), driven as the production caller drives them (sort_classify,
 * bwamem.cpp:1136-1163; mem_sam_pe_batch, bwamem_pair.cpp:634-704).
 *
 * Pinned: tests/test_kswv_oracle.py compares this file with the compiled, unmodified reference
 * (oracle/_ref/libkswv_ref_avx512.so, built by oracle/Makefile through kswv_ref_shim.cpp) on seeded batches and
 * holds golden vectors generated from that library (tests/golden/kswv_*.npz, scripts/make_kswv_golden.py).
 *
 * What one lane of the vector code computes (the restatement follows the lane, not the 64/32-lane batch;
 * each statement below is why the rest of the batch cannot change a lane's result):
 *  - the query is padded to a multiple of 16 (8-bit class, kswv.cpp:295-306) or 8 (16-bit class, :857-869)
 *    columns with a code that scores 0 against everything (DUMMY5 via five512 :68-69, DUMMY3 via the zero
 *    entries of perm512 :957-971); those columns are part of the lane's DP and can carry a row maximum;
 *  - columns past that quantum and rows past len1 hold 0xFF: the match term is forced to 0 there
 *    (:72-75, :93-95). Only E reaches the columns on the right and nothing flows back; on the rows below,
 *    every value is smaller than the one above it, so they never raise gmax, never count as a rising row
 *    for the row-maximum filter, and the second-best loops mask them (rlen, :653-672, :1171-1188);
 *  - no traceback; H, E, F as in ksw_u8/ksw_i16 with E reset to 0 at every row start and F carried per column;
 *    8-bit class in saturating unsigned arithmetic with the bias `shift`, 16-bit class in wrapping int16;
 *  - the row maximum imax and the FIRST column that reaches it (strict >, :76-78);
 *  - Block I (:510-523, :1063-1077): row i-1's maximum is kept in rowMax[] only if row i did not rise above it,
 *    the previous comparison did not keep a row, the row reached minsc (KSW_XSUBO) and the lane is still live;
 *  - Block II (:526-548, :1080-1097): gmax/te/qe on a strictly larger row maximum while live; the lane stops
 *    being live once gmax >= endsc (KSW_XSTOP) or, 8-bit class, gmax + shift saturates;
 *  - second best (:589-703, :1139-1212): the largest kept row maximum (first on ties) outside
 *    [te - val, te + val], val = ceil(score / max(a, b, ambig)); 8-bit class reports 0 as -1 and a saturated
 *    lane (score 255) as score2 = te2 = -1 (the reference does not write those two fields when all 64 lanes
 *    of a batch saturated, :586; the restatement and the shim's pre-set value agree on -1);
 *  - phase 1 (KSW_XSTART; bwamem_pair.cpp:660-699): the aligned prefixes ref[0..te], qer[0..qe] reversed in
 *    place, len2 = qe + 1, len1 UNCHANGED, h0 = KSW_XSTOP | score, same class; tb/qb only if the score matches.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KSW_XBYTE  0x10000
#define KSW_XSTOP  0x20000
#define KSW_XSUBO  0x40000
#define KSW_XSTART 0x80000
#define KSWV_AMBIG (-1)          /* DEFAULT_AMBIG, bandedSWA.h / kswv.cpp:131 */

typedef struct { int64_t idr, idq, id; int32_t len1, len2, h0, seqid, regid, score, tle, gtle, qle, gscore, max_off; } seqpair_t;
typedef struct { int32_t score, te, qe, score2, te2, tb, qb; } kswr_t;
typedef struct { int o_del, e_del, o_ins, e_ins, a, b; } kswv_par;   /* b = mismatch penalty, positive */

static inline int imax(int x, int y) { return x > y ? x : y; }
static inline int imin(int x, int y) { return x < y ? x : y; }

typedef struct { int score_raw, te, qe; } lane_best;

/* One lane of kswv512_u8 (byte=1) / kswv512_16 (byte=0) over t[0..tlen) x q[0..qlen). rowmax[0..tlen) gets the
 * lane's column of rowMax[]. Returns gmax (raw), te, qe. */
static lane_best kswv_lane(const kswv_par *P, const uint8_t *t, int tlen, const uint8_t *q, int qlen, int xtra,
                           int byte, int *rowmax, int *H0, int *H1, int *F) {
    const int quantum = byte ? 16 : 8;
    const int ncol = (qlen + quantum - 1) / quantum * quantum;
    const int a = P->a, b = -P->b, amb = KSWV_AMBIG;
    const int minv = imin(imin(a, b), amb);
    const int shift = byte ? (uint8_t)(256 - (uint8_t)minv) : 0;               /* kswv.cpp:396-404 */
    const int none = byte ? 0 : -1;
    const int lim = byte ? 255 : 32767;
    int v = (xtra & KSW_XSUBO) ? (xtra & 0xffff) : 0x10000;                    /* :422-437, :976-993 */
    const int has_minsc = v <= lim, minsc = v;
    v = (xtra & KSW_XSTOP) ? (xtra & 0xffff) : 0x10000;
    const int has_endsc = v <= lim, endsc = v;
    const int oe_del = P->o_del + P->e_del, oe_ins = P->o_ins + P->e_ins;

    for (int j = 0; j <= ncol; ++j) H0[j] = H1[j] = F[j] = 0;
    int gmax = 0, te = -1, qe = 0, alive = 1;
    int pimax = 0, mask = 0, minsc_ok = 0;
    int i;
    for (i = 0; i < tlen; ++i) {
        const int s1 = t[i];
        int e = 0, rmax = 0, iqe = byte ? 255 : -1;
        for (int j = 0; j < ncol; ++j) {
            int sc;
            if (j >= qlen) sc = 0;                                             /* the zero-score dummy columns */
            else if (s1 == 4 || q[j] == 4) sc = amb;
            else sc = (s1 == q[j]) ? a : b;
            const int h00 = H0[j], f = F[j + 1];
            int m, h;
            if (byte) {
                m = imin(255, h00 + (uint8_t)(sc + shift));                    /* adds_epu8 */
                m = imax(0, m - shift);                                        /* subs_epu8 */
                h = imax(imax(m, e), f);
            } else {
                m = (int16_t)(h00 + sc);
                h = imax(imax(imax(m, e), f), 0);
            }
            if (h > rmax) { rmax = h; iqe = byte ? (j & 255) : (int16_t)j; }
            if (byte) {
                e = imax(imax(0, h - oe_ins), imax(0, e - P->e_ins));
                F[j + 1] = imax(imax(0, h - oe_del), imax(0, f - P->e_del));
            } else {
                e = imax((int16_t)(h - oe_ins), (int16_t)(e - P->e_ins));
                F[j + 1] = imax((int16_t)(h - oe_del), (int16_t)(f - P->e_del));
            }
            H1[j + 1] = h;
        }
        if (i > 0) {                                                           /* Block I */
            const int msk = (rmax > pimax) | mask;
            rowmax[i - 1] = (!msk && minsc_ok && alive) ? pimax : none;
            mask = !msk;
        }
        pimax = rmax;
        minsc_ok = has_minsc && rmax >= minsc;
        if (alive && rmax > gmax) { gmax = rmax; te = i; qe = iqe; }           /* Block II */
        int stop = has_endsc && gmax >= endsc;
        if (byte && imin(255, gmax + shift) >= 255) stop = 1;
        if (stop) alive = 0;
        { int *S = H1; H1 = H0; H0 = S; }
        if (!alive) {
            /* every later row of this lane is stored as `none` (exit0 clear); nothing else depends on them */
            for (int k = i; k < tlen; ++k) rowmax[k] = none;
            lane_best r = { gmax, te, qe };
            return r;
        }
    }
    if (tlen > 0)                                                              /* the store after the loop */
        rowmax[tlen - 1] = (!mask && minsc_ok && alive) ? pimax : none;
    lane_best r = { gmax, te, qe };
    return r;
}

/* second best, kswv.cpp:589-703 / :1139-1212 */
static void kswv_second(const kswv_par *P, const int *rowmax, int tlen, int score_raw, int te, int byte,
                        int *score2, int *te2) {
    const int qmax = imax(imax(P->a, -P->b), KSWV_AMBIG);
    const int val = (score_raw + qmax - 1) / qmax;
    const int low = (int16_t)(te - val), high = (int16_t)(te + val);
    int mx = byte ? 0 : -1, at = -1;
    for (int i = 0; i < low && i < tlen; ++i)
        if (rowmax[i] > mx) { mx = rowmax[i]; at = i; }
    for (int i = imax(high + 1, 0); i < tlen; ++i)
        if (rowmax[i] > mx) { mx = rowmax[i]; at = i; }
    *score2 = byte ? (mx == 0 ? -1 : mx) : mx;
    *te2 = at;
}

static void reverse_copy(uint8_t *dst, const uint8_t *src, int n) {
    for (int i = 0; i < n; ++i) dst[i] = src[n - 1 - i];
}

int kswv_oracle_sizeof_seqpair(void) { return (int)sizeof(seqpair_t); }

/* params: {o_del, e_del, o_ins, e_ins, match, mismatch(+ve)}. aln[pairs[i].regid] is written for every pair.
 * Returns 0, or -1 when out of memory. cells (may be NULL) += DP cells visited (rows x padded columns). */
int kswv_oracle_batch(const int32_t *params, const seqpair_t *pairs, const uint8_t *ref, const uint8_t *qer,
                      int64_t n, kswr_t *aln, int nthreads, int64_t *cells) {
    const kswv_par P = { params[0], params[1], params[2], params[3], params[4], params[5] };
    int64_t total = 0;
    int fail = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : total) reduction(| : fail)
    for (int64_t k = 0; k < n; ++k) {
        const seqpair_t *sp = &pairs[k];
        const int xtra = sp->h0, byte = (xtra & KSW_XBYTE) != 0;
        const int tlen = sp->len1, qlen = sp->len2;
        kswr_t *r = &aln[sp->regid];
        int *rowmax = (int *)malloc(sizeof(int) * (size_t)(tlen + 1));
        int *H0 = (int *)malloc(sizeof(int) * (size_t)(qlen + 40) * 3), *H1 = H0 + qlen + 40, *F = H1 + qlen + 40;
        uint8_t *tr = (uint8_t *)malloc((size_t)tlen + 1), *qr = (uint8_t *)malloc((size_t)qlen + 1);
        if (!rowmax || !H0 || !tr || !qr) { fail = 1; free(rowmax); free(H0); free(tr); free(qr); continue; }
        const uint8_t *t = ref + sp->idr, *q = qer + sp->idq;
        r->tb = r->qb = -1;                                                    /* bwamem_pair.cpp:634-637 */
        lane_best best = kswv_lane(&P, t, tlen, q, qlen, xtra, byte, rowmax, H0, H1, F);
        const int quantum = byte ? 16 : 8;
        total += (int64_t)tlen * ((qlen + quantum - 1) / quantum * quantum);
        if (byte) {
            const int shift = (uint8_t)(256 - (uint8_t)imin(imin(P.a, -P.b), KSWV_AMBIG));
            r->score = best.score_raw + shift < 255 ? best.score_raw : 255;    /* :568 */
        } else r->score = best.score_raw;
        r->te = best.te; r->qe = best.qe;
        if (byte && r->score == 255) r->score2 = r->te2 = -1;
        else kswv_second(&P, rowmax, tlen, best.score_raw, best.te, byte, &r->score2, &r->te2);
        /* phase 1, bwamem_pair.cpp:660-699 */
        if ((xtra & KSW_XSTART) && !((xtra & KSW_XSUBO) && r->score < (xtra & 0xffff))) {
            const int qlen1 = r->qe + 1, rt = r->te + 1;
            if (qlen1 <= qlen && rt <= tlen && qlen1 >= 0 && rt >= 0) {
                reverse_copy(qr, q, qlen1);
                reverse_copy(tr, t, rt);
                memcpy(tr + rt, t + rt, (size_t)(tlen - rt));
                lane_best rev = kswv_lane(&P, tr, tlen, qr, qlen1, KSW_XSTOP | r->score, byte, rowmax, H0, H1, F);
                if (r->score == rev.score_raw) { r->tb = r->te - rev.te; r->qb = r->qe - rev.qe; }   /* :562-565, :1116-1119 */
            } else fail = 1;   /* 8-bit class with 256 or more padded columns: qe wrapped; outside the domain */
        }
        free(rowmax); free(H0); free(tr); free(qr);
    }
    if (cells) *cells += total;
    return fail ? -1 : 0;
}
