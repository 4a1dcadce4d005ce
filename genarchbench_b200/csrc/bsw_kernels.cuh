// CUDA kernels (sm_100a) for the bsw hot path: banded Smith-Waterman seed extension with affine
// gaps, band, z-drop, end bonus -- bit-exact per pair with the reference's getScores16
// (/root/reference/benchmarks/bsw/src/bandedSWA.cpp:2679-3365; semantics defined by
// scalarBandedSWA :132-253 plus the vector path's z-drop rule :1889-1902, band :2898-2919 and
// row budget :3035-3036,3130-3144).
//
// Design:
//  * one THREAD per pair. The H/E band rows of the pair live in shared memory, interleaved by
//    thread (word w of thread t at [w*NT + t]) so every access is bank-conflict free whatever column
//    each thread is at. Long pairs use the same code over a global-memory scratch.
//  * the two 16-bit lanes of every DPX instruction are two ADJACENT COLUMNS (2g, 2g+1) of the same
//    row of the same pair, so all row-sequential decisions of the reference (band clamp, row max /
//    last argmax, m==0 exit, z-drop, trailing-zero trimming) stay exact and per pair.
//      M  = Hd + min(s, Hd)            (== Hd ? Hd+s : 0 up to values <= 0, which behave like 0)
//      T  = max(M - oe, 0)             VIADDMNMX.S16x2.RELU
//      E' = max(E - e_del, T)          VIADDMNMX.S16x2
//      F  : two-step in-register scan  2 x VIADDMNMX.S16x2 (+1 IMAD, +1 shift)
//      H  = max(M, E, F)               VIMNMX3.S16x2
//    substitution scores for both lanes come from ONE PRMT that indexes an 8-byte LUT with the
//    per-lane (query ^ target) code (ambiguous base folded in through an OR on bit 2).
//  * packed sequences (2-bit, or 4-bit when a pair holds an ambiguous base) sit in 16-byte aligned
//    slots in the caller's order; each thread expands its own slot into shared memory once per pair.
#pragma once
#include <stdint.h>
#ifdef BSW_HOST_EMUL
// tests/host_emul compiles the per-pair code below with g++ against an emulation of the few CUDA
// intrinsics it uses, so the algorithm can be checked against the oracle without a GPU.
#include "dpx_host_emul.h"
#else
#include <cuda_runtime.h>
#endif

namespace bswk {

constexpr int kBlockPairs = 128;  // threads per block == pairs per block (host packs blobs per block)

struct KParams {
    int o_del, e_del, o_ins, e_ins, zdrop, end_bonus, match, mismatch, ambig, w;
    // max(0, match, -mismatch, ambig), the reference wrapper's `max` (bandedSWA.cpp:2790-2793).
    // Computed on the HOST: ptxas 12.9 (sm_100a) fuses max(max(max(match, -mismatch), ambig), 0)
    // into one VIMNMX3.RELU and drops the negation (seen on B200: band came out as w).
    int max_score;
    // Multipliers read from the parameter bank at run time, so that ptxas keeps the lane shifts
    // below as IMAD / IMAD.HI on the FMA pipe instead of folding them into ALU-pipe shifts (the
    // ALU pipe is what bounds this kernel): k16 = 65536, km = match + 1.
    uint32_t k16, km;
};
__host__ __device__ inline int max_score_of(int match, int mismatch, int ambig) {
    int mx = 0;
    if (mx < match) mx = match;
    if (mx < -mismatch) mx = -mismatch;
    if (mx < ambig) mx = ambig;
    return mx;
}

// 16 bytes per pair, sorted order (host: length-binned). `off` in 4-byte units from the blob base.
struct __align__(16) PairMeta {
    uint32_t off;     // start of this pair's packed [query | target] blob
    uint32_t id;      // index of the pair in the caller's order (results are written to out[id])
    uint16_t len2;    // query length
    uint16_t len1;    // target length
    int16_t  h0;
    uint16_t flags;   // bit0: blob is 4-bit ("wide": the pair contains an ambiguous base)
};

struct __align__(16) PairOut {  // one STG.128 per pair
    int16_t score, qle, tle, gtle, gscore, max_off;
    uint32_t cells;   // COUNT kernels only: DP cells the reference's scalar loop visits for this pair
};

__host__ __device__ inline uint32_t seq_bytes(uint32_t len, bool wide) {
    uint32_t b = wide ? (len + 1) >> 1 : (len + 3) >> 2;
    return (b + 3u) & ~3u;  // each sequence padded to 4 bytes
}

// 4-byte words a pair occupies in the (2-bit) slot area: query then target, 16-byte aligned
__host__ __device__ inline uint32_t slot_words(uint32_t len2, uint32_t len1) {
    uint32_t b = seq_bytes(len2, false) + seq_bytes(len1, false);
    return ((b + 15u) & ~15u) >> 2;
}

__device__ __forceinline__ uint32_t pack2(int v) {
    uint32_t u = (uint32_t)v & 0xFFFFu;
    return u | (u << 16);
}

// ---------------------------------------------------------------------------------------------
// Row storage. Words are interleaved by thread: word w of thread t lives at base[w * stride + t].
//   he[g]  : .x = { Hs[2g], Hs[2g+1] }  with Hs[j] = H(i-1, j-1)  (the reference's eh[j].h)
//            .y = { E[2g],  E[2g+1]  }  with E[j]  = E(i, j)       (the reference's eh[j].e)
//   qs[g]  : 16-bit PRMT selector seed of query columns (2g, 2g+1)
//   tg[w]  : target bases, 4 bits each, 8 per word
// ---------------------------------------------------------------------------------------------
struct Rows {
    uint2 *he;
    uint16_t *qs;
    uint32_t *tg;
    int stride;  // threads sharing the arrays (blockDim for shared memory, grid-wide for global)
    __device__ __forceinline__ uint2 &HE(int g) const { return he[(size_t)g * stride]; }
    // 16-bit views of the rows. They go through the SAME 32-bit words the packed loop reads and
    // writes (no differently-typed aliases the compiler could reorder around the uint2 accesses).
    __device__ __forceinline__ uint32_t getH16(int j) const {
        const uint32_t w = he[(size_t)(j >> 1) * stride].x;
        return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
    }
    __device__ __forceinline__ uint32_t getE16(int j) const {
        const uint32_t w = he[(size_t)(j >> 1) * stride].y;
        return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
    }
    // Hs[j] = 0, E[j] = 0 with two 16-bit stores (no read-modify-write latency in front of the row).
    // Issued as asm with a memory clobber so the packed 64-bit accesses are not moved across them.
    __device__ __forceinline__ void zeroHE16(int j) const {
#ifdef BSW_HOST_EMUL
        setHE16(j, 0u, 0u);
#else
        unsigned char *p = reinterpret_cast<unsigned char *>(&he[(size_t)(j >> 1) * stride]) + 2 * (j & 1);
        asm volatile("st.u16 [%0], %2;\n\tst.u16 [%1], %2;" ::"l"(p), "l"(p + 4), "h"((unsigned short)0) : "memory");
#endif
    }
    // sets Hs[j] = hv and E[j] = ev, leaving the other half of the word pair untouched
    __device__ __forceinline__ void setHE16(int j, uint32_t hv, uint32_t ev) const {
        uint2 &p = he[(size_t)(j >> 1) * stride];
        uint2 v = p;
        if (j & 1) { v.x = (v.x & 0xFFFFu) | (hv << 16); v.y = (v.y & 0xFFFFu) | (ev << 16); }
        else { v.x = (v.x & 0xFFFF0000u) | hv; v.y = (v.y & 0xFFFF0000u) | ev; }
        p = v;
    }
    __device__ __forceinline__ uint16_t &QS(int g) const { return qs[(size_t)g * stride]; }
    __device__ __forceinline__ uint32_t &TG(int w) const { return tg[(size_t)w * stride]; }
};

// PTX prmt.b32 (default mode): byte i of the result = byte (nibble_i & 7) of {b:a}; nibble bit 3 set
// => that byte's SIGN replicated instead. (__byte_perm() masks the selector with 0x7777 and loses
// the sign mode, so the instruction is issued directly.)
__device__ __forceinline__ uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef BSW_HOST_EMUL
    return emul::prmt(a, b, sel);
#else
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#endif
}
// one LOP3: per bit, mask ? (q | t) : (q ^ t)
__device__ __forceinline__ uint32_t sel_combine(uint32_t q, uint32_t t, uint32_t mask) {
#ifdef BSW_HOST_EMUL
    return (mask & (q | t)) | (~mask & (q ^ t));
#else
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xBC;" : "=r"(d) : "r"(q), "r"(t), "r"(mask));
    return d;
#endif
}

// selector nibble pattern of one base code (0..4): both nibbles carry the code
__device__ __forceinline__ uint32_t base_pat(uint32_t code) { return code * 0x11u; }

// Expands this thread's packed blob (already visible at `blob`, 4-byte words) into qs[] / tg[].
__device__ inline void unpack_pair(const uint32_t *blob, int qlen, int tlen, bool wide, const Rows &R) {
    const int ngroups = (qlen + 1) >> 1;
    if (!wide) {
        // query: 16 bases per word -> 8 selector seeds
        for (int w = 0, g = 0; g < ngroups; ++w) {
            uint32_t x = blob[w];
#pragma unroll
            for (int k = 0; k < 8; ++k, ++g) {
                if (g < ngroups) {
                    uint32_t two = (x >> (4 * k)) & 0xFu;
                    R.QS(g) = (uint16_t)(base_pat(two & 3u) | (base_pat(two >> 2) << 8));
                }
            }
        }
        const uint32_t *tb = blob + (seq_bytes(qlen, false) >> 2);
        const int twords = (tlen + 7) >> 3;
        for (int w = 0; w < twords; ++w) {
            uint32_t x = tb[w >> 1];
            x = (w & 1) ? (x >> 16) : (x & 0xFFFFu);   // 8 bases, 2 bits each
            x = (x | (x << 8)) & 0x00FF00FFu;
            x = (x | (x << 4)) & 0x0F0F0F0Fu;
            x = (x | (x << 2)) & 0x33333333u;           // -> 8 nibbles
            R.TG(w) = x;
        }
    } else {
        for (int w = 0, g = 0; g < ngroups; ++w) {
            uint32_t x = blob[w];
#pragma unroll
            for (int k = 0; k < 4; ++k, ++g) {
                if (g < ngroups) {
                    uint32_t two = (x >> (8 * k)) & 0xFFu;
                    R.QS(g) = (uint16_t)(base_pat(two & 0xFu) | (base_pat(two >> 4) << 8));
                }
            }
        }
        const uint32_t *tb = blob + (seq_bytes(qlen, true) >> 2);
        const int twords = (tlen + 7) >> 3;
        for (int w = 0; w < twords; ++w) R.TG(w) = tb[w];
    }
}

struct PairResult {
    int score, qle, tle, gtle, gscore, max_off;
    uint32_t cells;
};

// per-pair band, the vector wrapper's rule (bandedSWA.cpp:2898-2919): uint16 arithmetic, integer
// division, then +1.
__device__ __forceinline__ int pair_band(const KParams &P, int qlen) {
    const int mx = P.max_score;
    uint32_t q = (uint32_t)(qlen * mx) & 0xFFFFu;
    uint32_t a = (q + (uint32_t)(P.end_bonus - P.o_ins)) & 0xFFFFu;
    int band = min(P.w, max((int)(a / (uint32_t)P.e_ins) + 1, 1));
    uint32_t b = (q + (uint32_t)(P.end_bonus - P.o_del)) & 0xFFFFu;
    band = min(band, max((int)(b / (uint32_t)P.e_del) + 1, 1));
    return band;
}

// The DP of one pair over row storage R (already holding qs[] and tg[]).
//   FASTM : every score of the launch satisfies score * (match + 1) <= 32767, so the reference's
//           M = Hd ? Hd + s : 0 is ONE instruction, min(Hd + s, Hd * (match + 1)) (the product on the
//           FMA pipe): for Hd >= 1 the second term is >= Hd + match >= Hd + s, for Hd == 0 it caps M
//           at 0 -- and any M <= 0 behaves like 0 in max(M, E, F) and in max(M - oe, 0).
//   SYM   : o_del == o_ins && e_del == e_ins (one T for both gap kinds)
//   COUNT : also track the reference's exact leading trim and count the cells its scalar loop would
//           visit (bandedSWA.cpp:191-216; the commented SW_cells++ at :215) -- the unit of work of
//           the GCUPS metric. Used once per input outside any timed region.
template <bool FASTM, bool SYM, bool COUNT>
__device__ inline PairResult extend_pair(const Rows &R, int qlen, int tlen, int h0, const KParams &P) {
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const uint32_t NEG_OE_DEL = pack2(-oe_del), NEG_OE_INS = pack2(-oe_ins);
    const uint32_t NEG_E_DEL = pack2(-P.e_del), NEG_E_INS = pack2(-P.e_ins);
    // PRMT look-up table: index 0 match, 1..3 mismatch, 4..7 ambiguous (low byte of the score; PRMT
    // replicates its sign into the high byte of each lane)
    const uint32_t LUT_LO = ((uint32_t)P.match & 0xFFu) | (((uint32_t)(-P.mismatch) & 0xFFu) * 0x01010100u);
    const uint32_t LUT_HI = ((uint32_t)P.ambig & 0xFFu) * 0x01010101u;

    // row "-1" (bandedSWA.cpp:159-161) and zeroed E; one spare column so the hi lane of the last
    // word is always initialised
    {
        const int nwords = ((qlen + 1) >> 1) + 1;
        int hv = h0;
        for (int g = 0; g < nwords; ++g) {
            int a = hv;                                   // Hs[2g]
            if (g == 0) hv = h0 > oe_ins ? h0 - oe_ins : 0;
            else hv = max(hv - P.e_ins, 0);
            int b = hv;                                   // Hs[2g+1]
            hv = max(hv - P.e_ins, 0);
            if (2 * g > qlen) a = 0;                      // the reference's calloc'ed tail
            if (2 * g + 1 > qlen) b = 0;
            R.HE(g) = make_uint2((uint32_t)a | ((uint32_t)b << 16), 0u);
        }
    }

    const int band = pair_band(P, qlen);
    const int budget = min(qlen + band, tlen);
    const uint32_t K16 = P.k16, KM = P.km;

    int best = h0, best_i = -1, best_j = -1, g_i = -1, gsc = -1, off = 0;
    int beg = 0, end = qlen;
    uint32_t tword = 0;
    int hcol = h0 - P.o_del;  // first column: H(i,-1) = max(h0 - o_del - e_del*(i+1), 0)
    int xbeg = 0;             // COUNT: the reference's exact beg (ours lags it by whole words)
    uint32_t cells = 0;

    for (int i = 0; i < budget; ++i) {
        if (beg < i - band) beg = i - band;
        if (end > i + band + 1) end = i + band + 1;
        if (beg >= end) break;
        if (COUNT) {
            if (xbeg < i - band) xbeg = i - band;
            cells += (uint32_t)(end - xbeg);
        }

        if ((i & 7) == 0) tword = R.TG(i >> 3);
        const uint32_t tsel = (tword & 7u) * 0x1111u + 0x8080u;   // (code * 0x11 | 0x80) in both bytes
        tword >>= 4;

        hcol -= P.e_del;
        const int hleft = beg == 0 ? max(hcol, 0) : 0;

        // Lanes outside [beg, end) of the first / last word must see zero inputs: clear the stale
        // (never read again) entries instead of masking inside the loop.
        if (beg & 1) R.zeroHE16(beg - 1);
        if (end & 1) R.zeroHE16(end);

        const int g0 = beg >> 1, g1 = (end - 1) >> 1;
        uint32_t hprev = (uint32_t)hleft << 16;  // .hi = H(i, 2*g0 - 1)
        uint32_t A = 0;                          // { F(i, 2g), 0 }
        uint32_t rm = 0;                         // running max per lane (even / odd columns)
        int mjlo = -1, mjhi = -1;
        uint32_t h = 0, En = 0, Hst = 0;

        // One group = columns (2g, 2g+1). Only the F scan is serial along the row; the scores, M, T
        // and E' of different groups are independent. With 2-3 resident warps per scheduler the
        // kernel is bound by dependency latency, so groups are processed four at a time: all loads
        // first, then the independent parts of the four groups (interleavable), then the F chain.
        auto front = [&](const uint2 he, const uint32_t qsel, uint32_t &M, uint32_t &Tins, uint32_t &Enew) {
            const uint32_t Hd = he.x, Ev = he.y;
            // k = q ^ t on bits 0-1 (and the sign-replicate bit), q | t on bit 2 (ambiguous)
            const uint32_t sel = sel_combine(qsel, tsel, 0x4444u);
            const uint32_t s = prmt_sx(LUT_LO, LUT_HI, sel);
            if (FASTM) {
                M = __viaddmin_s16x2(Hd, s, Hd * KM);
            } else {
                const uint32_t sm = __vmins2(s, __vmins2(Hd, 0x00010001u) * (uint32_t)P.match);
                M = __vadd2(Hd, sm);
            }
            const uint32_t Tdel = __viaddmax_s16x2_relu(M, NEG_OE_DEL, NEG_OE_DEL);
            Tins = SYM ? Tdel : __viaddmax_s16x2_relu(M, NEG_OE_INS, NEG_OE_INS);
            Enew = __viaddmax_s16x2(Ev, NEG_E_DEL, Tdel);
        };
        auto back = [&](const int g, const uint32_t Ev, const uint32_t M, const uint32_t Tins, const uint32_t Enew) {
            const uint32_t W1 = __viaddmax_s16x2(A, NEG_E_INS, Tins);   // .lo = F(i, 2g+1)
            const uint32_t B = W1 * K16 + A;                             // { F(2g), F(2g+1) }  (IMAD)
            h = __vimax3_s16x2(M, Ev, B);
            const uint32_t W2 = __viaddmax_s16x2(B, NEG_E_INS, Tins);   // .hi = F(i, 2g+2)
            A = __umulhi(W2, K16);                                       // W2 >> 16           (IMAD.HI)
            Hst = __umulhi(hprev, K16) + h * K16;                        // { H(i,2g-1), H(i,2g) }
            En = Enew;
            R.HE(g) = make_uint2(Hst, En);
            hprev = h;
            bool phi, plo;
            rm = __vibmax_s16x2(h, rm, &phi, &plo);
            if (plo) mjlo = g;
            if (phi) mjhi = g;
        };
        int g = g0;
        for (; g + 3 <= g1; g += 4) {
            const uint2 he0 = R.HE(g), he1 = R.HE(g + 1), he2 = R.HE(g + 2), he3 = R.HE(g + 3);
            const uint32_t q0 = R.QS(g), q1 = R.QS(g + 1), q2 = R.QS(g + 2), q3 = R.QS(g + 3);
            uint32_t M0, M1, M2, M3, T0, T1, T2, T3, E0, E1, E2, E3;
            front(he0, q0, M0, T0, E0);
            front(he1, q1, M1, T1, E1);
            front(he2, q2, M2, T2, E2);
            front(he3, q3, M3, T3, E3);
            back(g, he0.y, M0, T0, E0);
            back(g + 1, he1.y, M1, T1, E1);
            back(g + 2, he2.y, M2, T2, E2);
            back(g + 3, he3.y, M3, T3, E3);
        }
        for (; g <= g1; ++g) {
            const uint2 he0 = R.HE(g);
            const uint32_t q0 = R.QS(g);
            uint32_t M0, T0, E0;
            front(he0, q0, M0, T0, E0);
            back(g, he0.y, M0, T0, E0);
        }

        // last computed column's H, and the reference's eh[end] = { h1, 0 }
        int hlast;
        if (end & 1) {
            hlast = (int)(h & 0xFFFFu);          // word g1 already holds { .., H(i,end-1) } / E[end]=0
        } else {
            hlast = (int)(h >> 16);
            R.setHE16(end, (uint32_t)hlast, 0u);
        }
        if (end == qlen) {                        // bandedSWA.cpp:218-221
            if (!(gsc > hlast)) g_i = i;
            gsc = max(gsc, hlast);
        }
        const int mlo = (int)(short)(rm & 0xFFFFu), mhi = (int)(short)(rm >> 16);
        const int m = max(mlo, mhi);
        if (m == 0) break;
        int mj;
        {
            const int jlo = 2 * mjlo, jhi = 2 * mjhi + 1;
            mj = mlo > mhi ? jlo : (mhi > mlo ? jhi : max(jlo, jhi));   // LAST column reaching m
        }
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            off = max(off, abs(mj - i));
        } else {
            // vector z-drop rule: no gap-extend factor, no zdrop > 0 guard (bandedSWA.cpp:1889-1902)
            const int di = i - best_i, dj = mj - best_j;
            if (best - m - abs(di - dj) > P.zdrop) break;
        }

        if (COUNT) {   // the reference's scan (bandedSWA.cpp:234-235), on the rows just written
            int j = xbeg;
            while (j < end && R.getH16(j) == 0 && R.getE16(j) == 0) ++j;
            xbeg = j;
        }
        // leading trim (not semantic: skipped cells are all-zero; done lazily, one word every 4th row)
        if ((i & 3) == 3) {
            const uint2 z = R.HE(g0);
            if ((z.x | z.y) == 0u && 2 * (g0 + 1) > beg) beg = 2 * (g0 + 1);
        }
        // trailing trim (semantic): j* = last j in [beg,end] with Hs[j] | E[j] != 0 (m > 0
        // guarantees one exists); the new end is min(j* + 2, qlen). Hs[end] = H(i,end-1) is almost
        // always non-zero, so that case is tested first.
        if (hlast) {
            end = min(end + 2, qlen);
        } else {
            int jstar;
            const uint32_t Wt = Hst | En;
            if (end & 1) jstar = (Wt & 0xFFFFu) ? end - 1 : -1;
            else jstar = (Wt >> 16) ? end - 1 : ((Wt & 0xFFFFu) ? end - 2 : -1);
            if (jstar < 0) {
                int g = g1 - 1;
                uint32_t wz = 0;
                for (; g >= 0; --g) {
                    const uint2 z = R.HE(g);
                    wz = z.x | z.y;
                    if (wz) break;
                }
                jstar = (wz >> 16) ? 2 * g + 1 : 2 * g;
            }
            end = min(jstar + 2, qlen);
        }
    }

    PairResult r;
    r.score = best; r.qle = best_j + 1; r.tle = best_i + 1;
    r.gtle = g_i + 1; r.gscore = gsc; r.max_off = off;
    r.cells = cells;
    return r;
}

#ifndef BSW_HOST_EMUL
__device__ __forceinline__ void store_result(PairOut *out, uint32_t id, const PairResult &r) {
    union { PairOut o; uint4 v; } u;
    u.o.score = (int16_t)r.score; u.o.qle = (int16_t)r.qle; u.o.tle = (int16_t)r.tle;
    u.o.gtle = (int16_t)r.gtle; u.o.gscore = (int16_t)r.gscore; u.o.max_off = (int16_t)r.max_off;
    u.o.cells = r.cells;
    reinterpret_cast<uint4 *>(out)[id] = u.v;
}

// ---------------------------------------------------------------------------------------------
// Short pairs: rows in shared memory. Launch: grid = ceil(n / kBlockPairs), block = kBlockPairs,
// dynamic smem = (8*row_words + 2*qs_words + 4*tg_words) * NT.
//   meta[k] for k in [0, n): this launch's pairs in sorted (length-binned) order; blob = the slab's
//   packed sequences. Each thread expands its own 16-byte aligned slot straight from global memory
//   (a few dozen bytes per pair, read once; the host packs slots in the caller's order so that its
//   own pass is a pure stream -- see DESIGN.md).
// ---------------------------------------------------------------------------------------------
template <bool FASTM, bool SYM, bool COUNT>
__global__ void __launch_bounds__(kBlockPairs)
bsw_short_kernel(const PairMeta *__restrict__ meta, const uint32_t *__restrict__ blob,
                 PairOut *__restrict__ out, int n, KParams P, int row_words, int qs_words,
                 int tg_words) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = kBlockPairs;
    const int tid = threadIdx.x;
    const int k = blockIdx.x * NT + tid;
    if (k >= n) return;
    const PairMeta m = meta[k];

    Rows R;
    R.stride = NT;
    R.he = reinterpret_cast<uint2 *>(smem) + tid;
    R.qs = reinterpret_cast<uint16_t *>(smem + (size_t)8 * row_words * NT) + tid;
    R.tg = reinterpret_cast<uint32_t *>(smem + (size_t)8 * row_words * NT + (size_t)2 * qs_words * NT) + tid;
    (void)tg_words;

    // a wide pair's slot only holds the word offset of its 4-bit blob in the overflow area
    const uint32_t *src = blob + m.off;
    if (m.flags & 1) src = blob + src[0];
    unpack_pair(src, m.len2, m.len1, m.flags & 1, R);

    PairResult r = extend_pair<FASTM, SYM, COUNT>(R, m.len2, m.len1, m.h0, P);
    store_result(out, m.id, r);
}

// ---------------------------------------------------------------------------------------------
// Long pairs (rows do not fit the shared-memory bins): same per-pair code over a global scratch,
// interleaved by thread across the whole grid so neighbouring threads touch neighbouring words.
//   scratch layout: he[row_words][nthreads] (uint2) | qs[qs_words][nthreads] (u16, padded to 4 B)
//                   | tg[tg_words][nthreads] (u32)
// ---------------------------------------------------------------------------------------------
template <bool FASTM, bool SYM, bool COUNT>
__global__ void __launch_bounds__(kBlockPairs)
bsw_long_kernel(const PairMeta *__restrict__ meta, const uint32_t *__restrict__ blob,
                PairOut *__restrict__ out, int n, KParams P, int row_words, int qs_words,
                int tg_words, unsigned char *__restrict__ scratch) {
    const int nthreads = gridDim.x * blockDim.x;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const PairMeta m = meta[k];
    Rows R;
    R.stride = nthreads;
    unsigned char *p = scratch;
    R.he = reinterpret_cast<uint2 *>(p) + k;
    p += (size_t)8 * row_words * nthreads;
    R.qs = reinterpret_cast<uint16_t *>(p) + k;
    p += (((size_t)2 * qs_words * nthreads) + 15) & ~(size_t)15;
    R.tg = reinterpret_cast<uint32_t *>(p) + k;
    unpack_pair((m.flags & 1) ? blob + blob[m.off] : blob + m.off, m.len2, m.len1, m.flags & 1, R);
    PairResult r = extend_pair<FASTM, SYM, COUNT>(R, m.len2, m.len1, m.h0, P);
    store_result(out, m.id, r);
}

// ---------------------------------------------------------------------------------------------
// Integer-pipe microbenchmark: `iters` x 8 independent chains of one instruction kind per thread.
// ---------------------------------------------------------------------------------------------
template <int WHICH>
__global__ void dpx_peak_kernel(uint32_t *sink, int iters, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed * (k + 1) + threadIdx.x;
    const uint32_t c1 = seed | 0x00010001u, c2 = seed ^ 0x00070003u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (WHICH == 0) a[k] = __viaddmax_s16x2_relu(a[k], c1, c2);
                else if (WHICH == 1) a[k] = __vimax3_s16x2(a[k], c1, c2 + u);
                else if (WHICH == 2) a[k] = __vadd2(a[k], c1);
                else if (WHICH == 3) a[k] = (a[k] & c1) ^ (c2 + u);
                else if (WHICH == 4) a[k] = __byte_perm(a[k], c1, c2 + u);
                else if (WHICH == 5) a[k] = a[k] * c1 + c2;
                else if (WHICH == 6) a[k] = __funnelshift_r(a[k], c1, 16) + 0;
                else if (WHICH == 7) a[k] = __umulhi(a[k], 65536u) + c1;   // IMAD.HI
                else if (WHICH == 8) {                                      // ALU + FMA pipe mix
                    a[k] = __viaddmax_s16x2_relu(a[k], c1, c2);
                    a[(k + 4) & 7] = a[(k + 4) & 7] * 65536u + c2;
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= a[k];
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;  // keep the chains alive
}

#endif  // !BSW_HOST_EMUL

}  // namespace bswk
