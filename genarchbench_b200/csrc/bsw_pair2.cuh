// Thread-per-TWO-pairs DP ("duo"): the two 16-bit lanes of every DPX instruction are the SAME cell
// (i, j) of two DIFFERENT pairs A (low halves) and B (high halves) that sit next to each other in the
// length-sorted launch order. Included by bsw_kernels.cuh.
//
// Against extend_pair (lanes = two adjacent columns of one pair) this removes every cross-lane move:
// the F recurrence is one VIADDMNMX per column with no lane shuffling, and the shifted store of H (the
// diagonal of the next row) is simply the previous column's register. Per 2 cells the inner loop issues
// ~13 instructions instead of ~20. What it costs: the two pairs walk their rows together, so a column is
// computed for both as long as either needs it. Pairs that are neighbours in (len2, len1, h0) order
// have nearly the same [beg, end) per row -- on config 3, 97.6 % of the lane slots do useful work.
//
// Exactness per pair is kept as in extend_pair:
//   * a lane whose pair is not at this row / column any more (other pair longer, pair finished, column at
//     or right of its `end`) runs in MASKED mode: its stored entries and its row maximum are not touched;
//   * columns left of a lane's `beg` hold all-zero entries (zero-trimmed, or cleared when the band clamp
//     moved beg past them) and so compute to zero, exactly as if they were skipped.
#pragma once

namespace bswk {

// Row storage of one thread (words interleaved by thread, stride = threads of the block):
//   he4[k] : uint4 = columns 2k, 2k+1:  .x = { HsA[2k],   HsB[2k]   }   .y = { EA[2k],   EB[2k]   }
//                                       .z = { HsA[2k+1], HsB[2k+1] }   .w = { EA[2k+1], EB[2k+1] }
//   qs[k]  : u32   = selector seeds of columns 2k, 2k+1, 16 bits each: byte 0 pair A, byte 1 pair B
struct Rows2 {
    uint4 *he4;
    uint32_t *qs;
    int stride;
    __device__ __forceinline__ uint4 &HE4(int k) const { return he4[(size_t)k * stride]; }
    __device__ __forceinline__ uint2 &HE(int j) const {   // column j: .x = Hs halves, .y = E halves
        return reinterpret_cast<uint2 *>(he4 + (size_t)(j >> 1) * stride)[j & 1];
    }
    __device__ __forceinline__ uint32_t &QS2(int k) const { return qs[(size_t)k * stride]; }
    __device__ __forceinline__ uint32_t getH16(int j, int a) const {
        const uint32_t w = HE(j).x;
        return a ? (w >> 16) : (w & 0xFFFFu);
    }
    __device__ __forceinline__ uint32_t getE16(int j, int a) const {
        const uint32_t w = HE(j).y;
        return a ? (w >> 16) : (w & 0xFFFFu);
    }
    // Hs_a[j] = hv, E_a[j] = ev, the other pair's halves untouched
    __device__ __forceinline__ void setHE16(int j, int a, uint32_t hv, uint32_t ev) const {
#ifdef BSW_HOST_EMUL
        uint2 &p = HE(j);
        uint2 v = p;
        if (a) { v.x = (v.x & 0xFFFFu) | (hv << 16); v.y = (v.y & 0xFFFFu) | (ev << 16); }
        else { v.x = (v.x & 0xFFFF0000u) | hv; v.y = (v.y & 0xFFFF0000u) | ev; }
        p = v;
#else
        unsigned char *p = reinterpret_cast<unsigned char *>(&HE(j)) + 2 * a;
        asm volatile("st.u16 [%0], %2;\n\tst.u16 [%1], %3;" ::"l"(p), "l"(p + 4), "h"((unsigned short)hv),
                     "h"((unsigned short)ev) : "memory");
#endif
    }
};

// number of he4 / qs elements (2 columns each) for columns 0 .. qlen
// (a whole number of 4-column blocks: the leading trim looks at a block at a time)
__host__ __device__ inline int duo_elems(int qlen) { return ((qlen + 4) >> 2) << 1; }
__host__ __device__ inline uint32_t duo_thread_bytes(int qlen) { return 20u * (uint32_t)duo_elems(qlen); }

// One pair of a duo thread.
struct DuoLane {
    int qlen, tlen, h0;          // qlen == 0: no pair in this lane
    const uint32_t *tb;          // packed target
    bool wide_blob;              // this pair's blob is 4-bit
    // state
    int band, budget, beg, end, best, best_i, best_j, g_i, gsc, off, hcol, xbeg;
    uint32_t tword, cells;
    bool done;
};

// target bases of rows 8w .. 8w+7 of one lane as 8 nibbles, in the THREAD's selector convention
// (TWIDE: the code itself; otherwise 4 - code)
template <bool TWIDE>
__device__ __forceinline__ uint32_t duo_target(const DuoLane &L, int w) {
    uint32_t x;
    if (L.wide_blob) {
        x = L.tb[w];
    } else {
        x = L.tb[w >> 1];
        x = (w & 1) ? (x >> 16) : (x & 0xFFFFu);
        x = (x | (x << 8)) & 0x00FF00FFu;
        x = (x | (x << 4)) & 0x0F0F0F0Fu;
        x = (x | (x << 2)) & 0x33333333u;
    }
    return (TWIDE || BSW_SEL_LOP3) ? x : 0x44444444u - x;
}

// base j of a packed query (2 or 4 bits per base)
__device__ __forceinline__ uint32_t duo_query_base(const uint32_t *blob, bool wide_blob, int j) {
    return wide_blob ? (blob[j >> 3] >> (4 * (j & 7))) & 0xFu : (blob[j >> 4] >> (2 * (j & 15))) & 3u;
}

// Fills qs[] from the two packed queries and points the lanes at their targets.
__device__ inline void duo_unpack(const uint32_t *blobA, const uint32_t *blobB, DuoLane *L, const Rows2 &R) {
    const int qmax = max(L[0].qlen, L[1].qlen);
    const int nel = duo_elems(qmax);
    for (int k = 0; k < nel; ++k) {
        uint32_t w = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = 2 * k + u;
            const uint32_t qa = j < L[0].qlen ? duo_query_base(blobA, L[0].wide_blob, j) : 0u;
            const uint32_t qb = j < L[1].qlen ? duo_query_base(blobB, L[1].wide_blob, j) : 0u;
            w |= ((qa * 0x11u) | ((qb * 0x11u) << 8)) << (16 * u);
        }
        R.QS2(k) = w;
    }
    L[0].tb = blobA + (seq_bytes((uint32_t)L[0].qlen, L[0].wide_blob) >> 2);
    L[1].tb = L[1].qlen ? blobB + (seq_bytes((uint32_t)L[1].qlen, L[1].wide_blob) >> 2) : blobA;
}

__device__ __forceinline__ uint32_t half_of(uint32_t w, int a) { return a ? (w >> 16) : (w & 0xFFFFu); }

// The DP of the two pairs of a thread. TWIDE: at least one of them may hold an ambiguous base (LOP3
// selector; see score_lut). Other template flags as in extend_pair. Results in res[0], res[1].
template <bool FASTM, bool SYM, bool COUNT, bool TWIDE>
__device__ inline void extend_duo(const Rows2 &R, DuoLane *L, const KParams &P, PairResult *res) {
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const uint32_t NEG_OE_DEL = pack2(-oe_del), NEG_OE_INS = pack2(-oe_ins);
    const uint32_t NEG_E_DEL = pack2(-P.e_del), NEG_E_INS = pack2(-P.e_ins);
    uint32_t LUT_LO, LUT_HI;
    score_lut<TWIDE>(P, LUT_LO, LUT_HI);
    const uint32_t K16 = P.k16, KM = P.km, K1 = P.k1;
    const int qmax = max(L[0].qlen, L[1].qlen);

    // ---- row "-1" (bandedSWA.cpp:159-161) and zeroed E for both pairs
    {
        const int nel = duo_elems(qmax);
        for (int k = 0; k < nel; ++k) {
            uint32_t hw[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * k + u;
                uint32_t v[2];
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    int x = j == 0 ? L[a].h0 : max(L[a].h0 - oe_ins - (j - 1) * P.e_ins, 0);
                    if (j > L[a].qlen || L[a].qlen == 0) x = 0;   // the reference's calloc'ed tail
                    v[a] = (uint32_t)x;
                }
                hw[u] = v[0] | (v[1] << 16);
            }
            uint4 w; w.x = hw[0]; w.y = 0u; w.z = hw[1]; w.w = 0u;
            R.HE4(k) = w;
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        DuoLane &l = L[a];
        l.band = pair_band(P, l.qlen);
        l.budget = l.qlen ? min(l.qlen + l.band, l.tlen) : 0;
        l.beg = 0; l.end = l.qlen;
        l.best = l.h0; l.best_i = -1; l.best_j = -1; l.g_i = -1; l.gsc = -1; l.off = 0;
        l.hcol = l.h0 - P.o_del;
        l.xbeg = 0; l.cells = 0; l.tword = 0;
        l.done = l.qlen == 0 || l.tlen == 0;
    }

    for (int i = 0;; ++i) {
        // ---- per pair: band clamp, retirement (bandedSWA.cpp:183-185, 3035-3036, 3130-3144)
        bool act[2];
        int tcode[2], hleft[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            DuoLane &l = L[a];
            if (!l.done && i >= l.budget) l.done = true;
            if (!l.done) {
                if (l.beg < i - l.band) {
                    l.beg = i - l.band;
                    R.setHE16(l.beg - 1, a, 0u, 0u);   // the entry the clamp just passed must read as zero
                }
                if (l.end > i + l.band + 1) l.end = i + l.band + 1;
                if (l.beg >= l.end) l.done = true;
            }
            act[a] = !l.done;
            tcode[a] = 0; hleft[a] = 0;
            if (act[a]) {
                if (COUNT) {
                    if (l.xbeg < i - l.band) l.xbeg = i - l.band;
                    l.cells += (uint32_t)(l.end - l.xbeg);
                }
                if ((i & 7) == 0) l.tword = duo_target<TWIDE>(l, i >> 3);
                tcode[a] = (int)(l.tword & 7u);
                l.tword >>= 4;
                l.hcol -= P.e_del;
                hleft[a] = l.beg == 0 ? max(l.hcol, 0) : 0;
            }
        }
        if (!act[0] && !act[1]) break;

        // the row's target seeds: byte a = code | (code | 8) << 4, in both 16-bit halves of the word
        const uint32_t tsel = (((uint32_t)tcode[0] * 0x11u + 0x80u) | (((uint32_t)tcode[1] * 0x11u + 0x80u) << 8)) * 0x00010001u;

        const int eA = act[0] ? L[0].end : 0, eB = act[1] ? L[1].end : 0;
        const int bA = act[0] ? L[0].beg : 0x7FFFFFFF, bB = act[1] ? L[1].beg : 0x7FFFFFFF;
        const int ub = min(bA, bB), ue = max(eA, eB);
        const int emin = (act[0] && act[1]) ? min(eA, eB) : 0;   // columns [ub, emin) are live for both pairs

        uint32_t hprev = (uint32_t)hleft[0] | ((uint32_t)hleft[1] << 16);   // { H_A(i, j-1), H_B(i, j-1) }
        uint32_t F = 0;                                                      // { F_A(i, j), F_B(i, j) }
        uint32_t rm = 0;                                                     // row max per pair
        int mjA = -1, mjB = -1;

        // one column of both pairs; returns h
        auto column = [&](const uint32_t Hd, const uint32_t Ev, const uint32_t sel, uint32_t &Enew) -> uint32_t {
            const uint32_t sc = prmt_sx(LUT_LO, LUT_HI, sel);
            uint32_t M;
            if (FASTM) {
                M = __viaddmin_s16x2(Hd, sc, Hd * KM);
            } else {
                const uint32_t sm = __vmins2(sc, __vmins2(Hd, 0x00010001u) * (uint32_t)P.match);
                M = __vadd2(Hd, sm);
            }
            const uint32_t Tdel = __viaddmax_s16x2_relu(M, NEG_OE_DEL, NEG_OE_DEL);
            const uint32_t Tins = SYM ? Tdel : __viaddmax_s16x2_relu(M, NEG_OE_INS, NEG_OE_INS);
            Enew = __viaddmax_s16x2(Ev, NEG_E_DEL, Tdel);
            const uint32_t h = __vimax3_s16x2(M, Ev, F);
            F = __viaddmax_s16x2(F, NEG_E_INS, Tins);
            return h;
        };

        int j = ub & ~3;
        // ---- blocks of four columns, both pairs live
        for (; j + 3 < emin; j += 4) {
            const int k = j >> 1;
            const uint4 a = R.HE4(k), b = R.HE4(k + 1);
            const uint32_t q01 = R.QS2(k), q23 = R.QS2(k + 1);
            uint32_t s0, s1, s2, s3;
            if (TWIDE || BSW_SEL_LOP3) {
                s0 = sel_combine(q01, tsel, 0x44444444u); s1 = __umulhi(s0, K16);
                s2 = sel_combine(q23, tsel, 0x44444444u); s3 = __umulhi(s2, K16);
            } else {
                s0 = q01 * K1 + tsel; s1 = __umulhi(s0, K16);
                s2 = q23 * K1 + tsel; s3 = __umulhi(s2, K16);
            }
            uint4 oa, ob;
            bool phi, plo;
            const uint32_t h0v = column(a.x, a.y, s0, oa.y);
            oa.x = hprev;
            rm = __vibmax_s16x2(h0v, rm, &phi, &plo); if (plo) mjA = j; if (phi) mjB = j;
            const uint32_t h1v = column(a.z, a.w, s1, oa.w);
            oa.z = h0v;
            rm = __vibmax_s16x2(h1v, rm, &phi, &plo); if (plo) mjA = j + 1; if (phi) mjB = j + 1;
            const uint32_t h2v = column(b.x, b.y, s2, ob.y);
            ob.x = h1v;
            rm = __vibmax_s16x2(h2v, rm, &phi, &plo); if (plo) mjA = j + 2; if (phi) mjB = j + 2;
            const uint32_t h3v = column(b.z, b.w, s3, ob.w);
            ob.z = h2v;
            rm = __vibmax_s16x2(h3v, rm, &phi, &plo); if (plo) mjA = j + 3; if (phi) mjB = j + 3;
            hprev = h3v;
            R.HE4(k) = oa;
            R.HE4(k + 1) = ob;
        }
        // ---- remaining columns one at a time, masked per pair: a pair is live at column j < its end
        uint32_t hl = hprev;    // per pair: H(i, end - 1), taken when its last column goes by
        for (; j < ue; ++j) {
            const uint2 he = R.HE(j);
            const uint32_t q = half_of(R.QS2(j >> 1), j & 1);
            const uint32_t sel = (TWIDE || BSW_SEL_LOP3) ? sel_combine(q, tsel, 0x44444444u) : q * K1 + tsel;
            const uint32_t keep = (j < eA ? 0x0000FFFFu : 0u) | (j < eB ? 0xFFFF0000u : 0u);
            uint32_t En;
            uint32_t h = column(he.x, he.y, sel, En);
            // a pair that is not live keeps its stored entries and stays out of the row maximum
            R.HE(j) = make_uint2((hprev & keep) | (he.x & ~keep), (En & keep) | (he.y & ~keep));
            h &= keep;
            bool phi, plo;
            rm = __vibmax_s16x2(h, rm, &phi, &plo);
            if (plo && j < eA) mjA = j;
            if (phi && j < eB) mjB = j;
            hprev = h;
            if (j == eA - 1) hl = (hl & 0xFFFF0000u) | (h & 0xFFFFu);
            if (j == eB - 1) hl = (hl & 0x0000FFFFu) | (h & 0xFFFF0000u);
        }

        // ---- row end, per pair
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (!act[a]) continue;
            DuoLane &l = L[a];
            const int end = l.end;
            const int hlast = (int)half_of(hl, a);
            R.setHE16(end, a, (uint32_t)hlast, 0u);       // the reference's eh[end] = { h1, 0 }
            if (end == l.qlen) {                           // bandedSWA.cpp:218-221
                if (!(l.gsc > hlast)) l.g_i = i;
                l.gsc = max(l.gsc, hlast);
            }
            const int m = (int)(short)half_of(rm, a);
            const int mj = a ? mjB : mjA;
            if (m == 0) { l.done = true; continue; }
            if (m > l.best) {
                l.best = m; l.best_i = i; l.best_j = mj;
                l.off = max(l.off, abs(mj - i));
            } else {
                // vector z-drop rule: no gap-extend factor, no zdrop > 0 guard (bandedSWA.cpp:1889-1902)
                const int di = i - l.best_i, dj = mj - l.best_j;
                if (l.best - m - abs(di - dj) > P.zdrop) { l.done = true; continue; }
            }
            if (COUNT) {   // the reference's scan (bandedSWA.cpp:234-235), on the rows just written
                int x = l.xbeg;
                while (x < end && R.getH16(x, a) == 0 && R.getE16(x, a) == 0) ++x;
                l.xbeg = x;
            }
            // leading trim (not semantic: skipped cells are all-zero; lazy, four columns at a time)
            {
                const int k = (l.beg >> 2) << 1;
                const uint4 z0 = R.HE4(k), z1 = R.HE4(k + 1);
                const uint32_t z = z0.x | z0.y | z0.z | z0.w | z1.x | z1.y | z1.z | z1.w;
                if (half_of(z, a) == 0u) l.beg = 2 * k + 4;
            }
            // trailing trim (semantic): j* = last j in [beg, end] with Hs[j] | E[j] != 0; new end = min(j* + 2, qlen)
            if (hlast) {
                l.end = min(end + 2, l.qlen);
            } else {
                int js = end - 1;
                while (js >= 0 && half_of(R.HE(js).x | R.HE(js).y, a) == 0u) --js;
                l.end = min(js + 2, l.qlen);
            }
        }
    }

#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const DuoLane &l = L[a];
        PairResult r;
        r.score = l.best; r.qle = l.best_j + 1; r.tle = l.best_i + 1;
        r.gtle = l.g_i + 1; r.gscore = l.gsc; r.max_off = l.off;
        r.cells = l.cells;
        res[a] = r;
    }
}

}  // namespace bswk
