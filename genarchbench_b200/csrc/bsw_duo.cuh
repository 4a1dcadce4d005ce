// Thread-per-TWO-pairs DP, second generation ("duo"): the two 16-bit lanes of every DPX instruction are the SAME
// cell (i, j) of two DIFFERENT pairs A (low halves) and B (high halves) that sit next to each other in the
// device-sorted (len2, len1, h0) launch order -- the lane mapping of the reference's inter-pair SIMD kernel
// (bandedSWA.cpp:2977-3365: one pair per 16-bit lane), two lanes wide. Included by bsw_kernels.cuh.
//
// Against extend_pair (lanes = two adjacent columns of one pair) every cross-lane move is gone: F is ONE
// VIADDMNMX per column (a 1-instruction dependency chain instead of 4), the diagonal of the next row is simply
// the previous column's register (no shifted store), the row decisions of two pairs share one pass. Per column of
// both pairs (2 cells) the inner loop issues 6 ALU-pipe instructions for the recurrence (PRMT score, M, T, E', H,
// F') + 0.75 for the keyed row argmax -- 3.4 per cell against 4.4.
//
// Exactness per pair (row-ordered decisions of bandedSWA.cpp:183-237 per lane):
//   * both pairs walk their rows together over columns [j0, max(endA, endB)), j0 a shared multiple of 4;
//   * entries LEFT of a lane's beg read as zero and so compute to zero, exactly as if skipped: the band clamp
//     zeroes the one entry it passes per row, the (lazy, joint) leading trim only skips blocks that are all-zero for
//     BOTH lanes -- the reference's own leading trim (:234-235) skips nothing but all-zero entries either;
//   * entries at or RIGHT of a lane's end keep their stale values (the reference reads them again when `end`
//     grows by its 2 columns per row): columns the other lane still needs are computed for both and BLENDED back
//     (masked trips); when both ends are equal -- the common case for sorted neighbours, ~80 % of the rows of
//     config 3 -- the last partial block uses predicated stores instead and costs no masks;
//   * a lane whose pair is finished (row budget, m == 0, z-drop, empty column range) has its rows zeroed once and
//     then computes zeros: no masks at all for the rest of the other pair's rows.
#pragma once

namespace bswk {

// Row storage of one duo thread (interleaved by thread, stride = threads of the block):
//   he4[k]: uint4 = columns 2k, 2k+1:   .x = { HsA[2k],   HsB[2k]   }   .y = { EA[2k],   EB[2k]   }
//                                       .z = { HsA[2k+1], HsB[2k+1] }   .w = { EA[2k+1], EB[2k+1] }
//           with Hs[j] = H(i-1, j-1) (the reference's eh[j].h) and E[j] = E(i, j) (eh[j].e)
//   qs[b] : uint2 = selector seeds of columns 4b .. 4b+3, 16 bits each (.x: 4b, 4b+1; .y: 4b+2, 4b+3):
//           byte 0 = pair A's base * 0x11, byte 1 = pair B's
struct RowsD {
    uint4 *he4;
    uint2 *qs;
    int stride;
    __device__ __forceinline__ uint4 &HE4(int k) const { return he4[(size_t)k * stride]; }
    __device__ __forceinline__ uint2 &HE(int j) const {   // column j: .x = Hs halves, .y = E halves
        return reinterpret_cast<uint2 *>(he4 + (size_t)(j >> 1) * stride)[j & 1];
    }
    __device__ __forceinline__ uint2 &QS(int b) const { return qs[(size_t)b * stride]; }
    // Hs_a[j] = hv, E_a[j] = ev, the other pair's halves untouched
    __device__ __forceinline__ void setHE16(int j, int a, uint32_t hv, uint32_t ev) const {
#ifdef BSW_HOST_EMUL
        uint2 &p = HE(j);
        uint2 v = p;
        if (a) { v.x = (v.x & 0xFFFFu) | (hv << 16); v.y = (v.y & 0xFFFFu) | (ev << 16); }
        else { v.x = (v.x & 0xFFFF0000u) | (hv & 0xFFFFu); v.y = (v.y & 0xFFFF0000u) | (ev & 0xFFFFu); }
        p = v;
#else
        unsigned char *p = reinterpret_cast<unsigned char *>(&HE(j)) + 2 * a;
        const uint32_t sp = (uint32_t)__cvta_generic_to_shared(p);
        asm volatile("st.shared.u16 [%0], %1;\n\tst.shared.u16 [%0+4], %2;" ::"r"(sp), "h"((unsigned short)hv),
                     "h"((unsigned short)ev) : "memory");
#endif
    }
    __device__ __forceinline__ void setHE(int j, uint32_t hw, uint32_t ew) const {   // both pairs' halves
#ifdef BSW_HOST_EMUL
        HE(j) = make_uint2(hw, ew);
#else
        const uint32_t sp = (uint32_t)__cvta_generic_to_shared(&HE(j));
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sp), "r"(hw), "r"(ew) : "memory");
#endif
    }
};

// 4-column blocks a duo thread needs for queries up to qmax bases: columns 0 .. qmax
__host__ __device__ inline int duo_blocks(int qmax) { return (qmax + 4) >> 2; }
__host__ __device__ inline uint32_t duo2_thread_bytes(int qmax) { return 40u * (uint32_t)duo_blocks(qmax); }

struct DuoIn {
    int qlen, tlen, h0;          // qlen == 0 || tlen == 0: no DP for this lane
    const uint32_t *blob;        // packed [query | target]
    bool wide;                   // the blob is 4-bit (the pair holds an ambiguous base)
};

// four consecutive bases of a packed query starting at base 4 * b4, one per byte
__device__ __forceinline__ uint32_t duo_bases4(const DuoIn &L, int b4) {
    uint32_t v;
    if (!L.wide) {
        v = (L.blob[b4 >> 2] >> (8 * (b4 & 3))) & 0xFFu;
        v = (v | (v << 12)) & 0x000F000Fu;
        v = (v | (v << 6)) & 0x03030303u;
    } else {
        v = (L.blob[b4 >> 1] >> (16 * (b4 & 1))) & 0xFFFFu;
        v = (v | (v << 8)) & 0x00FF00FFu;
        v = (v | (v << 4)) & 0x0F0F0F0Fu;
    }
    return v;
}

// The DP of the two pairs of a thread; results in res[0], res[1].
//   FASTM, SYM : as in extend_pair
//   TWIDE      : at least one pair of the WARP may hold an ambiguous base (LOP3 selector; see score_lut)
//   KEY        : row argmax by key = score << kbits | column (P.kkey = 1 << P.kbits): every score of the launch is
//                < 2^(16 - kbits) and every column index < 2^kbits
template <bool FASTM, bool SYM, bool TWIDE, bool KEY>
__device__ inline void extend_duo2(const RowsD &R, const DuoIn *L, const KParams &P, PairResult *res) {
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const uint32_t NEG_OE_DEL = pack2(-oe_del), NEG_OE_INS = pack2(-oe_ins);
    const uint32_t NEG_E_DEL = pack2(-P.e_del), NEG_E_INS = pack2(-P.e_ins);
    uint32_t LUT_LO, LUT_HI;
    score_lut<TWIDE>(P, LUT_LO, LUT_HI);
    uint32_t K16 = P.k16, KM = P.km, K1 = P.k1, KK = P.kkey;
#if !defined(BSW_HOST_EMUL) && BSW_PIN_CONSTS
    {   // as in extend_pair: keeps ptxas from re-loading them from the parameter bank inside every trip
        const uint32_t z = *reinterpret_cast<const volatile uint32_t *>(&g_zero);
        K16 ^= z; KM ^= z; K1 ^= z; KK ^= z;
    }
#endif
    const uint32_t KBITS = P.kbits;

    // ---- per lane state
    int qlen[2], budget[2], band[2], end[2], best[2], best_i[2], best_j[2], g_i[2], gsc[2], off[2];
    bool live[2];
    const uint32_t *tb[2];
    uint32_t traw[2], tnext[2];
    int tlast[2];                 // last word of the packed target that may be read
    int hcol0[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const DuoIn &l = L[a];
        qlen[a] = l.qlen;
        band[a] = pair_band(P, l.qlen);
        live[a] = l.qlen > 0 && l.tlen > 0;
        budget[a] = live[a] ? min(l.qlen + band[a], l.tlen) : 0;
        end[a] = l.qlen;
        best[a] = l.h0; best_i[a] = -1; best_j[a] = -1; g_i[a] = -1; gsc[a] = -1; off[a] = 0;
        tb[a] = l.blob + (seq_bytes((uint32_t)l.qlen, l.wide) >> 2);
        tlast[a] = l.tlen > 0 ? (l.wide ? (l.tlen - 1) >> 3 : (l.tlen - 1) >> 4) : 0;
        traw[a] = 0;
        tnext[a] = live[a] ? tb[a][0] : 0u;
        hcol0[a] = live[a] ? min(l.h0 - P.o_del, 32767) : -1;
    }
    const int qmax = max(live[0] ? qlen[0] : 0, live[1] ? qlen[1] : 0);
    const int nblk = duo_blocks(qmax);

    // ---- selector seeds and row "-1" (bandedSWA.cpp:159-161): Hs[0] = h0, Hs[j] = max(h0 - oe_ins - (j-1) e_ins, 0)
    // for 1 <= j <= qlen, 0 beyond (the reference's calloc'ed tail); E = 0
    for (int b = 0; b < nblk; ++b) {
        uint32_t va = 0, vb = 0;
        if (live[0] && 4 * b < qlen[0]) va = duo_bases4(L[0], b);
        if (live[1] && 4 * b < qlen[1]) vb = duo_bases4(L[1], b);
        // (bases past a query's end are padding zeros of the blob or, in a 4-bit blob, whatever follows: mask them)
        uint32_t hw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = 4 * b + u;
            uint32_t v[2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                int x = j == 0 ? L[a].h0 : max(L[a].h0 - oe_ins - (j - 1) * P.e_ins, 0);
                if (j > qlen[a] || !live[a]) x = 0;
                v[a] = (uint32_t)x;
            }
            hw[u] = v[0] | (v[1] << 16);
            if (j >= qlen[0]) va &= ~(0xFFu << (8 * u));
            if (j >= qlen[1]) vb &= ~(0xFFu << (8 * u));
        }
        va *= 0x11u; vb *= 0x11u;
        uint2 q;
        q.x = (va & 0xFFu) | ((vb & 0xFFu) << 8) | ((va & 0xFF00u) << 8) | ((vb & 0xFF00u) << 16);
        q.y = ((va >> 16) & 0xFFu) | (((vb >> 16) & 0xFFu) << 8) | (((va >> 24) & 0xFFu) << 16) | ((vb >> 24) << 24);
        R.QS(b) = q;
        uint4 w0, w1;
        w0.x = hw[0]; w0.y = 0u; w0.z = hw[1]; w0.w = 0u;
        w1.x = hw[2]; w1.y = 0u; w1.z = hw[3]; w1.w = 0u;
        R.HE4(2 * b) = w0;
        R.HE4(2 * b + 1) = w1;
    }

    // a finished lane computes zeros from here on: its halves of every entry are cleared once
    auto retire = [&](int a) {
        live[a] = false;
        if (!live[a ^ 1]) return;
        const uint32_t keep = a ? 0x0000FFFFu : 0xFFFF0000u;
        for (int k = 0; k < 2 * nblk; ++k) {
            uint4 w = R.HE4(k);
            w.x &= keep; w.y &= keep; w.z &= keep; w.w &= keep;
            R.HE4(k) = w;
        }
    };

    uint32_t HCOL = ((uint32_t)hcol0[0] & 0xFFFFu) | ((uint32_t)hcol0[1] << 16);   // h0 - o_del - e_del * i, floored at -1
    int j0 = 0;            // first column of the rows (multiple of 4): everything left of it is zero for both lanes
    uint32_t tcode[2] = {0u, 0u};

    for (int i = 0; live[0] || live[1]; ++i) {
        // ---- per lane: row budget, band clamp (bandedSWA.cpp:183-185, 3035-3036, 3130-3144), target base
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (!live[a]) continue;
            if (i >= budget[a]) { retire(a); continue; }
            const int cb = i - band[a];                     // the clamp's beg
            if (cb > 0) {
                R.setHE16(cb - 1, a, 0u, 0u);               // the entry the clamp passes must read as zero
                HCOL = a ? (HCOL | 0xFFFF0000u) : (HCOL | 0x0000FFFFu);   // beg > 0: H(i, beg - 1) = 0 from now on
            }
            if (end[a] > i + band[a] + 1) end[a] = i + band[a] + 1;
            if (cb >= end[a]) { retire(a); continue; }
            if (L[a].wide) {
                if ((i & 7) == 0) { traw[a] = tnext[a]; tnext[a] = tb[a][min((i >> 3) + 1, tlast[a])]; }
                tcode[a] = traw[a] & 7u;
                traw[a] >>= 4;
            } else {
                if ((i & 15) == 0) { traw[a] = tnext[a]; tnext[a] = tb[a][min((i >> 4) + 1, tlast[a])]; }
                tcode[a] = traw[a] & 3u;
                traw[a] >>= 2;
            }
        }
        if (!live[0] && !live[1]) break;
        // the row's target seeds: byte a = c | (c | 8) << 4 with c = the base (LOP3 selector) or 4 - base (add
        // selector), in both 16-bit halves of the word
        uint32_t tsel;
        if (TWIDE || BSW_SEL_LOP3) tsel = tcode[0] * 0x00110011u + tcode[1] * 0x11001100u + 0x80808080u;
        else tsel = (4u - tcode[0]) * 0x00110011u + (4u - tcode[1]) * 0x11001100u + 0x80808080u;

        // first column: H(i, -1) = max(h0 - o_del - e_del * (i + 1), 0) while beg == 0
        HCOL = __viaddmax_s16x2(HCOL, NEG_E_DEL, 0xFFFFFFFFu);
        const uint32_t hleft = __vmaxs2(HCOL, 0u);

        // a finished lane follows the other one's end (it needs no masks: all of its entries are zero)
        const int eA = live[0] ? end[0] : end[1], eB = live[1] ? end[1] : end[0];
        const int emin = min(eA, eB), emax = max(eA, eB);

        uint32_t hprev = hleft;     // { H_A(i, j-1), H_B(i, j-1) }
        uint32_t F = 0;             // { F_A(i, j), F_B(i, j) }
        uint32_t rm = 0;            // KEY: running max of the keys; else running max of the scores
        int mjA = -1, mjB = -1;     // !KEY: last column where the lane reached rm

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
        // selectors of the four columns of a block
        auto selectors = [&](const uint2 q, uint32_t &s0, uint32_t &s1, uint32_t &s2, uint32_t &s3) {
            if (TWIDE || BSW_SEL_LOP3) {
                s0 = sel_combine(q.x, tsel, 0x44444444u); s2 = sel_combine(q.y, tsel, 0x44444444u);
            } else {
                s0 = q.x * K1 + tsel; s2 = q.y * K1 + tsel;
            }
            s1 = __umulhi(s0, K16); s3 = __umulhi(s2, K16);
        };

        int j = j0;
        // ---- blocks of four columns that are live for both pairs
        if (j + 4 <= emin) {
            auto trip = [&](const uint4 &a, const uint4 &b, const uint2 &q, uint4 &na, uint4 &nb, uint2 &nq) -> bool {
                const int k = j >> 1;
                const bool more = j + 8 <= emin;
                if (more) { na = R.HE4(k + 2); nb = R.HE4(k + 3); nq = R.QS((j >> 2) + 1); }
                uint32_t s0, s1, s2, s3;
                selectors(q, s0, s1, s2, s3);
                uint4 oa, ob;
                oa.x = hprev;
                const uint32_t h0v = column(a.x, a.y, s0, oa.y);
                oa.z = h0v;
                const uint32_t h1v = column(a.z, a.w, s1, oa.w);
                ob.x = h1v;
                const uint32_t h2v = column(b.x, b.y, s2, ob.y);
                ob.z = h2v;
                const uint32_t h3v = column(b.z, b.w, s3, ob.w);
                hprev = h3v;
                R.HE4(k) = oa;
                R.HE4(k + 1) = ob;
                if (KEY) {
                    // the later column wins ties, as `h >= m` does in the reference (bandedSWA.cpp:204-205)
                    const uint32_t t3 = __vimax3_u16x2(h0v * KK, h1v * KK + 0x00010001u, h2v * KK + 0x00020002u);
                    const uint32_t t4 = __vmaxu2(t3, h3v * KK + 0x00030003u);
                    rm = __viaddmax_u16x2(t4, (uint32_t)j * 0x00010001u, rm);
                } else {
                    bool phi, plo;
                    rm = __vibmax_s16x2(h0v, rm, &phi, &plo); if (plo) mjA = j;     if (phi) mjB = j;
                    rm = __vibmax_s16x2(h1v, rm, &phi, &plo); if (plo) mjA = j + 1; if (phi) mjB = j + 1;
                    rm = __vibmax_s16x2(h2v, rm, &phi, &plo); if (plo) mjA = j + 2; if (phi) mjB = j + 2;
                    rm = __vibmax_s16x2(h3v, rm, &phi, &plo); if (plo) mjA = j + 3; if (phi) mjB = j + 3;
                }
                j += 4;
                return more;
            };
            uint4 a0 = R.HE4(j >> 1), b0 = R.HE4((j >> 1) + 1);
            uint2 q0 = R.QS(j >> 2);
            uint4 a1, b1;           // written by the first trip before the second reads them
            uint2 q1;
            for (;;) {
                if (!trip(a0, b0, q0, a1, b1, q1)) break;
                if (!trip(a1, b1, q1, a0, b0, q0)) break;
            }
        }

        uint32_t hl = hprev;        // per lane: H(i, end - 1)
        if (eA == eB) {
            const int n = eA - j;   // live columns left: 0 .. 3 (more only if the row starts right of `end`: n <= 0)
            if (n > 0) {
                // ---- the last, partial block: columns u < n are live, column n is the reference's eh[end] = { h1, 0 },
                // everything right of it keeps its stale value (predicated stores instead of masks)
                const int k = j >> 1;
                const uint4 a = R.HE4(k), b = R.HE4(k + 1);
                const uint2 q = R.QS(j >> 2);
                uint32_t s0, s1, s2, s3;
                selectors(q, s0, s1, s2, s3);
                (void)s3;
                const bool p1 = n > 1, p2 = n > 2;
                uint32_t E0, E1, E2;
                const uint32_t h0v = column(a.x, a.y, s0, E0);
                const uint32_t h1v = column(a.z, a.w, s1, E1);
                const uint32_t h2v = column(b.x, b.y, s2, E2);
                uint4 oa;
                oa.x = hprev; oa.y = E0; oa.z = h0v; oa.w = p1 ? E1 : 0u;
                R.HE4(k) = oa;
                if (p1) R.setHE(j + 2, h1v, p2 ? E2 : 0u);
                if (p2) R.setHE(j + 3, h2v, 0u);
                hl = p2 ? h2v : (p1 ? h1v : h0v);
                if (KEY) {
                    uint32_t t = h0v * KK;
                    if (p1) t = __vmaxu2(t, h1v * KK + 0x00010001u);
                    if (p2) t = __vmaxu2(t, h2v * KK + 0x00020002u);
                    rm = __viaddmax_u16x2(t, (uint32_t)j * 0x00010001u, rm);
                } else {
                    bool phi, plo;
                    rm = __vibmax_s16x2(h0v, rm, &phi, &plo); if (plo) mjA = j; if (phi) mjB = j;
                    if (p1) { rm = __vibmax_s16x2(h1v, rm, &phi, &plo); if (plo) mjA = j + 1; if (phi) mjB = j + 1; }
                    if (p2) { rm = __vibmax_s16x2(h2v, rm, &phi, &plo); if (plo) mjA = j + 2; if (phi) mjB = j + 2; }
                }
            } else {
                R.setHE(eA, hl, 0u);                       // eh[end] = { h1, 0 } of both lanes
            }
        } else {
            // ---- ends differ: blocks up to the larger end with per-lane masks. keep = the lane is live at this column
            for (; j < emax; j += 4) {
                const int k = j >> 1;
                const uint4 a = R.HE4(k), b = R.HE4(k + 1);
                const uint2 q = R.QS(j >> 2);
                uint32_t s[4];
                selectors(q, s[0], s[1], s[2], s[3]);
                const uint32_t hd[4] = {a.x, a.z, b.x, b.z}, ev[4] = {a.y, a.w, b.y, b.w};
                uint32_t oh[4], oe[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t keep = (j + u < eA ? 0x0000FFFFu : 0u) | (j + u < eB ? 0xFFFF0000u : 0u);
                    uint32_t En;
                    uint32_t h = column(hd[u], ev[u], s[u], En);
                    oh[u] = (hprev & keep) | (hd[u] & ~keep);
                    oe[u] = (En & keep) | (ev[u] & ~keep);
                    h &= keep;
                    hl = h | (hl & ~keep);
                    if (KEY) {
                        rm = __vmaxu2(rm, h * KK + (uint32_t)(j + u) * 0x00010001u);
                    } else {
                        bool phi, plo;
                        rm = __vibmax_s16x2(h, rm, &phi, &plo);
                        if (plo && j + u < eA) mjA = j + u;
                        if (phi && j + u < eB) mjB = j + u;
                    }
                    hprev = h;
                }
                uint4 oa, ob;
                oa.x = oh[0]; oa.y = oe[0]; oa.z = oh[1]; oa.w = oe[1];
                ob.x = oh[2]; ob.y = oe[2]; ob.z = oh[3]; ob.w = oe[3];
                R.HE4(k) = oa;
                R.HE4(k + 1) = ob;
            }
            if (live[0]) R.setHE16(end[0], 0, hl & 0xFFFFu, 0u);     // eh[end] = { h1, 0 }
            if (live[1]) R.setHE16(end[1], 1, hl >> 16, 0u);
        }

        // first block of the row for the (joint) leading trim below; loaded here so that its latency hides
        // behind the row decisions
        const uint4 z0 = R.HE4(j0 >> 1), z1 = R.HE4((j0 >> 1) + 1);

        // ---- row end, per lane (bandedSWA.cpp:217-237)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (!live[a]) continue;
            const int e = end[a];
            const int hlast = (int)(a ? hl >> 16 : hl & 0xFFFFu);
            if (e == qlen[a]) {                           // :218-221
                if (!(gsc[a] > hlast)) g_i[a] = i;
                gsc[a] = max(gsc[a], hlast);
            }
            int m, mj;
            if (KEY) {
                const uint32_t kx = a ? rm >> 16 : rm & 0xFFFFu;
                m = (int)(kx >> KBITS);
                mj = (int)(kx & (KK - 1u));
            } else {
                m = (int)(short)(a ? rm >> 16 : rm & 0xFFFFu);
                mj = a ? mjB : mjA;
            }
            if (m == 0) { retire(a); continue; }
            if (m > best[a]) {
                best[a] = m; best_i[a] = i; best_j[a] = mj;
                off[a] = max(off[a], abs(mj - i));
            } else {
                // vector z-drop rule: no gap-extend factor, no zdrop > 0 guard (bandedSWA.cpp:1889-1902)
                const int di = i - best_i[a], dj = mj - best_j[a];
                if (best[a] - m - abs(di - dj) > P.zdrop) { retire(a); continue; }
            }
            // trailing trim (semantic): j* = last j <= end with Hs[j] | E[j] != 0 (m > 0 guarantees one); the new end
            // is min(j* + 2, qlen). Hs[end] = H(i, end - 1) is almost always non-zero: tested first.
            if (hlast) {
                end[a] = min(e + 2, qlen[a]);
            } else {
                int js = e - 1;
                for (; js >= 0; --js) {
                    const uint2 w = R.HE(js);
                    const uint32_t x = w.x | w.y;
                    if (a ? x >> 16 : x & 0xFFFFu) break;
                }
                end[a] = min(js + 2, qlen[a]);
            }
        }
        // ---- joint leading trim (not semantic: skipped cells are all-zero for both lanes; lazy, a block at a time)
        if ((z0.x | z0.y | z0.z | z0.w | z1.x | z1.y | z1.z | z1.w) == 0u) {
            const int lim = min(live[0] ? end[0] : 0x7FFFFFFF, live[1] ? end[1] : 0x7FFFFFFF);
            if (j0 + 4 <= lim && j0 + 4 <= emin) j0 += 4;
        }
    }

#pragma unroll
    for (int a = 0; a < 2; ++a) {
        PairResult r;
        r.score = best[a]; r.qle = best_j[a] + 1; r.tle = best_i[a] + 1;
        r.gtle = g_i[a] + 1; r.gscore = gsc[a]; r.max_off = off[a];
        r.cells = 0;
        res[a] = r;
    }
}

}  // namespace bswk
