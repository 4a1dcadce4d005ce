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
    uint32_t sbase;      // shared-space address of he4 (the 16- and 64-bit stores below use 32-bit addresses)
    __device__ __forceinline__ uint32_t saddr(int j) const {
        return sbase + (uint32_t)(j >> 1) * (uint32_t)(stride * 16) + (uint32_t)(j & 1) * 8u;
    }
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
        const uint32_t sp = saddr(j) + 2u * (uint32_t)a;
        asm volatile("st.shared.u16 [%0], %1;\n\tst.shared.u16 [%0+4], %2;" ::"r"(sp), "h"((unsigned short)hv),
                     "h"((unsigned short)ev) : "memory");
#endif
    }
    __device__ __forceinline__ void setHE(int j, uint32_t hw, uint32_t ew) const {   // both pairs' halves
#ifdef BSW_HOST_EMUL
        HE(j) = make_uint2(hw, ew);
#else
        const uint32_t sp = saddr(j);
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

// Clears pair a's halves of every entry of a duo thread's rows (see extend_duo2: a finished lane computes zeros).
__device__ __noinline__ void duo_clear_lane(uint4 *he4, int stride, int nelem, int a) {
    const uint32_t keep = a ? 0x0000FFFFu : 0xFFFF0000u;
    for (int k = 0; k < nelem; ++k) {
        uint4 w = he4[(size_t)k * stride];
        w.x &= keep; w.y &= keep; w.z &= keep; w.w &= keep;
        he4[(size_t)k * stride] = w;
    }
}

// byte u of the result = 0xFF iff u < n (n may be negative or beyond 4)
__device__ __forceinline__ uint32_t duo_bytemask(int n) {
#ifdef BSW_HOST_EMUL
    return n <= 0 ? 0u : (n >= 4 ? 0xFFFFFFFFu : ((1u << (8 * n)) - 1u));
#else
    return __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)(8 * max(n, 0)));
#endif
}

// Per-lane row decisions of bandedSWA.cpp:217-237 on scalars (straight-line; the caller handles `dead`):
//   gscore / gtle when the row reached the query end, m == 0, a new best with its max_off, else the VECTOR z-drop
//   rule (no gap-extend factor, no zdrop > 0 guard; bandedSWA.cpp:1889-1902), and the trailing trim's common case
//   end = min(end + 2, qlen) (Hs[end] = H(i, end - 1) != 0). Returns true when the pair is finished; *scan = the
//   trailing trim has to look for the last non-zero entry (hlast == 0).
struct DuoLaneState {
    int qlen, budget, band1, end, best, best_i, best_j, g_i, gsc, off;
};
__device__ __forceinline__ bool duo_lane_end(DuoLaneState &l, int i, int hlast, int m, int mj, int zdrop, bool *scan) {
    const int e = l.end;
    if (e == l.qlen) {                            // :218-221
        if (!(l.gsc > hlast)) l.g_i = i;
        l.gsc = max(l.gsc, hlast);
    }
    const bool better = m > l.best;
    const int di = i - l.best_i, dj = mj - l.best_j;
    const bool drop = l.best - m - abs(di - dj) > zdrop;
    const bool dead = m == 0 || (!better && drop);
    if (better) {
        l.best = m; l.best_i = i; l.best_j = mj;
        l.off = max(l.off, abs(mj - i));
    }
    *scan = !dead && hlast == 0;
    l.end = min(e + 2, l.qlen);
    return dead;
}

// The DP of the two pairs of a thread; results in res[0], res[1].
//   FASTM, SYM : as in extend_pair
//   TWIDE      : at least one pair of the WARP may hold an ambiguous base (LOP3 selector; see score_lut)
//   KEY        : row argmax by key = score << kbits | column (P.kkey = 1 << P.kbits): every score of the launch is
//                < 2^(16 - kbits) and every column index (up to the end of the last 4-column block) < 2^kbits
// Structure of a row (both pairs at once):
//   1. rare events by row number (a lane's row budget ends; its band clamp starts), the clamp's one zeroed entry
//      per row, the row's target bases -> selector seed, the first column's H;
//   2. FAST trips of four columns over [j0, min(endA, endB)) rounded down to whole blocks: no masks at all;
//   3. the LAST block(s) up to max(endA, endB): per-column byte masks say which lane is still left of its end; the
//      other lane's stored entries are blended back (they keep their stale values, which the reference reads again
//      when `end` grows) and its H stays out of the row maximum. Every thread runs this block every row, so the
//      threads of a warp whose two ends differ cost the others nothing;
//   4. the reference's eh[end] = { h1, 0 }, then the row decisions per lane, the joint leading trim.
// The rows run in two loops: a lean one while BOTH pairs are live (no liveness tests anywhere; it ends when a pair
// does), then, for the few rows one pair outlasts the other by, a compact general one (masked blocks only).
template <bool FASTM, bool SYM, bool TWIDE, bool KEY>
__device__ inline void extend_duo2(const RowsD &R, const DuoIn *L, const KParams &P, PairResult *res) {
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    uint32_t NEG_OE_DEL = pack2(-oe_del), NEG_OE_INS = pack2(-oe_ins);
    uint32_t NEG_E_DEL = pack2(-P.e_del), NEG_E_INS = pack2(-P.e_ins);
    uint32_t LUT_LO, LUT_HI;
    score_lut<TWIDE>(P, LUT_LO, LUT_HI);
    uint32_t K16 = P.k16, KM = P.km, K1 = P.k1, KK = P.kkey;
    int ZDROP = P.zdrop;
#if !defined(BSW_HOST_EMUL) && BSW_PIN_CONSTS
    {   // made opaque by a run-time zero from global memory: ptxas otherwise re-loads (or re-derives) them from
        // the parameter bank inside every trip
        const uint32_t z = *reinterpret_cast<const volatile uint32_t *>(&g_zero);
        K16 ^= z; KM ^= z; K1 ^= z; KK ^= z;
        NEG_OE_DEL ^= z; NEG_OE_INS ^= z; NEG_E_DEL ^= z; NEG_E_INS ^= z; LUT_LO ^= z; LUT_HI ^= z;
        ZDROP ^= (int)z;
    }
#endif
    const uint32_t KBITS = P.kbits;
    const int S = R.stride;

    // ---- per lane state
    DuoLaneState A, B;
    bool liveA, liveB, clampA = false, clampB = false;
    const uint32_t *tbA, *tbB;
    uint32_t trawA = 0, trawB = 0, tnextA, tnextB;
    int tlastA, tlastB;           // last word of the packed target that may be read
    int hcol0[2];
    {
        DuoLaneState *ls[2] = {&A, &B};
        bool lv[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const DuoIn &l = L[a];
            DuoLaneState &x = *ls[a];
            x.qlen = l.qlen;
            x.band1 = pair_band(P, l.qlen) + 1;
            lv[a] = l.qlen > 0 && l.tlen > 0;
            x.budget = lv[a] ? min(l.qlen + x.band1 - 1, l.tlen) : 0;
            x.end = l.qlen;
            x.best = l.h0; x.best_i = -1; x.best_j = -1; x.g_i = -1; x.gsc = -1; x.off = 0;
            hcol0[a] = lv[a] ? min(l.h0 - P.o_del, 32767) : -1;
        }
        liveA = lv[0]; liveB = lv[1];
        tbA = L[0].blob + (seq_bytes((uint32_t)L[0].qlen, L[0].wide) >> 2);
        tbB = L[1].blob + (seq_bytes((uint32_t)L[1].qlen, L[1].wide) >> 2);
        tlastA = L[0].tlen > 0 ? (L[0].wide ? (L[0].tlen - 1) >> 3 : (L[0].tlen - 1) >> 4) : 0;
        tlastB = L[1].tlen > 0 ? (L[1].wide ? (L[1].tlen - 1) >> 3 : (L[1].tlen - 1) >> 4) : 0;
        tnextA = liveA ? tbA[0] : 0u;
        tnextB = liveB ? tbB[0] : 0u;
    }
    const int qmax = max(liveA ? A.qlen : 0, liveB ? B.qlen : 0);
    const int nblk = duo_blocks(qmax);

    // ---- selector seeds and row "-1" (bandedSWA.cpp:159-161): Hs[0] = h0, Hs[j] = max(h0 - oe_ins - (j-1) e_ins, 0)
    // for 1 <= j <= qlen, 0 beyond (the reference's calloc'ed tail); E = 0
    for (int b = 0; b < nblk; ++b) {
        uint32_t va = 0, vb = 0;
        if (liveA && 4 * b < A.qlen) va = duo_bases4(L[0], b);
        if (liveB && 4 * b < B.qlen) vb = duo_bases4(L[1], b);
        uint32_t hw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = 4 * b + u;
            int xa = j == 0 ? L[0].h0 : max(L[0].h0 - oe_ins - (j - 1) * P.e_ins, 0);
            int xb = j == 0 ? L[1].h0 : max(L[1].h0 - oe_ins - (j - 1) * P.e_ins, 0);
            if (j > A.qlen || !liveA) xa = 0;
            if (j > B.qlen || !liveB) xb = 0;
            hw[u] = (uint32_t)xa | ((uint32_t)xb << 16);
            // (bases past a query's end: padding of the blob or, in a 4-bit blob, the target that follows)
            if (j >= A.qlen) va &= ~(0xFFu << (8 * u));
            if (j >= B.qlen) vb &= ~(0xFFu << (8 * u));
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

    uint32_t HCOL = ((uint32_t)hcol0[0] & 0xFFFFu) | ((uint32_t)hcol0[1] << 16);   // h0 - o_del - e_del * i, floored at -1
    int i = 0;
    int j0 = 0;            // first column of the rows (multiple of 4): everything left of it is zero for both lanes
    uint4 *p0 = R.he4;     // &HE4(j0 >> 1), &QS(j0 >> 2)
    uint2 *pq0 = R.qs;

    // the row's target base of a lane (packed 2 bits per base; 4 in a blob that holds an ambiguous base)
    auto target_code = [&](const bool wide, const uint32_t *tb, int tlast, uint32_t &traw, uint32_t &tnext) -> uint32_t {
        uint32_t c;
        if (TWIDE && wide) {
            if ((i & 7) == 0) { traw = tnext; tnext = tb[min((i >> 3) + 1, tlast)]; }
            c = traw & 7u;
            traw >>= 4;
        } else {
            if ((i & 15) == 0) { traw = tnext; tnext = tb[min((i >> 4) + 1, tlast)]; }
            c = traw & 3u;
            traw >>= 2;
        }
        return c;
    };
    // selector seed of the row: byte a = c | (c | 8) << 4 with c = the base (LOP3 selector) or 4 - base (add selector),
    // in both 16-bit halves of the word
    auto target_seed = [&](uint32_t ca, uint32_t cb) -> uint32_t {
        uint32_t tsel;
        if (TWIDE || BSW_SEL_LOP3) tsel = ca * 0x00110011u + cb * 0x11001100u + 0x80808080u;
        else tsel = 0xC4C4C4C4u - (ca * 0x00110011u + cb * 0x11001100u);
#if !defined(BSW_HOST_EMUL)
        asm volatile("" : "+r"(tsel));    // computed once per row, not re-derived inside the trips
#endif
        return tsel;
    };
    uint32_t tsel = 0, hprev = 0, F = 0, rm = 0;
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
    // masked blocks over columns [j, emax): keep = the lane is still left of its end; -> hl = per lane H(i, end - 1)
    auto masked_blocks = [&](int j, const int eA, const int eB, const int emax) -> uint32_t {
        uint32_t hl = hprev;
#pragma unroll 1
        for (; j < emax; j += 4) {
            const int k = j >> 1;
            const uint4 a = R.HE4(k), b = R.HE4(k + 1);
            const uint2 q = R.QS(j >> 2);
            const uint32_t MA = duo_bytemask(eA - j), MB = duo_bytemask(eB - j);
            uint32_t s[4];
            selectors(q, s[0], s[1], s[2], s[3]);
            const uint32_t hd[4] = {a.x, a.z, b.x, b.z}, ev[4] = {a.y, a.w, b.y, b.w};
            uint32_t oh[4], oe[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t keep = __byte_perm(MA, MB, 0x4400u + 0x1111u * (uint32_t)u);   // { A, A, B, B } of byte u
                uint32_t En;
                uint32_t h = column(hd[u], ev[u], s[u], En);
                oh[u] = (hprev & keep) | (hd[u] & ~keep);
                oe[u] = (En & keep) | (ev[u] & ~keep);
                hprev = h;
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
            }
            uint4 oa, ob;
            oa.x = oh[0]; oa.y = oe[0]; oa.z = oh[1]; oa.w = oe[1];
            ob.x = oh[2]; ob.y = oe[2]; ob.z = oh[3]; ob.w = oe[3];
            R.HE4(k) = oa;
            R.HE4(k + 1) = ob;
        }
        // the reference's eh[end] = { h1, 0 }
        if (eA == eB) {
            R.setHE(eA, hl, 0u);
        } else {
            R.setHE16(eA, 0, hl & 0xFFFFu, 0u);
            R.setHE16(eB, 1, hl >> 16, 0u);
        }
        return hl;
    };
    // row maximum and its last column of a lane
    auto row_max = [&](const int a, int &m, int &mj) {
        if (KEY) {
            const uint32_t kx = a ? rm >> 16 : rm & 0xFFFFu;
            m = (int)(kx >> KBITS);
            mj = (int)(kx & (KK - 1u));
        } else {
            m = (int)(short)(a ? rm >> 16 : rm & 0xFFFFu);
            mj = a ? mjB : mjA;
        }
    };
    // trailing trim, the rare case (bandedSWA.cpp:236-237): j* = last j < e with Hs[j] | E[j] != 0 (Hs[e] = E[e] = 0 were
    // just written; m > 0 guarantees an entry); the new end is min(j* + 2, qlen)
    auto trim_scan = [&](const int a, const int e, const int qlen) -> int {
        int js = e - 1;
#pragma unroll 1
        for (; js >= 0; --js) {
            const uint2 w = R.HE(js);
            const uint32_t x = w.x | w.y;
            if (a ? x >> 16 : x & 0xFFFFu) break;
        }
        return min(js + 2, qlen);
    };

    // =========================== the rows ===========================
    // A pair that is finished (row budget, m == 0, z-drop, empty column range) hands in its results and its lane
    // becomes a GHOST of the other pair: its halves of every entry are cleared once (so it computes zeros and needs
    // no masks), its first-column H is pinned at zero, and its control state (end, band, budget) mirrors the live
    // lane's, so the thread stays on the same code path as the rest of its warp until its second pair ends too.
    bool ghostA = !liveA, ghostB = !liveB;
    auto hand_in = [&](const int a, const DuoLaneState &l) {
        res[a].score = l.best; res[a].qle = l.best_j + 1; res[a].tle = l.best_i + 1;
        res[a].gtle = l.g_i + 1; res[a].gscore = l.gsc; res[a].max_off = l.off; res[a].cells = 0;
    };
    if (ghostA) { hand_in(0, A); HCOL |= 0x0000FFFFu; A = B; }
    if (ghostB) { hand_in(1, B); HCOL |= 0xFFFF0000u; B = A; }
    auto kill = [&](const int a) {
        if (a == 0) {
            hand_in(0, A); ghostA = true; HCOL |= 0x0000FFFFu;
            if (!ghostB) { duo_clear_lane(R.he4, S, 2 * nblk, 0); A = B; clampA = clampB; }
        } else {
            hand_in(1, B); ghostB = true; HCOL |= 0xFFFF0000u;
            if (!ghostA) { duo_clear_lane(R.he4, S, 2 * nblk, 1); B = A; clampB = clampA; }
        }
    };
    int event = 0;         // next row at which something rare happens (recomputed on the first row)
#pragma unroll 1
    while (!(ghostA && ghostB)) {
        // ---- rare: a row budget ends (bandedSWA.cpp:3035-3036), a band clamp starts to move beg (:183)
        if (i >= event) {
            if (!ghostA && i >= A.budget) kill(0);
            if (!ghostB && i >= B.budget) kill(1);
            if (ghostA && ghostB) break;
            if (i >= A.band1) { clampA = true; HCOL |= 0x0000FFFFu; }   // beg > 0 from now on: H(i, beg - 1) = 0
            if (i >= B.band1) { clampB = true; HCOL |= 0xFFFF0000u; }
            event = min(min(A.budget, B.budget), min(clampA ? 0x7FFFFFFF : A.band1, clampB ? 0x7FFFFFFF : B.band1));
        }
        // ---- band clamp (bandedSWA.cpp:183-185, 3130-3144)
        A.end = min(A.end, i + A.band1);
        B.end = min(B.end, i + B.band1);
        if (clampA | clampB) {
            const int cbA = i - A.band1 + 1, cbB = i - B.band1 + 1;     // the clamps' beg
            bool gone = false;
            if (clampA && !ghostA && cbA >= A.end) { kill(0); gone = true; }   // an empty column range ends a pair
            if (clampB && !ghostB && cbB >= B.end) { kill(1); gone = true; }
            if (gone) {
                if (ghostA && ghostB) break;
                event = i;             // the ghost took over the other lane's state: sort the events out again
                continue;
            }
            if (clampA) R.setHE16(cbA - 1, 0, 0u, 0u);      // the entry the clamp passes must read as zero
            if (clampB) R.setHE16(cbB - 1, 1, 0u, 0u);
        }
        tsel = target_seed(target_code(L[0].wide, tbA, tlastA, trawA, tnextA),
                           target_code(L[1].wide, tbB, tlastB, trawB, tnextB));
        // first column: H(i, -1) = max(h0 - o_del - e_del * (i + 1), 0) while beg == 0
        HCOL = __viaddmax_s16x2(HCOL, NEG_E_DEL, 0xFFFFFFFFu);
        hprev = __vmaxs2(HCOL, 0u);     // { H_A(i, j-1), H_B(i, j-1) }
        F = 0;                          // { F_A(i, j), F_B(i, j) }
        rm = 0;                         // KEY: running max of the keys; else running max of the scores
        const int eA = A.end, eB = B.end;
        const int emin = min(eA, eB), emax = max(eA, eB);

        int j = j0;
        // ---- FAST trips: blocks of four columns left of both ends
        int nt = (emin - j0) >> 2;
        if (nt > 0) {
            j += 4 * nt;
            uint4 *p = p0;
            uint2 *pq = pq0;
            uint32_t J2 = (uint32_t)j0 * 0x00010001u;
            int jj = j0;             // !KEY only
            auto trip = [&](const uint4 &a, const uint4 &b, const uint2 &q, uint4 &na, uint4 &nb, uint2 &nq) {
                if (nt > 1) { na = p[2 * S]; nb = p[3 * S]; nq = pq[S]; }
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
                p[0] = oa;
                p[S] = ob;
                if (KEY) {
                    // the later column wins ties, as `h >= m` does in the reference (bandedSWA.cpp:204-205)
                    const uint32_t t3 = __vimax3_u16x2(h0v * KK, h1v * KK + 0x00010001u, h2v * KK + 0x00020002u);
                    const uint32_t t4 = __vmaxu2(t3, h3v * KK + 0x00030003u);
                    rm = __viaddmax_u16x2(t4, J2, rm);
                    J2 += 0x00040004u;
                } else {
                    bool phi, plo;
                    rm = __vibmax_s16x2(h0v, rm, &phi, &plo); if (plo) mjA = jj;     if (phi) mjB = jj;
                    rm = __vibmax_s16x2(h1v, rm, &phi, &plo); if (plo) mjA = jj + 1; if (phi) mjB = jj + 1;
                    rm = __vibmax_s16x2(h2v, rm, &phi, &plo); if (plo) mjA = jj + 2; if (phi) mjB = jj + 2;
                    rm = __vibmax_s16x2(h3v, rm, &phi, &plo); if (plo) mjA = jj + 3; if (phi) mjB = jj + 3;
                    jj += 4;
                }
                p += 2 * S; pq += S;
                return --nt > 0;
            };
            uint4 a0 = p[0], b0 = p[S];
            uint2 q0 = pq[0];
            uint4 a1, b1;           // written by the first trip before the second reads them
            uint2 q1;
            for (;;) {
                if (!trip(a0, b0, q0, a1, b1, q1)) break;
                if (!trip(a1, b1, q1, a0, b0, q0)) break;
            }
        }
        // ---- LAST block(s) and eh[end]
        const uint32_t hl = masked_blocks(j, eA, eB, emax);
        // first block of the row for the (joint) leading trim below; loaded here so that its latency hides
        // behind the row decisions
        const uint4 z0 = p0[0], z1 = p0[S];
        // ---- row end per lane
        int mA, mB, cA, cB;
        row_max(0, mA, cA);
        row_max(1, mB, cB);
        bool scanA, scanB;
        const bool deadA = duo_lane_end(A, i, (int)(hl & 0xFFFFu), mA, cA, ZDROP, &scanA) && !ghostA;
        const bool deadB = duo_lane_end(B, i, (int)(hl >> 16), mB, cB, ZDROP, &scanB) && !ghostB;
        if (scanA && !ghostA) A.end = trim_scan(0, eA, A.qlen);
        if (scanB && !ghostB) B.end = trim_scan(1, eB, B.qlen);
        ++i;
        if (deadA | deadB) {           // rare
            if (deadA) kill(0);
            if (deadB) kill(1);
            event = i;
        }
        if (ghostA) A.end = B.end;     // a ghost follows the live lane
        if (ghostB) B.end = A.end;
        // ---- joint leading trim (not semantic: skipped cells are all-zero for both lanes; lazy, a block at a time)
        if ((z0.x | z0.y | z0.z | z0.w | z1.x | z1.y | z1.z | z1.w) == 0u && j0 + 4 <= min(emin, min(A.end, B.end))) {
            j0 += 4; p0 += 2 * S; pq0 += S;
        }
    }
}

}  // namespace bswk
