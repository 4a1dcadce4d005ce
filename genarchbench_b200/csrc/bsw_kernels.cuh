// CUDA kernels (sm_100a) for the bsw hot path: banded Smith-Waterman seed extension with affine
// gaps, band, z-drop, end bonus -- bit-exact per pair with the reference's getScores16
// (/root/reference/benchmarks/bsw/src/bandedSWA.cpp:2679-3365; semantics defined by
// scalarBandedSWA :132-253 plus the vector path's z-drop rule :1889-1902, band :2898-2919 and
// row budget :3035-3036,3130-3144).
//
// Design:
//  * bsw_short_kernel: one THREAD per pair. The H/E rows of the pair live in shared memory,
//    interleaved by thread (element k of thread t at [k*NT + t]) so every access is bank-conflict free
//    whatever column each thread is at. bsw_win_kernel is the same code over a sliding window of the
//    rows (long queries under a narrow band); bsw_long_kernel gives everything longer one WARP per
//    pair (tiles of 128 columns, max-plus warp scan of F, REDUX row decisions).
//  * the two 16-bit lanes of every DPX instruction are two ADJACENT COLUMNS (2g, 2g+1) of the same
//    row of the same pair, so all row-sequential decisions of the reference (band clamp, row max /
//    last argmax, m==0 exit, z-drop, trailing-zero trimming) stay exact and per pair.
//      M  = min(Hd + s, Hd * (match+1))  VIADDMNMX.S16x2 (== Hd ? Hd+s : 0 up to values <= 0, which
//                                        behave like 0); general form when the product may overflow
//      T  = max(M - oe, 0)               VIADDMNMX.S16x2.RELU
//      E' = max(E - e_del, T)            VIADDMNMX.S16x2
//      F  : two-step in-register scan    2 x VIADDMNMX.S16x2 (+ IMAD, IMAD.HI lane moves)
//      H  = max(M, E, F)                 VIMNMX3.S16x2
//    substitution scores for both lanes come from ONE PRMT that indexes an 8-byte LUT; the selector is
//    query seed + target seed (an IMAD) for plain pairs, a LOP3 for pairs holding an ambiguous base.
//  * packed sequences (2-bit, or 4-bit when a pair holds an ambiguous base) sit in 16-byte aligned
//    slots of the slab blob; each thread expands its query into selector seeds once per pair and
//    reads its target 16 rows at a time.
//  * bsw_key_kernel + cub radix sort bin the pairs by length on the device; extend_duo2
//    (bsw_duo.cuh) is an evaluated alternative with two pairs per thread.
// The BSW_* macros below are the A/B switches of the experiments recorded in DESIGN.md 5.6.
#pragma once
#include <stdint.h>
#ifndef BSW_HIER_ARGMAX
#define BSW_HIER_ARGMAX 0      // experiment (8-group trips): one argmax event per trip + a post-row scan
#endif
#ifndef BSW_SEL_LOP3      // experiment: LOP3 selector for narrow pairs too
#define BSW_SEL_LOP3 0
#endif
#ifndef BSW_TRIM_EVERY    // leading trim every N-th row (power of two)
#define BSW_TRIM_EVERY 1
#endif
#ifndef BSW_SCALAR_F      // F chain as two scalar VIADDMNMX per group (T lanes split on the FMA pipe)
#define BSW_SCALAR_F 0
#endif
#ifndef BSW_PREFETCH      // issue the next block's loads before computing the current block
#define BSW_PREFETCH 1
#endif
#ifndef BSW_WIN_GROUPS    // groups per inner-loop trip of the windowed kernel
#define BSW_WIN_GROUPS 8
#endif
#ifndef BSW_TAIL_NOUNROLL  // keep the single-group tail loop rolled (smaller code)
#define BSW_TAIL_NOUNROLL 1
#endif
#ifndef BSW_SHORT_GROUPS  // groups per inner-loop trip of the whole-row thread-per-pair kernel
#define BSW_SHORT_GROUPS 4
#endif
#ifndef BSW_KEY_REL        // keyed argmax: index field relative to the row's first group (would also key config 2's
#define BSW_KEY_REL 0      // 250-300-base queries: 34.44 vs 34.49 ms there, 7.22 vs 7.16 ms on config 3 -- off)
#endif
#ifndef BSW_PINGPONG_WIDE_TRIPS  // the same for trips of more than four groups: config 4 44.9 ms vs 41.5 ms (110 registers)
#define BSW_PINGPONG_WIDE_TRIPS 0
#endif
#ifndef BSW_PIN_CONSTS    // keep the lane-move multipliers in registers across the row loop
#define BSW_PIN_CONSTS 1
#endif
#ifndef BSW_PIN_ROW_CONSTS  // also pin the per-row scalars (kbits, zdrop, e_del: three LDCU per row): 7.16 vs 7.10 ms, off
#define BSW_PIN_ROW_CONSTS 0
#endif
#ifndef BSW_PINGPONG      // two copies of the four-group trip alternate between two register sets
#define BSW_PINGPONG 1
#endif
#ifndef BSW_HALF_TRIP      // a two-group step between the four-group trips and the single-group tail
#define BSW_HALF_TRIP 1
#endif
#ifndef BSW_ST_SHARED     // 16-bit row stores through st.shared (32-bit addresses) instead of generic stores
#define BSW_ST_SHARED 1
#endif
#ifdef BSW_HOST_EMUL
// tests/host_emul compiles the per-pair code below with g++ against an emulation of the few CUDA
// intrinsics it uses, so the algorithm can be checked against the oracle without a GPU.
#include "dpx_host_emul.h"
#else
#include <cuda_runtime.h>
#endif

namespace bswk {

#ifndef BSW_NT            // threads per block == pairs per block of the thread-per-pair kernel
#define BSW_NT 128
#endif
#ifndef BSW_HST_PRMT      // shifted H store through one PRMT (ALU pipe) instead of IMAD.HI + IMAD (FMA pipe)
#define BSW_HST_PRMT 1
#endif
constexpr int kBlockPairs = BSW_NT;
#ifndef BSW_WIN_NT        // threads per block of the windowed-rows kernel: its blocks are shared-memory bound, and
#define BSW_WIN_NT 32     // 41 KB blocks of one warp pack five warps on an SM where a 164 KB block packs four
#endif
constexpr int kWinBlockPairs = BSW_WIN_NT;

struct KParams {
    int o_del, e_del, o_ins, e_ins, zdrop, end_bonus, match, mismatch, ambig, w;
    // max(0, match, -mismatch, ambig), the reference wrapper's `max` (bandedSWA.cpp:2790-2793).
    // Computed on the HOST: ptxas 12.9 (sm_100a) fuses max(max(max(match, -mismatch), ambig), 0)
    // into one VIMNMX3.RELU and drops the negation (seen on B200: band came out as w).
    int max_score;
    // Multipliers read from the parameter bank at run time, so that ptxas keeps the lane shifts
    // below as IMAD / IMAD.HI on the FMA pipe instead of folding them into ALU-pipe shifts (the
    // ALU pipe is what bounds this kernel): k16 = 65536, km = match + 1, k1 = 1.
    uint32_t k16, km, k1;
    // extend_pair<.., KEY>: row argmax by key = score << kbits | group (set per launch, 0 otherwise)
    uint32_t kkey, kbits;
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
                      // bit1: the 4-bit blob sits AT off (packed input); otherwise the 2-bit slot at off holds,
                      //       in its first word, the word offset of the 4-bit copy
};
// start of the packed [query | target] a pair's kernel reads (2-bit for plain pairs, 4-bit for wide ones)
__device__ __forceinline__ const uint32_t *pair_blob(const uint32_t *__restrict__ blob, const PairMeta &m) {
    const uint32_t *src = blob + m.off;
    if ((m.flags & 3) == 1) src = blob + src[0];
    return src;
}

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
// Row storage. Everything is interleaved by thread so that a warp's accesses are conflict free
// whatever column each thread is at:
//   he4[k] : uint4 = the reference's eh[] for FOUR adjacent columns 4k .. 4k+3 (two "groups"):
//              .x = { Hs[4k],   Hs[4k+1] }   .y = { E[4k],   E[4k+1] }
//              .z = { Hs[4k+2], Hs[4k+3] }   .w = { E[4k+2], E[4k+3] }
//            with Hs[j] = H(i-1, j-1) (eh[j].h) and E[j] = E(i, j) (eh[j].e), int16 each;
//            element k of this thread lives at he4[k * stride]                      (LDS/STS.128)
//   qs[k]  : u32 = the PRMT selector seeds of query columns 4k .. 4k+3 (16 bits per group)
//   tb     : the pair's PACKED target in the slab blob (global memory, read 8 rows at a time)
//   kmask  : -1 for whole rows. A WINDOWED row (long query, narrow band: extend_pair<.., WIN>) keeps only
//            kmask + 1 elements -- element k lives in slot k & kmask -- because the live entries of a row
//            all lie within the band around the diagonal; see extend_pair.
//   qb     : the pair's packed query (windowed rows expand selector seeds as columns enter the window)
// ---------------------------------------------------------------------------------------------
struct Rows {
    uint4 *he4;
    uint32_t *qs;
    const uint32_t *tb;
    int stride;  // threads sharing the arrays (blockDim for shared memory, grid-wide for global)
    int kmask;
    const uint32_t *qb;
    __device__ __forceinline__ uint4 &HE4(int k) const { return he4[(size_t)(k & kmask) * stride]; }
    // group g = columns (2g, 2g+1): the .xy or .zw half of element g >> 1
    __device__ __forceinline__ uint2 &HE(int g) const {
        return reinterpret_cast<uint2 *>(he4 + (size_t)((g >> 1) & kmask) * stride)[g & 1];
    }
    // 16-bit views of the rows. They go through the SAME 32-bit words the packed loop reads and
    // writes (no differently-typed aliases the compiler could reorder around the packed accesses).
    __device__ __forceinline__ uint32_t getH16(int j) const {
        const uint32_t w = HE(j >> 1).x;
        return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
    }
    __device__ __forceinline__ uint32_t getE16(int j) const {
        const uint32_t w = HE(j >> 1).y;
        return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
    }
    // Hs[j] = hv, E[j] = ev with two 16-bit stores (no read-modify-write latency in front of the
    // row). Issued as asm with a memory clobber so the packed accesses are not moved across them.
    __device__ __forceinline__ void setHE16(int j, uint32_t hv, uint32_t ev) const {
#ifdef BSW_HOST_EMUL
        uint2 &p = HE(j >> 1);
        uint2 v = p;
        if (j & 1) { v.x = (v.x & 0xFFFFu) | (hv << 16); v.y = (v.y & 0xFFFFu) | (ev << 16); }
        else { v.x = (v.x & 0xFFFF0000u) | hv; v.y = (v.y & 0xFFFF0000u) | ev; }
        p = v;
#else
        unsigned char *p = reinterpret_cast<unsigned char *>(&HE(j >> 1)) + 2 * (j & 1);
#if BSW_ST_SHARED
        const uint32_t sp = (uint32_t)__cvta_generic_to_shared(p);
        asm volatile("st.shared.u16 [%0], %1;\n\tst.shared.u16 [%0+4], %2;" ::"r"(sp), "h"((unsigned short)hv),
                     "h"((unsigned short)ev) : "memory");
#else
        asm volatile("st.u16 [%0], %2;\n\tst.u16 [%1], %3;" ::"l"(p), "l"(p + 4), "h"((unsigned short)hv),
                     "h"((unsigned short)ev) : "memory");
#endif
#endif
    }
    __device__ __forceinline__ uint32_t &QS2(int k) const { return qs[(size_t)(k & kmask) * stride]; }
    __device__ __forceinline__ uint32_t QS(int g) const {   // 16-bit seed of one group
        const uint32_t w = QS2(g >> 1);
        return (g & 1) ? (w >> 16) : (w & 0xFFFFu);
    }
    // The row's target seed for the PRMT selector, in both 16-bit halves: nibbles c, c | 8 with
    // c = 4 - code (narrow pairs, see score_lut) or the code itself (wide pairs). `traw` caches the packed
    // target word (16 rows of 2 bits, or 8 rows of 4 bits) and is refilled when it runs out: one LDG per
    // 16 / 8 rows and three ALU ops per row.
    // tnext (initially tb[0]) holds the word after the current one (loaded a refill period ahead, so the global-load
    // latency is off the row's critical path; it may be the word past the target: slots are padded).
    template <bool WIDE>
    __device__ __forceinline__ uint32_t row_seed(int i, uint32_t &traw, uint32_t &tnext) const {
        uint32_t code;
        if (WIDE) {
            if ((i & 7) == 0) { traw = tnext; tnext = tb[(i >> 3) + 1]; }
            code = traw & 7u;
            traw >>= 4;
        } else {
            if ((i & 15) == 0) { traw = tnext; tnext = tb[(i >> 4) + 1]; }
            code = BSW_SEL_LOP3 ? (traw & 3u) : 4u - (traw & 3u);
            traw >>= 2;
        }
        return code * 0x11111111u + 0x80808080u;
    }
};

// dst = p ? g * k1 + c : dst   with k1 == 1 read from the parameter bank: a PREDICATED IMAD on the FMA
// pipe instead of the SEL the compiler emits for `if (p) dst = g + c` (the ALU pipe bounds the kernel).
#ifndef BSW_SEL_IMAD
#define BSW_SEL_IMAD 0   // measured: ptxas if-converts it back into IMAD + SEL
#endif
template <int C>
__device__ __forceinline__ void mov_if(int &dst, bool p, int g, uint32_t k1) {
#if defined(BSW_HOST_EMUL) || !BSW_SEL_IMAD
    (void)k1;
    if (p) dst = g + C;
#else
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q mad.lo.s32 %0, %2, %3, %4;\n\t}"
        : "+r"(dst) : "r"((unsigned)p), "r"(g), "r"((int)k1), "n"(C));
#endif
}

// PTX prmt.b32 (default mode): byte i of the result = byte (nibble_i & 7) of {b:a}; nibble bit 3 set
// => that byte's SIGN replicated instead. Only the low 16 bits of the selector are used.
// (__byte_perm() masks the selector with 0x7777 and loses the sign mode, so the instruction is
// issued directly.)
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

// Substitution scores come from ONE PRMT per group: an 8-byte look-up table indexed, per 16-bit
// lane, by a 3-bit code in the selector nibbles (value byte: the code; sign byte: code | 8).
//   narrow pairs (bases 0..3): code = q + (4 - t) in 1..7, == 4 iff q == t. The selector is a plain
//     ADD of the query seed (nibbles q) and the row's target seed (nibbles 4 - t, sign nibbles + 8):
//     no nibble ever carries, and the add is issued as IMAD on the FMA pipe (see KParams::k1).
//     LUT: byte 4 = match, every other byte = -mismatch.
//   wide pairs (a base may be 4 = ambiguous): code = (q ^ t) on bits 0-1 | (q | t) on bit 2, one LOP3.
//     LUT: byte 0 = match, 1..3 = -mismatch, 4..7 = ambiguous.
template <bool WIDE>
__device__ __forceinline__ void score_lut(const KParams &P, uint32_t &lo, uint32_t &hi) {
    const uint32_t mm = (uint32_t)(-P.mismatch) & 0xFFu, ma = (uint32_t)P.match & 0xFFu;
    if (WIDE || BSW_SEL_LOP3) {
        lo = ma | (mm * 0x01010100u);
        hi = ((uint32_t)P.ambig & 0xFFu) * 0x01010101u;
    } else {
        lo = mm * 0x01010101u;
        hi = ma | (mm * 0x01010100u);
    }
}

// selector seeds of query columns 4k .. 4k+3 (element k) from a packed query: a base b becomes the
// byte b * 0x11 (value nibble and sign nibble of one PRMT lane)
template <bool WIDE>
__device__ __forceinline__ uint32_t sel_word(const uint32_t *qb, int k) {
    uint32_t v;
    if (!WIDE) {
        v = (qb[k >> 2] >> (8 * (k & 3))) & 0xFFu;      // 4 bases, 2 bits each
        v = (v | (v << 12)) & 0x000F000Fu;
        v = (v | (v << 6)) & 0x03030303u;
    } else {
        v = (qb[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;   // 4 bases, 4 bits each
        v = (v | (v << 8)) & 0x00FF00FFu;
        v = (v | (v << 4)) & 0x0F0F0F0Fu;
    }
    return v * 0x11u;
}

// Expands the query of this thread's packed blob (4-byte words: query then target, each padded to 4
// bytes) into qs[] and points R.tb at the target. Narrow blobs hold 2 bits per base, wide blobs 4.
template <bool WIDE>
__device__ inline void unpack_pair(const uint32_t *blob, int qlen, Rows &R) {
    R.tb = blob + (seq_bytes((uint32_t)qlen, WIDE) >> 2);
    R.qb = blob;
    int nsel = (((qlen + 1) >> 1) + 1) >> 1;   // selector words: two groups each (== sel_words)
    if (R.kmask >= 0 && nsel > R.kmask + 1) nsel = R.kmask + 1;   // windowed rows: the first window only
    if (!WIDE) {
        // 16 bases per word -> 8 groups -> 4 selector words; a base b becomes the byte b * 0x11
        for (int w = 0, k = 0; k < nsel; ++w) {
            const uint32_t x = blob[w];
#pragma unroll
            for (int u = 0; u < 4; ++u, ++k) {
                if (k < nsel) {
                    uint32_t v = (x >> (8 * u)) & 0xFFu;          // 4 bases
                    v = (v | (v << 12)) & 0x000F000Fu;
                    v = (v | (v << 6)) & 0x03030303u;             // one base per byte
                    R.QS2(k) = v * 0x11u;
                }
            }
        }
    } else {
        // 8 bases per word -> 4 groups -> 2 selector words
        for (int w = 0, k = 0; k < nsel; ++w) {
            const uint32_t x = blob[w];
#pragma unroll
            for (int u = 0; u < 2; ++u, ++k) {
                if (k < nsel) {
                    uint32_t v = (x >> (16 * u)) & 0xFFFFu;       // 4 bases, 4 bits each
                    v = (v | (v << 8)) & 0x00FF00FFu;
                    v = (v | (v << 4)) & 0x0F0F0F0Fu;
                    R.QS2(k) = v * 0x11u;
                }
            }
        }
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

// bits needed to hold v (at least 1)
__host__ __device__ inline int bits_for(uint32_t v) { int b = 1; while ((v >> b) != 0u) ++b; return b; }
// number of he4 elements (4 columns each) a pair with qlen query bases needs: columns 0 .. qlen
// (one spare so that the hi lane of the last group is always initialised), rounded up
__host__ __device__ inline int row_elems(int qlen) { return (qlen + 4) >> 2; }
// elements of a windowed row (extend_pair<.., WIN>) for band width w: a power of two, 4 * nk >= 2 w + 16
__host__ __device__ inline int window_elems(int w) {
    if (w > 65536) w = 65536;   // a pair's band never exceeds its query length; keeps the arithmetic in range
    int nk = 8;
    while (4 * nk < 2 * w + 16) nk <<= 1;
    return nk;
}
// number of qs words (selector seeds of two groups = 4 columns each) for qlen query bases
__host__ __device__ inline int sel_words(int qlen) { return (((qlen + 1) >> 1) + 1) >> 1; }

// The DP of one pair over row storage R (already holding the selector seeds qs[]; R.tb = packed target).
//   FASTM : every score of the launch satisfies score * (match + 1) <= 32767, so the reference's
//           M = Hd ? Hd + s : 0 is ONE instruction, min(Hd + s, Hd * (match + 1)) (the product on the
//           FMA pipe): for Hd >= 1 the second term is >= Hd + match >= Hd + s, for Hd == 0 it caps M
//           at 0 -- and any M <= 0 behaves like 0 in max(M, E, F) and in max(M - oe, 0).
//   SYM   : o_del == o_ins && e_del == e_ins (one T for both gap kinds)
//   COUNT : also track the reference's exact leading trim and count the cells its scalar loop would
//           visit (bandedSWA.cpp:191-216; the commented SW_cells++ at :215) -- the unit of work of
//           the GCUPS metric. Used once per input outside any timed region.
//   WIDE  : the pair may hold ambiguous bases (4-bit blob, LOP3 selector; see score_lut)
//   WIN   : windowed rows for long queries under a narrow band. Row i only touches entries
//           [beg, end] with i - band <= beg and end <= i + band + 1, everything left of beg is dead,
//           and an entry right of every `end` seen so far still holds its row "-1" value, which has a
//           closed form. So R keeps kmask + 1 elements (4 * (kmask + 1) >= 2 * band + 16 columns, slot =
//           element & kmask): before a row, the elements that `end` newly reaches are (re)initialised
//           -- row "-1" values, zero E, selector seeds from the packed query -- over slots whose old
//           columns have fallen out of the band. A 1000-base query under w = 100 then needs 1.3 KB of
//           shared memory instead of 5 KB and stays on the thread-per-pair kernel.
//   NB    : groups per trip of the inner loop, 4 or 8. Eight help the windowed launches, which run at one
//           warp per scheduler and have nothing else to hide latency with (config 4: 81 -> 73 ms), and
//           cost the whole-row launches 8 % (more registers, a longer single-group tail).
//   KEY   : every score of the launch is < 2^(16 - kbits) and every group index < 2^kbits (P.kkey =
//           1 << P.kbits): the row's "last column reaching the maximum" (bandedSWA.cpp:204-205) is then
//           the unsigned lane maximum of score << kbits | group -- per trip four IMADs on the FMA pipe
//           and three ALU-pipe maxima instead of 4 x (VIMNMX with predicates + 2 SEL + index add).
#ifndef BSW_HOST_EMUL
__device__ uint32_t g_zero = 0;   // see BSW_PIN_CONSTS
#endif
template <bool FASTM, bool SYM, bool COUNT, bool WIDE, bool WIN = false, int NB = 4, bool KEY = false>
__device__ inline PairResult extend_pair(const Rows &R, int qlen, int tlen, int h0, const KParams &P) {
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const uint32_t NEG_OE_DEL = pack2(-oe_del), NEG_OE_INS = pack2(-oe_ins);
    const uint32_t NEG_E_DEL = pack2(-P.e_del);
    const int NEG_E_INS_S = -P.e_ins;
    (void)NEG_E_INS_S;   // BSW_SCALAR_F only
    const uint32_t NEG_E_INS = pack2(-P.e_ins);
    uint32_t LUT_LO, LUT_HI;
    score_lut<WIDE>(P, LUT_LO, LUT_HI);

    // row "-1" (bandedSWA.cpp:159-161) and zeroed E of element k (columns 4k .. 4k+3):
    // Hs[0] = h0, Hs[j] = max(h0 - oe_ins - (j-1) e_ins, 0) for 1 <= j <= qlen, 0 beyond (calloc'ed tail)
    auto init_elem = [&](int k) {
        uint32_t hv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = 4 * k + u;
            int v = j == 0 ? h0 : max(h0 - oe_ins - (j - 1) * P.e_ins, 0);
            if (j > qlen) v = 0;
            hv[u] = (uint32_t)v;
        }
        uint4 w; w.x = hv[0] | (hv[1] << 16); w.y = 0u; w.z = hv[2] | (hv[3] << 16); w.w = 0u;
        R.HE4(k) = w;
    };
    int kinit = row_elems(qlen);            // elements [0, kinit) hold valid entries
    if (WIN) kinit = min(kinit, R.kmask + 1);
    for (int k = 0; k < kinit; ++k) init_elem(k);

    const int band = pair_band(P, qlen);
    const int budget = min(qlen + band, tlen);
    static_assert(!KEY || (!WIN && !COUNT), "keyed argmax: whole rows only");
    uint32_t K16 = P.k16, KM = P.km, K1 = P.k1, KK = P.kkey;
    uint32_t KBITS = P.kbits;          // per-row scalars, pinned the same way (BSW_PIN_ROW_CONSTS)
    int ZDROP = P.zdrop, EDEL = P.e_del;
#if !defined(BSW_HOST_EMUL) && BSW_PIN_CONSTS
    // made opaque by a run-time zero from global memory: ptxas otherwise re-loads all four from the
    // parameter bank inside every trip (three LDC per trip)
    {
        const uint32_t z = *reinterpret_cast<const volatile uint32_t *>(&g_zero);
        K16 ^= z; KM ^= z; K1 ^= z; KK ^= z;
#if BSW_PIN_ROW_CONSTS
        KBITS ^= z; ZDROP ^= (int)z; EDEL ^= (int)z;
#endif
    }
#endif

    int best = h0, best_i = -1, best_j = -1, g_i = -1, gsc = -1, off = 0;
    int beg = 0, end = qlen;
    uint32_t tword = 0, tnext = R.tb[0];
    int hcol = h0 - P.o_del;  // first column: H(i,-1) = max(h0 - o_del - e_del*(i+1), 0)
    int xbeg = 0;             // COUNT: the reference's exact beg (ours lags it by up to 3 columns)
    uint32_t cells = 0;

    for (int i = 0; i < budget; ++i) {
        if (beg < i - band) beg = i - band;
        if (end > i + band + 1) end = i + band + 1;
        if (beg >= end) break;
        if (COUNT) {
            if (xbeg < i - band) xbeg = i - band;
            cells += (uint32_t)(end - xbeg);
        }

        if (WIN) {
            // elements that `end` reaches for the first time enter the window
            const int kneed = min(end, qlen) >> 2;
            while (kinit <= kneed) {
                init_elem(kinit);
                R.QS2(kinit) = sel_word<WIDE>(R.qb, kinit);
                ++kinit;
            }
        }
        const uint32_t tsel = R.template row_seed<WIDE>(i, tword, tnext);

        hcol -= EDEL;
        const int hleft = beg == 0 ? max(hcol, 0) : 0;

        // The row is computed over whole groups starting at a 4-column boundary. Lanes outside
        // [beg, end) must see zero inputs: instead of masking inside the loop, the stale (never read
        // again) entry just left of beg is cleared -- everything further left, down to the boundary, is
        // already zero (cleared by earlier rows while the band clamp moved beg one column per row, or
        // zero when the leading trim moved beg, which it only does in steps of 4) -- and so is the
        // entry at `end` when it shares a group with column end - 1.
        if (beg & 3) R.setHE16(beg - 1, 0u, 0u);
        if (end & 1) R.setHE16(end, 0u, 0u);

        const int g0 = (beg >> 2) << 1, g1 = (end - 1) >> 1;
        uint32_t hprev = (uint32_t)hleft << 16;  // .hi = H(i, 2*g0 - 1)
        int F = 0;                               // F(i, 2g), a plain int: the only serial chain of the row
        uint32_t rm = 0;                         // running max per lane (even / odd columns)
        int ilo = g0, ihi = g0;                  // last group where a lane reached rm
        const int KEY_G0 = BSW_KEY_REL ? g0 : 0;   // base of the key's index field
        constexpr bool HIER = BSW_HIER_ARGMAX && NB >= 8;   // measured slower: DESIGN.md 5.6
        uint32_t h = 0, En = 0, Hst = 0;

        // One group = columns (2g, 2g+1). The scores, M, T and E' of different groups are independent;
        // only F runs along the row, as TWO dependent scalar VIADDMNMX per group (the T lanes are split
        // off the chain on the FMA pipe). Groups are processed four at a time; the loads of the next
        // four (2 x LDS.128 + 2 x LDS.32) are issued before the current four are computed.
        auto front = [&](const uint32_t Hd, const uint32_t Ev, const uint32_t sel, uint32_t &M, uint32_t &Tins,
                         uint32_t &Enew) {
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
        // F scan of one group and its H; returns the word { H(i,2g-1), H(i,2g) } to store
        auto back = [&](const uint32_t Ev, const uint32_t M, const uint32_t Tins) -> uint32_t {
#if BSW_SCALAR_F
            const int tlo = (int)__umulhi(Tins * K16, K16);              // T(2g)   (T >= 0; two IMADs)
            const int thi = (int)__umulhi(Tins, K16);                    // T(2g+1)
            const int F1 = __viaddmax_s32(F, NEG_E_INS_S, tlo);          // F(i, 2g+1)
            const uint32_t B = (uint32_t)F1 * K16 + (uint32_t)F;         // { F(2g), F(2g+1) }  (IMAD)
            F = __viaddmax_s32(F1, NEG_E_INS_S, thi);                    // F(i, 2g+2)
#else
            const uint32_t A = (uint32_t)F;                              // { F(2g), 0 }
            const uint32_t W1 = __viaddmax_s16x2(A, NEG_E_INS, Tins);   // .lo = F(i, 2g+1)
            const uint32_t B = W1 * K16 + A;                             // { F(2g), F(2g+1) }  (IMAD)
            const uint32_t W2 = __viaddmax_s16x2(B, NEG_E_INS, Tins);   // .hi = F(i, 2g+2)
            F = (int)__umulhi(W2, K16);                                  // W2 >> 16           (IMAD.HI)
#endif
            h = __vimax3_s16x2(M, Ev, B);
#if BSW_HST_PRMT
            const uint32_t st = __byte_perm(hprev, h, 0x5432);           // { H(i,2g-1), H(i,2g) }
#else
            const uint32_t st = __umulhi(hprev, K16) + h * K16;          // { H(i,2g-1), H(i,2g) }
#endif
            hprev = h;
            return st;
        };
        int g = g0;
        if (NB == 4) {
            if (g + 3 <= g1) {
                // one trip over the elements (a, b); the next trip's elements are loaded into (na, nb)
                // before the arithmetic. Two copies alternate between two register sets (BSW_PINGPONG), so
                // no moves rotate the prefetched values.
                auto trip = [&](const uint4 &a, const uint4 &b, const uint32_t q01, const uint32_t q23, uint4 &na,
                                uint4 &nb, uint32_t &nq01, uint32_t &nq23) -> bool {
                    const int k = g >> 1;
                    const bool more = g + 7 <= g1;
                    if (BSW_PREFETCH && more) {
                        na = R.HE4(k + 2); nb = R.HE4(k + 3);
                        nq01 = R.QS2(k + 2); nq23 = R.QS2(k + 3);
                    }
                    uint32_t s0, s1, s2, s3;
                    if (WIDE || BSW_SEL_LOP3) {
                        s0 = sel_combine(q01, tsel, 0x44444444u); s1 = __umulhi(s0, K16);
                        s2 = sel_combine(q23, tsel, 0x44444444u); s3 = __umulhi(s2, K16);
                    } else {
                        // tsel carries the row's seed in both halves: the upper half of the sum is group 1's selector
                        s0 = q01 * K1 + tsel; s1 = __umulhi(s0, K16);
                        s2 = q23 * K1 + tsel; s3 = __umulhi(s2, K16);
                    }
                    uint32_t M0, M1, M2, M3, T0, T1, T2, T3, E0, E1, E2, E3;
                    front(a.x, a.y, s0, M0, T0, E0);
                    front(a.z, a.w, s1, M1, T1, E1);
                    front(b.x, b.y, s2, M2, T2, E2);
                    front(b.z, b.w, s3, M3, T3, E3);
                    uint4 oa, ob;
                    oa.x = back(a.y, M0, T0); const uint32_t h0v = h;
                    oa.z = back(a.w, M1, T1); const uint32_t h1v = h;
                    ob.x = back(b.y, M2, T2); const uint32_t h2v = h;
                    ob.z = back(b.w, M3, T3);
                    oa.y = E0; oa.w = E1; ob.y = E2; ob.w = E3;
                    R.HE4(k) = oa;
                    R.HE4(k + 1) = ob;
                    Hst = ob.z; En = E3;
                    if (KEY) {
                        // the later group wins ties, as `h >= m` does in the reference
                        const uint32_t t3 = __vimax3_u16x2(h0v * KK, h1v * KK + 0x00010001u, h2v * KK + 0x00020002u);
                        const uint32_t t4 = __vmaxu2(t3, h * KK + 0x00030003u);
                        rm = __viaddmax_u16x2(t4, (uint32_t)(g - KEY_G0) * 0x00010001u, rm);
                    } else {
                        bool phi, plo;
                        rm = __vibmax_s16x2(h0v, rm, &phi, &plo); mov_if<0>(ilo, plo, g, K1); mov_if<0>(ihi, phi, g, K1);
                        rm = __vibmax_s16x2(h1v, rm, &phi, &plo); mov_if<1>(ilo, plo, g, K1); mov_if<1>(ihi, phi, g, K1);
                        rm = __vibmax_s16x2(h2v, rm, &phi, &plo); mov_if<2>(ilo, plo, g, K1); mov_if<2>(ihi, phi, g, K1);
                        rm = __vibmax_s16x2(h, rm, &phi, &plo);   mov_if<3>(ilo, plo, g, K1); mov_if<3>(ihi, phi, g, K1);
                    }
                    if (!BSW_PREFETCH && more) {
                        na = R.HE4(k + 2); nb = R.HE4(k + 3);
                        nq01 = R.QS2(k + 2); nq23 = R.QS2(k + 3);
                    }
                    g += 4;
                    return more;
                };
                uint4 a0 = R.HE4(g >> 1), b0 = R.HE4((g >> 1) + 1);
                uint32_t q0 = R.QS2(g >> 1), r0 = R.QS2((g >> 1) + 1);
#if BSW_PINGPONG
                uint4 a1, b1;           // written by the first trip before the second reads them
                uint32_t q1, r1;
                for (;;) {
                    if (!trip(a0, b0, q0, r0, a1, b1, q1, r1)) break;
                    if (!trip(a1, b1, q1, r1, a0, b0, q0, r0)) break;
                }
#else
                bool more;
                do {
                    uint4 na = a0, nb = b0;
                    uint32_t nq = q0, nr = r0;
                    more = trip(a0, b0, q0, r0, na, nb, nq, nr);
                    a0 = na; b0 = nb; q0 = nq; r0 = nr;
                } while (more);
#endif
            }
        } else {
            // NB groups (NB / 2 elements) per trip
            constexpr int NE = NB / 2;
            if (g + NB - 1 <= g1) {
                auto trip = [&](const uint4 (&cur)[NE], const uint32_t (&cq)[NE], uint4 (&nxt)[NE],
                                uint32_t (&nq)[NE]) -> bool {
                    const int k = g >> 1;
                    const bool more = g + 2 * NB - 1 <= g1;
                    if (BSW_PREFETCH && more) {
    #pragma unroll
                        for (int e = 0; e < NE; ++e) { nxt[e] = R.HE4(k + NE + e); nq[e] = R.QS2(k + NE + e); }
                    }
                    uint32_t M[NB], T[NB], E[NB], hv[NB];
    #pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        uint32_t s0, s1;
                        if (WIDE || BSW_SEL_LOP3) { s0 = sel_combine(cq[e], tsel, 0x44444444u); s1 = __umulhi(s0, K16); }
                        else { s0 = cq[e] * K1 + tsel; s1 = __umulhi(s0, K16); }   // tsel: the row's seed in both halves
                        front(cur[e].x, cur[e].y, s0, M[2 * e], T[2 * e], E[2 * e]);
                        front(cur[e].z, cur[e].w, s1, M[2 * e + 1], T[2 * e + 1], E[2 * e + 1]);
                    }
    #pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        uint4 o;
                        o.x = back(cur[e].y, M[2 * e], T[2 * e]); hv[2 * e] = h;
                        o.z = back(cur[e].w, M[2 * e + 1], T[2 * e + 1]); hv[2 * e + 1] = h;
                        o.y = E[2 * e]; o.w = E[2 * e + 1];
                        R.HE4(k + e) = o;
                        if (e == NE - 1) { Hst = o.z; En = o.w; }
                    }
                    if (KEY) {
                        uint32_t t = hv[0] * KK;
    #pragma unroll
                        for (int u = 1; u + 1 < NB; u += 2)
                            t = __vimax3_u16x2(t, hv[u] * KK + (uint32_t)u * 0x00010001u,
                                               hv[u + 1] * KK + (uint32_t)(u + 1) * 0x00010001u);
                        if (!(NB & 1)) t = __vmaxu2(t, hv[NB - 1] * KK + (uint32_t)(NB - 1) * 0x00010001u);
                        rm = __viaddmax_u16x2(t, (uint32_t)(g - KEY_G0) * 0x00010001u, rm);
                    } else if (HIER) {
                        // one >= event per trip: the post-row scan finds the group inside it
                        uint32_t tm = hv[0];
    #pragma unroll
                        for (int u = 1; u + 1 < NB; u += 2) tm = __vimax3_s16x2(tm, hv[u], hv[u + 1]);
                        if (!(NB & 1)) tm = __vmaxs2(tm, hv[NB - 1]);
                        bool phi, plo;
                        rm = __vibmax_s16x2(tm, rm, &phi, &plo);
                        if (plo) ilo = g + NB - 1;
                        if (phi) ihi = g + NB - 1;
                    } else {
    #pragma unroll
                        for (int u = 0; u < NB; ++u) {
                            bool phi, plo;
                            rm = __vibmax_s16x2(hv[u], rm, &phi, &plo);
                            if (plo) ilo = g + u;
                            if (phi) ihi = g + u;
                        }
                    }
                    if (!BSW_PREFETCH && more) {
    #pragma unroll
                        for (int e = 0; e < NE; ++e) { nxt[e] = R.HE4(k + NE + e); nq[e] = R.QS2(k + NE + e); }
                    }
                    g += NB;
                    return more;
                };
                uint4 c0[NE];
                uint32_t d0[NE];
    #pragma unroll
                for (int e = 0; e < NE; ++e) { c0[e] = R.HE4((g >> 1) + e); d0[e] = R.QS2((g >> 1) + e); }
#if BSW_PINGPONG_WIDE_TRIPS
                uint4 c1[NE];
                uint32_t d1[NE];
                for (;;) {
                    if (!trip(c0, d0, c1, d1)) break;
                    if (!trip(c1, d1, c0, d0)) break;
                }
#else
                bool more;
                do {
                    uint4 nxt[NE];
                    uint32_t nq[NE];
    #pragma unroll
                    for (int e = 0; e < NE; ++e) { nxt[e] = c0[e]; nq[e] = d0[e]; }
                    more = trip(c0, d0, nxt, nq);
    #pragma unroll
                    for (int e = 0; e < NE; ++e) { c0[e] = nxt[e]; d0[e] = nq[e]; }
                } while (more);
#endif
            }
        }
#if BSW_HALF_TRIP
        // whole elements (two groups) of what the trips left, so that at most one group goes through
        // the (per group much dearer) single-group loop below
#pragma unroll 1
        for (; g + 1 <= g1; g += 2) {
            const int k = g >> 1;
            const uint4 a = R.HE4(k);
            const uint32_t q01 = R.QS2(k);
            const uint32_t s0 = (WIDE || BSW_SEL_LOP3) ? sel_combine(q01, tsel, 0x44444444u) : q01 * K1 + tsel;
            const uint32_t s1 = __umulhi(s0, K16);
            uint32_t M0, M1, T0, T1, E0, E1;
            front(a.x, a.y, s0, M0, T0, E0);
            front(a.z, a.w, s1, M1, T1, E1);
            uint4 oa;
            oa.x = back(a.y, M0, T0); const uint32_t h0v = h;
            oa.z = back(a.w, M1, T1);
            oa.y = E0; oa.w = E1;
            R.HE4(k) = oa;
            Hst = oa.z; En = E1;
            if (KEY) {
                const uint32_t t2 = __vmaxu2(h0v * KK, h * KK + 0x00010001u);
                rm = __viaddmax_u16x2(t2, (uint32_t)(g - KEY_G0) * 0x00010001u, rm);
            } else {
                bool phi, plo;
                rm = __vibmax_s16x2(h0v, rm, &phi, &plo); if (plo) ilo = g; if (phi) ihi = g;
                rm = __vibmax_s16x2(h, rm, &phi, &plo); if (plo) ilo = g + 1; if (phi) ihi = g + 1;
            }
        }
#endif
#if BSW_TAIL_NOUNROLL
#pragma unroll 1
#endif
        for (; g <= g1; ++g) {
            const uint2 he0 = R.HE(g);
            const uint32_t q = R.QS(g);
            const uint32_t s0 = (WIDE || BSW_SEL_LOP3) ? sel_combine(q, tsel, 0x44444444u) : q * K1 + tsel;
            uint32_t M0, T0, E0;
            front(he0.x, he0.y, s0, M0, T0, E0);
            Hst = back(he0.y, M0, T0);
            En = E0;
            R.HE(g) = make_uint2(Hst, En);
            if (KEY) {
                rm = __vmaxu2(rm, h * KK + (uint32_t)(g - KEY_G0) * 0x00010001u);
            } else {
                bool phi, plo;
                rm = __vibmax_s16x2(h, rm, &phi, &plo);
                if (plo) ilo = g;
                if (phi) ihi = g;
            }
        }

        // first element of the row for the leading trim below, loaded here so that its latency hides behind
        // the row decisions (the store of entry `end` that follows is why the trim then skips an element
        // holding that entry)
        const uint4 ztrim = R.HE4(g0 >> 1);
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
        int mlo = 0, mhi = 0, m, mj = 0;
        if (KEY) {
            // the larger key is the larger score; at equal scores the larger group index, and at equal
            // indices the odd column -- exactly "the last column reaching the maximum"
            const uint32_t klo = rm & 0xFFFFu, khi = rm >> 16;
            const bool odd = khi >= klo;
            const uint32_t kb = max(klo, khi);
            m = (int)(kb >> KBITS);
            mj = 2 * (KEY_G0 + (int)(kb & (KK - 1u))) + (odd ? 1 : 0);
        } else {
            mlo = (int)(short)(rm & 0xFFFFu); mhi = (int)(short)(rm >> 16);
            m = max(mlo, mhi);
        }
        if (m == 0) break;
        // LAST column reaching m (the keyed path has it already)
        if (!KEY && HIER) {
            // H(i, j) now sits in Hs[j + 1]; the lane's last >= event happened in the (at most NB)
            // groups ending at ilo / ihi, so the scan below stops within them.
            // (the hi lane of the last group is column `end` when end is odd: never a candidate)
            mj = -1;
            if (mlo >= mhi) {
                int gg = ilo;
#pragma unroll 1
                for (int t = 0; t < NB - 1 && R.getH16(2 * gg + 1) != (uint32_t)mlo; ++t) --gg;
                mj = 2 * gg;
            }
            if (mhi >= mlo) {
                int gg = min(ihi, (end - 2) >> 1);
#pragma unroll 1
                for (int t = 0; t < NB - 1 && R.getH16(2 * gg + 2) != (uint32_t)mhi; ++t) --gg;
                mj = max(mj, 2 * gg + 1);
            }
        } else if (!KEY) {
            const int jlo = 2 * ilo, jhi = 2 * ihi + 1;
            mj = mlo > mhi ? jlo : (mhi > mlo ? jhi : max(jlo, jhi));
        }
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            off = max(off, abs(mj - i));
        } else {
            // vector z-drop rule: no gap-extend factor, no zdrop > 0 guard (bandedSWA.cpp:1889-1902)
            const int di = i - best_i, dj = mj - best_j;
            if (best - m - abs(di - dj) > ZDROP) break;
        }

        if (COUNT) {   // the reference's scan (bandedSWA.cpp:234-235), on the rows just written
            int j = xbeg;
            while (j < end && R.getH16(j) == 0 && R.getE16(j) == 0) ++j;
            xbeg = j;
        }
        // leading trim (not semantic: skipped cells are all-zero; done lazily, four columns at a time)
        if ((i & (BSW_TRIM_EVERY - 1)) == BSW_TRIM_EVERY - 1) {
            const uint4 z = ztrim;
            if ((z.x | z.y | z.z | z.w) == 0u && 2 * g0 + 4 > beg && 2 * g0 + 4 <= end) beg = 2 * g0 + 4;
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
                int gz = g1 - 1;
                uint32_t wz = 0;
                const int glow = WIN ? (beg >> 1) : 0;   // everything left of beg is zero (and, windowed, gone)
                for (; gz >= glow; --gz) {
                    const uint2 z = R.HE(gz);
                    wz = z.x | z.y;
                    if (wz) break;
                }
                jstar = (wz >> 16) ? 2 * gz + 1 : 2 * gz;
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

}  // namespace bswk
#include "bsw_duo.cuh"
namespace bswk {

#ifndef BSW_HOST_EMUL
__device__ __forceinline__ void store_result(PairOut *out, uint32_t id, const PairResult &r) {
    union { PairOut o; uint4 v; } u;
    u.o.score = (int16_t)r.score; u.o.qle = (int16_t)r.qle; u.o.tle = (int16_t)r.tle;
    u.o.gtle = (int16_t)r.gtle; u.o.gscore = (int16_t)r.gscore; u.o.max_off = (int16_t)r.max_off;
    u.o.cells = r.cells;
    reinterpret_cast<uint4 *>(out)[id] = u.v;
}

// ---------------------------------------------------------------------------------------------
// Short pairs: rows in shared memory. Launch: block = kBlockPairs threads, dynamic smem =
// (16*row_el + 4*qs_words) * NT.
//   meta[]: the slab's pairs in the caller's order; ord[0 .. n_wide + n_narrow): this launch's pairs
//   (indices into meta) in the device-sorted, length-binned order, the n_wide pairs holding an
//   ambiguous base first. Threads [0, roundup32(n_wide)) take the wide pairs, the threads
//   after them the narrow ones, so every WARP runs one instantiation of the DP (no divergence between
//   the LOP3-selector and the add-selector code). blob = the slab's packed sequences: each thread
//   expands its own 16-byte aligned slot straight from global memory (a few dozen bytes per pair, read
//   once; the host packs slots in the caller's order so that its own pass is a pure stream -- see
//   DESIGN.md). A wide pair's slot holds the word offset of its 4-bit blob in the slab's overflow area.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int launch_threads(int n_wide, int n_narrow) { return ((n_wide + 31) & ~31) + n_narrow; }

template <bool FASTM, bool SYM, bool COUNT, bool KEY = false>
__global__ void __launch_bounds__(kBlockPairs)
bsw_short_kernel(const PairMeta *__restrict__ meta, const uint32_t *__restrict__ ord,
                 const uint32_t *__restrict__ blob, PairOut *__restrict__ out, int n_wide, int n_narrow,
                 KParams P, int row_el, int qs_words) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = kBlockPairs;
    const int tid = threadIdx.x;
    const int t = blockIdx.x * NT + tid;
    const int nwr = (n_wide + 31) & ~31;
    const bool wide = t < nwr;                       // warp-uniform
    const int k = wide ? t : t - nwr + n_wide;
    if (wide ? t >= n_wide : k >= n_wide + n_narrow) return;
    const PairMeta m = meta[ord[k]];

    Rows R;
    R.stride = NT;
    R.he4 = reinterpret_cast<uint4 *>(smem) + tid;
    R.qs = reinterpret_cast<uint32_t *>(smem + (size_t)16 * row_el * NT) + tid;
    R.kmask = -1; R.qb = nullptr;
    (void)qs_words;

    const uint32_t *src = pair_blob(blob, m);
    PairResult r;
    if (wide) {
        unpack_pair<true>(src, m.len2, R);
        r = extend_pair<FASTM, SYM, COUNT, true, false, BSW_SHORT_GROUPS, KEY>(R, m.len2, m.len1, m.h0, P);
    } else {
        unpack_pair<false>(src, m.len2, R);
        r = extend_pair<FASTM, SYM, COUNT, false, false, BSW_SHORT_GROUPS, KEY>(R, m.len2, m.len1, m.h0, P);
    }
    store_result(out, m.id, r);
}

// ---------------------------------------------------------------------------------------------
// Long queries under a narrow band: the same thread-per-pair code over WINDOWED rows (extend_pair<.., WIN>).
// Launch as bsw_short_kernel but with kWinBlockPairs threads, dynamic smem = 20 * nk * kWinBlockPairs with nk = window elements (a power of
// two, 4 * nk >= 2 * w + 16).
// ---------------------------------------------------------------------------------------------
template <bool FASTM, bool SYM, bool COUNT>
__global__ void __launch_bounds__(kWinBlockPairs)
bsw_win_kernel(const PairMeta *__restrict__ meta, const uint32_t *__restrict__ ord,
               const uint32_t *__restrict__ blob, PairOut *__restrict__ out, int n_wide, int n_narrow,
               KParams P, int nk) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = kWinBlockPairs;
    const int tid = threadIdx.x;
    const int t = blockIdx.x * NT + tid;
    const int nwr = (n_wide + 31) & ~31;
    const bool wide = t < nwr;                       // warp-uniform
    const int k = wide ? t : t - nwr + n_wide;
    if (wide ? t >= n_wide : k >= n_wide + n_narrow) return;
    const PairMeta m = meta[ord[k]];

    Rows R;
    R.stride = NT;
    R.he4 = reinterpret_cast<uint4 *>(smem) + tid;
    R.qs = reinterpret_cast<uint32_t *>(smem + (size_t)16 * nk * NT) + tid;
    R.kmask = nk - 1; R.qb = nullptr;

    const uint32_t *src = pair_blob(blob, m);
    PairResult r;
    if (wide) {
        unpack_pair<true>(src, m.len2, R);
        r = extend_pair<FASTM, SYM, COUNT, true, true, BSW_WIN_GROUPS>(R, m.len2, m.len1, m.h0, P);
    } else {
        unpack_pair<false>(src, m.len2, R);
        r = extend_pair<FASTM, SYM, COUNT, false, true, BSW_WIN_GROUPS>(R, m.len2, m.len1, m.h0, P);
    }
    store_result(out, m.id, r);
}

// ---------------------------------------------------------------------------------------------
// Short pairs, two per thread (extend_duo2, bsw_duo.cuh; an evaluated alternative). Launch: block = kDuo2Threads, dynamic
// smem = 40 * nblk * kDuo2Threads (nblk = duo_blocks(longest query of the launch)). Thread t takes the launch's
// sorted pairs 2t and 2t + 1 (neighbours in (len2, len1, h0) order). A warp whose first thread still falls among
// the n_wide pairs holding an ambiguous base runs the LOP3-selector instantiation for all of its threads.
// ---------------------------------------------------------------------------------------------
#ifndef BSW_DUO2_NT
#define BSW_DUO2_NT 64
#endif
constexpr int kDuo2Threads = BSW_DUO2_NT;

template <bool FASTM, bool SYM, bool KEY>
__global__ void __launch_bounds__(kDuo2Threads)
bsw_duo2_kernel(const PairMeta *__restrict__ meta, const uint32_t *__restrict__ ord,
                const uint32_t *__restrict__ blob, PairOut *__restrict__ out, int n_wide, int n_narrow,
                KParams P, int nblk) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = kDuo2Threads;
    const int tid = threadIdx.x;
    const int t = blockIdx.x * NT + tid;
    const int n = n_wide + n_narrow;
    if (2 * t >= n) return;
    const bool twide = 2 * (t & ~31) < n_wide;       // warp-uniform
    const PairMeta mA = meta[ord[2 * t]];
    const bool hasB = 2 * t + 1 < n;
    PairMeta mB = mA;
    if (hasB) mB = meta[ord[2 * t + 1]];

    RowsD R;
    R.stride = NT;
    R.he4 = reinterpret_cast<uint4 *>(smem) + tid;
    R.qs = reinterpret_cast<uint2 *>(smem + (size_t)32 * nblk * NT) + tid;
    R.sbase = (uint32_t)__cvta_generic_to_shared(R.he4);

    DuoIn L[2];
    L[0].qlen = mA.len2; L[0].tlen = mA.len1; L[0].h0 = mA.h0; L[0].wide = mA.flags & 1;
    L[1].qlen = hasB ? mB.len2 : 0; L[1].tlen = hasB ? mB.len1 : 0; L[1].h0 = hasB ? mB.h0 : 0;
    L[1].wide = hasB && (mB.flags & 1);
    L[0].blob = pair_blob(blob, mA); L[1].blob = pair_blob(blob, mB);
    PairResult r[2];
    if (twide) extend_duo2<FASTM, SYM, true, KEY>(R, L, P, r);
    else extend_duo2<FASTM, SYM, false, KEY>(R, L, P, r);
    store_result(out, mA.id, r[0]);
    if (hasB) store_result(out, mB.id, r[1]);
}

// ---------------------------------------------------------------------------------------------
// Long pairs: ONE WARP per pair. The rows of the pair (same he4 / qs layout, stride 1) live in the
// warp's slice of shared memory; a row is swept left to right in tiles of 32 elements (128 columns),
// lane L owning element kb + L = columns 4(kb+L) .. +3 = two groups.
//
// The reference's decisions stay row by row (band clamp, row max / last argmax, m == 0, z-drop,
// leading / trailing trim): every lane holds the same beg / end / best / ... and the row is finished
// before the next one starts. Inside a row only F runs along the columns, and F is max-plus linear:
//     F(i, j + d) = max( F_local(j + d), F(i, j) - e_ins * d )
// so each lane first computes its four columns from F = 0 (pass 1: scores, M, T, E' and the local F
// chain), a 5-step warp scan (SHFL.UP + VIADDMNMX) then delivers every lane's true incoming F,
// and pass 2 corrects the lane's F with one VIADDMNMX per group before H = max(M, E, F), the shifted
// store of H (the diagonal of the next row; the left neighbour's last H arrives by SHFL.UP) and the row
// statistics. Row max and its LAST column, first / last non-zero entry are warp reductions
// (REDUX / ballots) instead of per-thread scans.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t warp_pair_bytes(int qlen) {
    return (uint32_t)(16 * row_elems(qlen) + 4 * sel_words(qlen) + 15) & ~15u;
}

template <bool FASTM, bool SYM, bool COUNT, bool WIDE>
__device__ inline PairResult warp_extend_pair(Rows &R, const uint32_t *__restrict__ blob, int qlen, int tlen,
                                              int h0, const KParams &P) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const uint32_t NEG_OE_DEL = pack2(-oe_del), NEG_OE_INS = pack2(-oe_ins);
    const uint32_t NEG_E_DEL = pack2(-P.e_del), NEG_E_INS = pack2(-P.e_ins);
    // pass-2 offsets of a lane's four columns from its first one: { 0, -e }, { -2e, -3e }
    const uint32_t OFF01 = ((uint32_t)(-P.e_ins) & 0xFFFFu) << 16;
    const uint32_t OFF23 = ((uint32_t)(-2 * P.e_ins) & 0xFFFFu) | (((uint32_t)(-3 * P.e_ins) & 0xFFFFu) << 16);
    const int dec = 4 * P.e_ins;    // F decay across one lane (4 columns)
    uint32_t LUT_LO, LUT_HI;
    score_lut<WIDE>(P, LUT_LO, LUT_HI);
    uint32_t K16 = P.k16, KM = P.km, K1 = P.k1;
#if !defined(BSW_HOST_EMUL) && BSW_PIN_CONSTS
    {   // as in extend_pair: keeps ptxas from re-loading them from the parameter bank inside the loops
        const uint32_t z = *reinterpret_cast<const volatile uint32_t *>(&g_zero);
        K16 ^= z; KM ^= z; K1 ^= z;
    }
#endif

    // ---- query selector seeds and row "-1" (bandedSWA.cpp:159-161), all lanes
    {
        const int nsel = sel_words(qlen);
        if (!WIDE) {
            for (int k = lane; k < nsel; k += 32) {
                uint32_t v = (blob[k >> 2] >> (8 * (k & 3))) & 0xFFu;   // 4 bases, 2 bits each
                v = (v | (v << 12)) & 0x000F000Fu;
                v = (v | (v << 6)) & 0x03030303u;
                R.QS2(k) = v * 0x11u;
            }
        } else {
            for (int k = lane; k < nsel; k += 32) {
                uint32_t v = (blob[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;  // 4 bases, 4 bits each
                v = (v | (v << 8)) & 0x00FF00FFu;
                v = (v | (v << 4)) & 0x0F0F0F0Fu;
                R.QS2(k) = v * 0x11u;
            }
        }
        R.tb = blob + (seq_bytes((uint32_t)qlen, WIDE) >> 2);
        const int nel = row_elems(qlen);
        for (int k = lane; k < nel; k += 32) {
            uint32_t hv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = 4 * k + u;
                int v = j == 0 ? h0 : max(h0 - oe_ins - (j - 1) * P.e_ins, 0);
                if (j > qlen) v = 0;
                hv[u] = (uint32_t)v;
            }
            uint4 w; w.x = hv[0] | (hv[1] << 16); w.y = 0u; w.z = hv[2] | (hv[3] << 16); w.w = 0u;
            R.HE4(k) = w;
        }
    }
    __syncwarp();

    const int band = pair_band(P, qlen);
    const int budget = min(qlen + band, tlen);
    int best = h0, best_i = -1, best_j = -1, g_i = -1, gsc = -1, off = 0;
    int beg = 0, end = qlen;
    uint32_t tword = 0, tnext = R.tb[0];
    int hcol = h0 - P.o_del;
    uint32_t cells = 0;

    for (int i = 0; i < budget; ++i) {
        if (beg < i - band) beg = i - band;
        if (end > i + band + 1) end = i + band + 1;
        if (beg >= end) break;
        if (COUNT) cells += (uint32_t)(end - beg);   // beg is the reference's exact beg in this kernel

        const uint32_t tsel = R.template row_seed<WIDE>(i, tword, tnext);
        hcol -= P.e_del;
        const int hleft = beg == 0 ? max(hcol, 0) : 0;

        // entries outside [beg, end) that share an element / a group with live columns must read as
        // zero (see extend_pair); everything left of beg is zero already
        if (lane == 0) {
            if (beg & 3) R.setHE16(beg - 1, 0u, 0u);
            if (end & 1) R.setHE16(end, 0u, 0u);
        }
        __syncwarp();

        const int k0 = beg >> 2, k1 = (end - 1) >> 2, g1 = (end - 1) >> 1;
        int Ftile = 0;                              // F(i, 4 * kb) entering the tile
        uint32_t hcarry = (uint32_t)hleft << 16;    // .hi = H(i, 4 * kb - 1)
        uint32_t rm = 0;                            // per lane: max per 16-bit half over its columns
        int klo = 0, khi = 0;                       // per lane: element where a half last reached rm
        uint32_t hw0 = 0, hw1 = 0;                  // this lane's H words of the last tile

        for (int kb = k0; kb <= k1; kb += 32) {
            const int k = kb + lane;
            const bool act = k <= k1;
            uint32_t M0 = 0, M1 = 0, T0 = 0, T1 = 0, E0 = 0, E1 = 0, Ev0 = 0, Ev1 = 0, B0 = 0, B1 = 0;
            int Fout = 0;
            if (act) {
                // ---- pass 1: everything that does not need the incoming F
                const uint4 a = R.HE4(k);
                const uint32_t q01 = R.QS2(k);
                uint32_t s0, s1;
                if (WIDE || BSW_SEL_LOP3) { s0 = sel_combine(q01, tsel, 0x44444444u); s1 = __umulhi(s0, K16); }
                else { s0 = q01 * K1 + tsel; s1 = __umulhi(s0, K16); }
                auto front = [&](const uint32_t Hd, const uint32_t Ev, const uint32_t sel, uint32_t &M, uint32_t &Tins,
                                 uint32_t &Enew) {
                    const uint32_t sc = prmt_sx(LUT_LO, LUT_HI, sel);
                    if (FASTM) {
                        M = __viaddmin_s16x2(Hd, sc, Hd * KM);
                    } else {
                        const uint32_t sm = __vmins2(sc, __vmins2(Hd, 0x00010001u) * (uint32_t)P.match);
                        M = __vadd2(Hd, sm);
                    }
                    const uint32_t Tdel = __viaddmax_s16x2_relu(M, NEG_OE_DEL, NEG_OE_DEL);
                    Tins = SYM ? Tdel : __viaddmax_s16x2_relu(M, NEG_OE_INS, NEG_OE_INS);
                    Enew = __viaddmax_s16x2(Ev, NEG_E_DEL, Tdel);
                };
                Ev0 = a.y; Ev1 = a.w;
                front(a.x, a.y, s0, M0, T0, E0);
                front(a.z, a.w, s1, M1, T1, E1);
                // local F chain from F = 0 at the lane's first column
                uint32_t A = 0;
                uint32_t W1 = __viaddmax_s16x2(A, NEG_E_INS, T0);
                B0 = W1 * K16 + A;
                uint32_t W2 = __viaddmax_s16x2(B0, NEG_E_INS, T0);
                A = __umulhi(W2, K16);
                W1 = __viaddmax_s16x2(A, NEG_E_INS, T1);
                B1 = W1 * K16 + A;
                W2 = __viaddmax_s16x2(B1, NEG_E_INS, T1);
                Fout = (int)__umulhi(W2, K16);      // F at the next lane's first column, given F = 0 here
            }
            // ---- max-plus scan of the lanes' outgoing F (inclusive), then this lane's incoming F
            int x = Fout;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(FULL, x, d);
                if (lane >= d) x = max(x, y - dec * d);
            }
            int Fin = __shfl_up_sync(FULL, x, 1);
            Fin = lane == 0 ? Ftile : max(Fin, Ftile - dec * lane);
            Ftile = max(__shfl_sync(FULL, x, 31), Ftile - dec * 32);
            // ---- pass 2: H, shifted store, statistics
            uint32_t h0w = 0, h1w = 0;
            if (act) {
                const uint32_t FinBB = (uint32_t)Fin * 0x00010001u;
                B0 = __viaddmax_s16x2(FinBB, OFF01, B0);
                B1 = __viaddmax_s16x2(FinBB, OFF23, B1);
                h0w = __vimax3_s16x2(M0, Ev0, B0);
                h1w = __vimax3_s16x2(M1, Ev1, B1);
            }
            uint32_t hp = __shfl_up_sync(FULL, h1w, 1);   // left neighbour's { H(4k-2), H(4k-1) }
            if (lane == 0) hp = hcarry;
            hcarry = __shfl_sync(FULL, h1w, 31);
            if (act) {
                const bool g1ok = 2 * k + 1 <= g1;          // the lane's second group holds a live column
                uint4 o;
                o.x = __umulhi(hp, K16) + h0w * K16;        // { H(4k-1), H(4k) }
                o.y = E0;
                o.z = __umulhi(h0w, K16) + h1w * K16;       // { H(4k+1), H(4k+2) }
                o.w = E1;
                if (g1ok) R.HE4(k) = o;
                else R.HE(2 * k) = make_uint2(o.x, o.y);    // entries right of `end` keep their stale values
                if (!g1ok) h1w = 0;                          // not part of the row
                bool phi, plo;
                rm = __vibmax_s16x2(__vmaxs2(h0w, h1w), rm, &phi, &plo);
                if (plo) klo = k;
                if (phi) khi = k;
            }
            hw0 = h0w; hw1 = h1w;
        }
        __syncwarp();

        // ---- row end, all lanes with the same values
        // H of the last column, and the reference's eh[end] = { h1, 0 }
        int hlast;
        {
            const int src = (k1 - k0) & 31;                   // lane that owns element k1 in the last tile
            const uint32_t w = __shfl_sync(FULL, (g1 & 1) ? hw1 : hw0, src);
            hlast = (end & 1) ? (int)(w & 0xFFFFu) : (int)(w >> 16);
            if (!(end & 1) && lane == 0) R.setHE16(end, (uint32_t)hlast, 0u);
        }
        if (end == qlen) {                                     // bandedSWA.cpp:218-221
            if (!(gsc > hlast)) g_i = i;
            gsc = max(gsc, hlast);
        }
        const int mlo = (int)(short)(rm & 0xFFFFu), mhi = (int)(short)(rm >> 16);
        const int m = __reduce_max_sync(FULL, max(mlo, mhi));
        if (m == 0) break;
        // LAST column reaching m: H(i, j) sits in Hs[j + 1]; a lane whose half reached m looks its
        // element up (columns at or right of `end` are never candidates: their H is below m)
        __syncwarp();
        int cand = -1;
        if (mlo == m) {                                        // even columns 4k, 4k+2
            const int c2 = 4 * klo + 2;
            cand = (c2 < end && R.getH16(c2 + 1) == (uint32_t)m) ? c2 : 4 * klo;
        }
        if (mhi == m) {                                        // odd columns 4k+1, 4k+3
            const int c3 = 4 * khi + 3;
            cand = max(cand, (c3 < end && R.getH16(c3 + 1) == (uint32_t)m) ? c3 : 4 * khi + 1);
        }
        const int mj = __reduce_max_sync(FULL, cand);
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            off = max(off, abs(mj - i));
        } else {
            const int di = i - best_i, dj = mj - best_j;
            if (best - m - abs(di - dj) > P.zdrop) break;
        }

        // leading trim, exact (bandedSWA.cpp:234): first j in [beg, end) with Hs[j] | E[j] != 0
        {
            int nb = end;
            for (int kb = k0; kb <= k1; kb += 32) {
                const int k = kb + lane;
                uint4 z = make_uint4(0u, 0u, 0u, 0u);
                if (k <= k1) z = R.HE4(k);
                const uint32_t w0 = z.x | z.y, w1 = z.z | z.w;
                int first = 0x7FFFFFFF;
                if (w0 | w1) first = 4 * k + ((w0 & 0xFFFFu) ? 0 : (w0 ? 1 : ((w1 & 0xFFFFu) ? 2 : 3)));
                const int f = __reduce_min_sync(FULL, first);
                if (f != 0x7FFFFFFF) { nb = min(f, end); break; }
            }
            beg = max(beg, nb);
        }
        // trailing trim (bandedSWA.cpp:236-237): j* = last j in [beg, end] with Hs[j] | E[j] != 0
        if (hlast) {
            end = min(end + 2, qlen);
        } else {
            int jstar = -1;
            for (int kb = end >> 2; kb >= 0 && jstar < 0; kb -= 32) {
                const int k = kb - lane;
                int last = -1;
                if (k >= 0) {
                    const uint4 z = R.HE4(k);
                    uint32_t w0 = z.x | z.y, w1 = z.z | z.w;
                    // entries right of `end` are stale
                    if (4 * k + 3 > end) w1 &= 0xFFFFu;
                    if (4 * k + 2 > end) w1 = 0u;
                    if (4 * k + 1 > end) w0 &= 0xFFFFu;
                    if (w1 >> 16) last = 4 * k + 3;
                    else if (w1) last = 4 * k + 2;
                    else if (w0 >> 16) last = 4 * k + 1;
                    else if (w0) last = 4 * k;
                }
                jstar = __reduce_max_sync(FULL, last);
            }
            end = min(jstar + 2, qlen);
        }
        __syncwarp();
    }

    PairResult r;
    r.score = best; r.qle = best_j + 1; r.tle = best_i + 1;
    r.gtle = g_i + 1; r.gscore = gsc; r.max_off = off;
    r.cells = cells;
    return r;
}

// Launch: block = 32 * warps_per_block threads, dynamic smem = warps_per_block * pair_bytes; pair p =
// blockIdx.x * warps_per_block + warp of the launch's n_wide + n_narrow pairs (wide ones first).
template <bool FASTM, bool SYM, bool COUNT>
__global__ void __launch_bounds__(256)
bsw_long_kernel(const PairMeta *__restrict__ meta, const uint32_t *__restrict__ ord,
                const uint32_t *__restrict__ blob, PairOut *__restrict__ out, int n_wide, int n_narrow,
                KParams P, int row_el, int pair_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    const int p = blockIdx.x * (blockDim.x >> 5) + warp;
    if (p >= n_wide + n_narrow) return;
    const PairMeta m = meta[ord[p]];
    Rows R;
    R.stride = 1;
    R.he4 = reinterpret_cast<uint4 *>(smem + (size_t)warp * pair_bytes);
    R.qs = reinterpret_cast<uint32_t *>(smem + (size_t)warp * pair_bytes + (size_t)16 * row_el);
    R.tb = nullptr; R.kmask = -1; R.qb = nullptr;
    const uint32_t *src = pair_blob(blob, m);
    PairResult r;
    // one pair per warp: the pair's own flag picks the instantiation (launches of this kernel merge several
    // length bins, so the pairs holding an ambiguous base are not all in front)
    if (m.flags & 1) r = warp_extend_pair<FASTM, SYM, COUNT, true>(R, src, m.len2, m.len1, m.h0, P);
    else r = warp_extend_pair<FASTM, SYM, COUNT, false>(R, src, m.len2, m.len1, m.h0, P);
    if ((threadIdx.x & 31) == 0) store_result(out, m.id, r);
}

// ---------------------------------------------------------------------------------------------
// Length binning on the device: sort key of every pair of a slab (descending order = launch order) and
// the identity permutation; cub's radix sort then yields ord[]. Pairs with an empty sequence (answered
// on the host) get key 0 and sort behind everything that is launched.
//   key = (bin << 5 | wide << 4 | (len2 - 1) % 16) << (b1 + b0) | len1 << b0 | h0,  bin = (len2 - 1) / 16
// len1 orders the pairs of a warp by their number of rows, h0 by the width of their first rows (the
// zero frontier of row i sits about h0 + i columns right of the diagonal): threads of a warp then run
// the same number of inner-loop trips (measured on config 3: 86 % of the lane slots busy, against 81 %
// without h0).
// ---------------------------------------------------------------------------------------------
// The fields are packed as tightly as the slab's largest len1 / h0 allow (b1, b0 bits, chosen by the
// host from its pass), so the radix sort runs over as few 8-bit digits as possible.
// Launch bins >= long_bin0 run as ONE windowed-rows launch: their top field is 1 | wide | len2 - 1 (17
// bits), which sorts them in front of everything else, their wide pairs first.
__host__ __device__ inline uint64_t sort_key(uint32_t len2, uint32_t len1, uint32_t h0, uint32_t wide, int b1, int b0,
                                             int long_bin0) {
    if (len2 == 0 || len1 == 0) return 0ull;
    const uint32_t v = len2 - 1;
    const uint32_t top = (int)(v >> 4) >= long_bin0 ? (0x10000u | (wide << 15) | v) : (((v >> 4) << 5) | (wide << 4) | (v & 15u));
    return ((uint64_t)top << (b1 + b0)) | ((uint64_t)len1 << b0) | h0;
}
#ifndef BSW_HOST_EMUL
// chunk_base (nullable; packed input): meta[k].off is relative to its chunk of 4096 pairs, the chunk's first word is
// chunk_base[k >> 12] -- added here, so that the host writes the records in one pass without a prefix sum in between
__global__ void bsw_key_kernel(PairMeta *__restrict__ meta, int n, uint64_t *__restrict__ keys,
                               uint32_t *__restrict__ idx, int b1, int b0, int long_bin0,
                               PairOut *__restrict__ out, const uint32_t *__restrict__ chunk_base) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    PairMeta m = meta[k];
    if (chunk_base) {
        m.off += chunk_base[k >> 12];
        meta[k].off = m.off;
    }
    keys[k] = sort_key(m.len2, m.len1, (uint32_t)m.h0, m.flags & 1u, b1, b0, long_bin0);
    idx[k] = (uint32_t)k;
    // empty target or query: no DP kernel is launched for the pair (key 0 sorts behind every launch), the
    // DP loop of the reference never runs (bandedSWA.cpp:181 with tlen == 0 / end == 0). Its answer is
    // written here so that the slab's result records are complete on the device.
    if (m.len2 == 0 || m.len1 == 0) {
        PairResult r;
        r.score = m.h0; r.qle = 0; r.tle = 0; r.gtle = 0; r.gscore = -1; r.max_off = 0; r.cells = 0;
        store_result(out, m.id, r);
    }
}
#endif

// ---------------------------------------------------------------------------------------------
// Packed input (bsw_gpu_batch_packed): the slab's 12-byte records -> PairMeta, ON THE DEVICE, so that the host's work
// per slab does not grow with the pairs (it only waits for the small statistics block below and plans the launches).
//   bsw_rec_meta_kernel : one block per chunk of 4096 pairs. Thread t owns 16 consecutive records: sizes them,
//                         a block-wide exclusive sum gives its first word inside the chunk, it writes the 16
//                         PairMeta (offsets chunk-relative: bsw_key_kernel adds the chunk's base), validates, and
//                         feeds the (wide, len2 bin) histogram (privatised in shared memory) and the slab maxima.
//   bsw_chunk_scan_kernel: exclusive sum of the chunk totals -> chunk bases, total words.
// ---------------------------------------------------------------------------------------------
struct PackedRecDev { uint16_t len1, len2; int32_t h0; uint32_t flags; };   // == bsw_packed_rec
constexpr int kRecChunk = 4096;          // pairs per chunk (the host's cut granularity as well)
constexpr int kRecBins = 2049;           // == BSW_MAX_SEQ_LEN / 16 + 2 (launch bins of 16 query lengths)
struct SlabStatsDev {
    uint32_t hist[2][kRecBins];          // [wide][(len2 - 1) / 16]
    int32_t maxq, maxsc, maxt, maxh;     // longest query, largest h0 + min(len1, len2) * match, longest target, largest h0
    uint32_t ntriv, bad;                 // pairs with an empty sequence; records outside the domain
    uint64_t total_words;
};

#ifndef BSW_HOST_EMUL
__global__ void __launch_bounds__(256)
bsw_rec_meta_kernel(const PackedRecDev *__restrict__ rec, int n, int match, PairMeta *__restrict__ meta,
                    uint32_t *__restrict__ chunk_total, SlabStatsDev *__restrict__ stats) {
    __shared__ uint32_t s_hist[2][kRecBins];
    __shared__ uint32_t s_warp[8];
    __shared__ int s_max[4];
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * kRecBins; i += 256) (&s_hist[0][0])[i] = 0u;
    if (tid < 4) s_max[tid] = 0;
    __syncthreads();
    const int base = blockIdx.x * kRecChunk + tid * 16;
    uint32_t words[16];
    uint32_t sum = 0, ntriv = 0, bad = 0;
    int maxq = 0, maxsc = 0, maxt = 0, maxh = 0;
    PackedRecDev r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = base + i;
        words[i] = 0;
        if (k < n) {
            r[i] = rec[k];
            const bool wide = r[i].flags & 1u;
            words[i] = (seq_bytes(r[i].len2, wide) + seq_bytes(r[i].len1, wide)) >> 2;
            const int sc = r[i].h0 + (int)min(r[i].len1, r[i].len2) * match;
            if (r[i].len1 > 32767 || r[i].len2 > 32767 || r[i].h0 < 0 || sc > 32767) bad = 1;
            else if (r[i].len1 == 0 || r[i].len2 == 0) ++ntriv;
            else {
                atomicAdd(&s_hist[wide ? 1 : 0][(r[i].len2 - 1) >> 4], 1u);
                maxq = max(maxq, (int)r[i].len2); maxsc = max(maxsc, sc);
                maxt = max(maxt, (int)r[i].len1); maxh = max(maxh, r[i].h0);
            }
        }
        sum += words[i];
    }
    // block-wide exclusive sum of the threads' totals (warp scan + scan of the 8 warp totals)
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((tid & 31) >= d) incl += y;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < (tid >> 5)) woff += s_warp[w];
        total += s_warp[w];
    }
    uint32_t off = woff + incl - sum;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = base + i;
        if (k < n) {
            PairMeta m;
            m.off = off; m.id = (uint32_t)k; m.len2 = r[i].len2; m.len1 = r[i].len1;
            m.h0 = (int16_t)r[i].h0; m.flags = (uint16_t)((r[i].flags & 1u) ? 3u : 0u);
            meta[k] = m;
            off += words[i];
        }
    }
    // statistics: warp-reduced, then one atomic per warp / block
    maxq = __reduce_max_sync(0xFFFFFFFFu, maxq); maxsc = __reduce_max_sync(0xFFFFFFFFu, maxsc);
    maxt = __reduce_max_sync(0xFFFFFFFFu, maxt); maxh = __reduce_max_sync(0xFFFFFFFFu, maxh);
    ntriv = __reduce_add_sync(0xFFFFFFFFu, ntriv); bad = __reduce_or_sync(0xFFFFFFFFu, bad);
    if ((tid & 31) == 0) {
        atomicMax(&s_max[0], maxq); atomicMax(&s_max[1], maxsc); atomicMax(&s_max[2], maxt); atomicMax(&s_max[3], maxh);
        if (ntriv) atomicAdd(&stats->ntriv, ntriv);
        if (bad) atomicOr(&stats->bad, 1u);
    }
    __syncthreads();
    if (tid == 0) {
        chunk_total[blockIdx.x] = total;
        atomicMax(&stats->maxq, s_max[0]); atomicMax(&stats->maxsc, s_max[1]);
        atomicMax(&stats->maxt, s_max[2]); atomicMax(&stats->maxh, s_max[3]);
    }
    for (int i = tid; i < 2 * kRecBins; i += 256) {
        const uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&stats->hist[0][0] + i, c);
    }
}

__global__ void __launch_bounds__(256)
bsw_chunk_scan_kernel(const uint32_t *__restrict__ chunk_total, int nch, uint32_t *__restrict__ chunk_base,
                      SlabStatsDev *__restrict__ stats) {
    __shared__ uint64_t s_part[256];
    const int tid = threadIdx.x;
    const int per = (nch + 255) / 256;
    uint64_t sum = 0;
    for (int i = 0; i < per; ++i) { const int c = tid * per + i; if (c < nch) sum += chunk_total[c]; }
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        uint64_t run = 0;
        for (int t = 0; t < 256; ++t) { const uint64_t x = s_part[t]; s_part[t] = run; run += x; }
        stats->total_words = run;
    }
    __syncthreads();
    uint64_t run = s_part[tid];
    for (int i = 0; i < per; ++i) {
        const int c = tid * per + i;
        if (c < nch) { chunk_base[c] = (uint32_t)run; run += chunk_total[c]; }
    }
}
#endif

// ---------------------------------------------------------------------------------------------
// The scalar class: pairs whose score bound h0 + len2 * match leaves int16. bwa-mem2 sorts them out a priori
// (bwamem.cpp:2218-2228, the third class of sortPairsLenExt :1846-1925) and runs them through the SCALAR kernel
// (scalarBandedSWAWrapper, bwamem.cpp:2384-2390 -> bandedSWA.cpp:132-276), so its rules apply, not the vector
// path's: int32 arithmetic, the band of :164-172, the z-drop test WITH the gap-extend factor and the zdrop > 0 guard
// (:226-231), ambiguous bases scored from the matrix (= P.ambig), no row budget. One thread per pair over
// byte-per-base sequences and int32 rows in global memory: the class is rare by construction (a seed score above
// ~32 000), so this path is about not failing the batch, not about speed.
// ---------------------------------------------------------------------------------------------
struct BigMeta {
    uint64_t toff, qoff;   // byte offsets of target / query in `seq`
    uint64_t soff;         // int32 offset of this pair's rows in `scratch` (2 * (len2 + 2) entries)
    int32_t len1, len2, h0;
    uint32_t pad;
};
struct BigOut { int32_t score, qle, tle, gtle, gscore, max_off; };

#ifndef BSW_HOST_EMUL
__global__ void bsw_big_kernel(const BigMeta *__restrict__ meta, int n, const uint8_t *__restrict__ seq,
                               int32_t *__restrict__ scratch, BigOut *__restrict__ out, KParams P) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const BigMeta m = meta[k];
    const uint8_t *tgt = seq + m.toff, *qry = seq + m.qoff;
    const int qlen = m.len2, tlen = m.len1, h0 = m.h0;
    int32_t *Hd = scratch + m.soff, *Ev = Hd + (qlen + 2);
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    for (int j = 0; j < qlen + 2; ++j) { Hd[j] = 0; Ev[j] = 0; }
    // row "-1" (bandedSWA.cpp:159-161)
    Hd[0] = h0;
    if (qlen >= 1) Hd[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (int j = 2; j <= qlen && Hd[j - 1] > P.e_ins; ++j) Hd[j] = Hd[j - 1] - P.e_ins;
    // band (:164-172)
    int band = P.w;
    {
        const int mx = P.max_score;
        int max_ins = (int)((double)(qlen * mx + P.end_bonus - P.o_ins) / P.e_ins + 1.);
        int max_del = (int)((double)(qlen * mx + P.end_bonus - P.o_del) / P.e_del + 1.);
        band = min(min(band, max(max_ins, 1)), max(max_del, 1));
    }
    int best = h0, best_i = -1, best_j = -1, g_i = -1, g = -1, off = 0;
    int beg = 0, end = qlen;
    for (int i = 0; i < tlen; ++i) {
        if (beg < i - band) beg = i - band;
        if (end > i + band + 1) end = i + band + 1;
        if (end > qlen) end = qlen;
        int hleft = 0;                                  // H(i, beg-1)
        if (beg == 0) hleft = max(h0 - (P.o_del + P.e_del * (i + 1)), 0);
        const int t = tgt[i];
        int f = 0, rowmax = 0, rowarg = -1, j;
        for (j = beg; j < end; ++j) {
            const int d = Hd[j], e = Ev[j];
            Hd[j] = hleft;
            const int q = qry[j];
            const int sc = (t >= 4 || q >= 4) ? P.ambig : (t == q ? P.match : -P.mismatch);
            const int M = d ? d + sc : 0;
            const int h = max(max(M, e), f);
            hleft = h;
            if (h >= rowmax) { rowmax = h; rowarg = j; }   // LAST column reaching the row max (:204-205)
            int tt = max(M - oe_del, 0);
            Ev[j] = max(e - P.e_del, tt);
            tt = max(M - oe_ins, 0);
            f = max(f - P.e_ins, tt);
        }
        Hd[end] = hleft; Ev[end] = 0;
        if (j == qlen) {                                 // :218-221
            if (!(g > hleft)) g_i = i;
            g = max(g, hleft);
        }
        if (rowmax == 0) break;
        if (rowmax > best) {
            best = rowmax; best_i = i; best_j = rowarg;
            off = max(off, abs(rowarg - i));
        } else if (P.zdrop > 0) {                        // :226-231
            const int di = i - best_i, dj = rowarg - best_j;
            const int pen = di > dj ? (di - dj) * P.e_del : (dj - di) * P.e_ins;
            if (best - rowmax - pen > P.zdrop) break;
        }
        for (j = beg; j < end && Hd[j] == 0 && Ev[j] == 0; ++j) {}
        beg = j;
        for (j = end; j >= beg && Hd[j] == 0 && Ev[j] == 0; --j) {}
        end = min(j + 2, qlen);
    }
    BigOut r;
    r.score = best; r.qle = best_j + 1; r.tle = best_i + 1; r.gtle = g_i + 1; r.gscore = g; r.max_off = off;
    out[k] = r;
}
#endif

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

// WHICH == 9 of bsw_gpu_dpx_peak: the arithmetic of one inner-loop trip of extend_pair<.., KEY> (four
// groups = eight cells: selector, PRMT, M, T, E', the F scan, H, shifted store word, keyed row max) on
// registers only -- no shared memory, no row bookkeeping, no divergence. Its rate is the ceiling the
// thread-per-pair kernel could reach if everything but the recurrence were free.
__global__ void bsw_trip_peak_kernel(uint32_t *sink, int iters, uint32_t seed, uint32_t k16, uint32_t km, uint32_t k1,
                                     uint32_t kk) {
    uint32_t hd[4], ev[4], q[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) { hd[k] = (seed * (k + 3) + threadIdx.x) & 0x00FF00FFu; ev[k] = (seed * (k + 7)) & 0x003F003Fu; }
    q[0] = 0x11002233u & (seed | 0x33333333u); q[1] = 0x22113300u;
    const uint32_t LUT_LO = 0xFCFCFCFCu, LUT_HI = 0xFCFCFC01u, NEG_OE = pack2(-7), NEG_E = pack2(-1);
    uint32_t rm = 0, A = 0, hprev = 0, tsel = 0x94949494u;
    {   // as in extend_pair (BSW_PIN_CONSTS)
        const uint32_t z = *reinterpret_cast<const volatile uint32_t *>(&g_zero);
        k16 ^= z; km ^= z; k1 ^= z; kk ^= z;
    }
    for (int it = 0; it < iters; ++it) {
        uint32_t M[4], T[4], E[4], hv[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const uint32_t s0 = q[e] * k1 + tsel, s1 = __umulhi(s0, k16);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int g = 2 * e + u;
                const uint32_t sc = prmt_sx(LUT_LO, LUT_HI, u ? s1 : s0);
                M[g] = __viaddmin_s16x2(hd[g], sc, hd[g] * km);
                T[g] = __viaddmax_s16x2_relu(M[g], NEG_OE, NEG_OE);
                E[g] = __viaddmax_s16x2(ev[g], NEG_E, T[g]);
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint32_t W1 = __viaddmax_s16x2(A, NEG_E, T[g]);
            const uint32_t B = W1 * k16 + A;
            const uint32_t h = __vimax3_s16x2(M[g], ev[g], B);
            const uint32_t W2 = __viaddmax_s16x2(B, NEG_E, T[g]);
            A = __umulhi(W2, k16);
            hd[g] = __byte_perm(hprev, h, 0x5432);   // next "row" reads what this one stored
            hprev = h;
            hv[g] = h;
            ev[g] = E[g];
        }
        const uint32_t t3 = __vimax3_u16x2(hv[0] * kk, hv[1] * kk + 0x00010001u, hv[2] * kk + 0x00020002u);
        const uint32_t t4 = __vmaxu2(t3, hv[3] * kk + 0x00030003u);
        rm = __viaddmax_u16x2(t4, (uint32_t)it * 0x00010001u, rm);
        tsel += 0x01010101u & (uint32_t)it;
    }
    if ((rm ^ A) == 0x12345678u) sink[threadIdx.x] = rm;
}

// WHICH == 10 / 11 of bsw_gpu_dpx_peak: the arithmetic of one inner-loop trip of extend_duo2<.., KEY> (four
// columns of two pairs = eight cells: selector, PRMT, M, T, E', H, F', keyed row max) on registers only; 10 at
// full occupancy, 11 at five one-warp blocks per SM (what the shared memory of the longest config-3 bins allows).
__global__ void bsw_duo_trip_peak_kernel(uint32_t *sink, int iters, uint32_t seed, uint32_t k16, uint32_t km, uint32_t k1,
                                         uint32_t kk) {
    uint32_t hd[4], ev[4], q[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) { hd[k] = (seed * (k + 3) + threadIdx.x) & 0x00FF00FFu; ev[k] = (seed * (k + 7)) & 0x003F003Fu; }
    q[0] = 0x11002233u & (seed | 0x33333333u); q[1] = 0x22113300u;
    const uint32_t LUT_LO = 0xFCFCFCFCu, LUT_HI = 0xFCFCFC01u, NEG_OE = pack2(-7), NEG_E = pack2(-1);
    uint32_t rm = 0, F = 0, hprev = 0, tsel = 0x94949494u, J2 = 0;
    {
        const uint32_t z = *reinterpret_cast<const volatile uint32_t *>(&g_zero);
        k16 ^= z; km ^= z; k1 ^= z; kk ^= z;
    }
    for (int it = 0; it < iters; ++it) {
        uint32_t s[4], hv[4];
        s[0] = q[0] * k1 + tsel; s[1] = __umulhi(s[0], k16);
        s[2] = q[1] * k1 + tsel; s[3] = __umulhi(s[2], k16);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t sc = prmt_sx(LUT_LO, LUT_HI, s[u]);
            const uint32_t M = __viaddmin_s16x2(hd[u], sc, hd[u] * km);
            const uint32_t T = __viaddmax_s16x2_relu(M, NEG_OE, NEG_OE);
            const uint32_t En = __viaddmax_s16x2(ev[u], NEG_E, T);
            hv[u] = __vimax3_s16x2(M, ev[u], F);
            F = __viaddmax_s16x2(F, NEG_E, T);
            hd[u] = hprev;          // next "row" reads what this one stored
            hprev = hv[u];
            ev[u] = En;
        }
        const uint32_t t3 = __vimax3_u16x2(hv[0] * kk, hv[1] * kk + 0x00010001u, hv[2] * kk + 0x00020002u);
        const uint32_t t4 = __vmaxu2(t3, hv[3] * kk + 0x00030003u);
        rm = __viaddmax_u16x2(t4, J2, rm);
        J2 += 0x00040004u;
        tsel += 0x01010101u & (uint32_t)it;
    }
    if ((rm ^ F) == 0x12345678u) sink[threadIdx.x] = rm;
}

#endif  // !BSW_HOST_EMUL

}  // namespace bswk
