// kswv on the GPU: bwa-mem2's batched mate-rescue Smith-Waterman (class kswv,
// /root/reference/benchmarks/fmi/bwa-mem2/x86_64/src/kswv.cpp), one WARP per pair.
//
// What is computed is what one lane of the reference's 64/32-lane vector kernels computes (kswv512_u8
// kswv.cpp:371-716, kswv512_16 :933-1215; see oracle/kswv_oracle.c for the statement-by-statement
// restatement): a full (unbanded) local-alignment DP over len1 rows x the query padded to a multiple of 16
// (8-bit class) or 8 (16-bit class) zero-score columns, the row maxima with their first column, gmax/te/qe,
// the early stop on KSW_XSTOP / 8-bit saturation, the second-best row maximum from the reference's
// rising-row filter, and the reverse pass for tb/qb (phase 1 of mem_sam_pe_batch, bwamem_pair.cpp:660-699).
//
// Mapping. The query's padded columns are split into 32 strips of C consecutive columns, one per lane
// (C = ceil(columns / 32), 1..8; queries above 256 columns take several 256-column passes with the boundary
// column kept in global memory). Lane k works on row s - k at step s: H and F of its strip live in
// registers, and what a row needs from the left -- H and E of the strip's last column and the running row
// maximum -- arrives by one shuffle pair per step. A cell is 10 integer instructions: one PRMT looks up the
// signed score, VIADDMNMX adds it to the diagonal and clamps (the 8-bit class's saturation is that clamp, at
// 255 - shift), VIMNMX3 with relu forms H, two IADD + two VIADDMNMX update E and F, and LEA + VIADDMNMX keep
// the row maximum as a key (H << 16 | 0xFFFF - column) whose maximum is the FIRST column of the largest H.
// Columns added on the left to fill the strips score negative against everything and therefore stay at zero.
// The last active lane sees the finished row: it updates gmax/te/qe, stores the row maximum (2 bytes) for the
// second-best pass and raises the stop flag, which ends the warp's loop at the next step.
#pragma once
#include <stdint.h>

#ifdef BSW_HOST_EMUL
#include "dpx_host_emul.h"
#include "warp_fibers.h"
static inline int __viaddmin_s32(int a, int b, int c) { return std::min((int)((uint32_t)a + (uint32_t)b), c); }
static inline int __viaddmax_s32_relu(int a, int b, int c) { return std::max(std::max((int)((uint32_t)a + (uint32_t)b), c), 0); }
static inline int __vimax3_s32_relu(int a, int b, int c) { return std::max(std::max(std::max(a, b), c), 0); }
static inline int __vimax3_s32(int a, int b, int c) { return std::max(std::max(a, b), c); }
#else
#include <cuda_runtime.h>
#endif

namespace kswvk {

constexpr int kXByte = 0x10000, kXStop = 0x20000, kXSubo = 0x40000, kXStart = 0x80000;   // ksw.h:31-34
constexpr int kPassCols = 256;      // columns one pass of a warp covers (32 lanes x 8)
constexpr int kNoStop = 0x7FFFFFFF;
constexpr int kScratchSlack = 48;   // words of per-group scratch past the longest reference (LUT words are read ahead)

struct KParams {
    int32_t a, b, amb;                        // match, mismatch (negative), ambiguous (-1, kswv.cpp:131)
    int32_t oe_del, e_del, oe_ins, e_ins;
    int32_t qmax, shift;                      // g_qmax (kswv.cpp:132-133), the 8-bit bias (kswv.cpp:396-404)
    int32_t one;                              // 1, opaque to the compiler (see kswv_dp)
    uint32_t lut_mis, lut_ab, lut_amb, lut_hi;
};

struct Task { uint32_t roff, qoff; int32_t tlen, qlen, xtra, out; };
struct Result { int32_t score, te, qe, score2, te2, tb, qb; };   // == kswr_t, ksw.h:45-50

// ------------------------------------------------------------------ lane-group primitives
// A pair is worked on by a group of W = 8, 16 or 32 adjacent lanes (32 / W pairs per warp). Every collective of
// the per-pair code names its group's mask, so groups of one warp may run different trip counts (the task order keeps
// them close); they meet again where the warp fetches its next tasks.
#ifdef BSW_HOST_EMUL
inline int hw_lane() { return wf::lane(); }
inline uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t sel) { return emul::prmt(a, b, sel); }
inline void sync_all() { wf::syncwarp(0xFFFFFFFFu); }
#else
__device__ __forceinline__ int hw_lane() { return (int)(threadIdx.x & 31u); }
// PTX prmt.b32, default mode: a selector nibble with bit 3 set replicates the selected byte's sign bit
__device__ __forceinline__ uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ void sync_all() { __syncwarp(0xFFFFFFFFu); }
#endif

template <int W>
struct Grp {
    static __device__ __forceinline__ int lane() { return hw_lane() & (W - 1); }
    static __device__ __forceinline__ int base() { return hw_lane() & ~(W - 1); }
    static __device__ __forceinline__ uint32_t mask() {
        return W == 32 ? 0xFFFFFFFFu : (((1u << (W & 31)) - 1u) << base());
    }
#ifdef BSW_HOST_EMUL
    static uint32_t up1(uint32_t v) { return wf::shfl_up1(v, mask(), W); }
    static uint32_t from(uint32_t v, int src) { return wf::shfl(v, src, mask(), W); }
    static bool any(bool p) { return wf::any(p, mask()); }
    static uint32_t ballot(bool p) { return W == 32 ? wf::ballot(p, mask()) : (wf::ballot(p, mask()) >> base()); }
    static uint32_t maxu(uint32_t v) { return wf::reduce_max(v, mask()); }
    static void sync() { wf::syncwarp(mask()); }
#else
    static __device__ __forceinline__ uint32_t up1(uint32_t v) { return __shfl_up_sync(mask(), v, 1, W); }
    static __device__ __forceinline__ uint32_t from(uint32_t v, int src) { return __shfl_sync(mask(), v, src, W); }
    static __device__ __forceinline__ bool any(bool p) { return __any_sync(mask(), p) != 0; }
    static __device__ __forceinline__ uint32_t ballot(bool p) {
        return W == 32 ? __ballot_sync(mask(), p) : (__ballot_sync(mask(), p) >> base());
    }
    static __device__ __forceinline__ uint32_t maxu(uint32_t v) { return __reduce_max_sync(mask(), v); }
    static __device__ __forceinline__ void sync() { __syncwarp(mask()); }
#endif
};

// host side of KParams: the two LUT words. Byte q of lut_lo(r) is the score of reference base r against query
// code q = 0..3; lut_hi holds query codes 4 (ambiguous), 5 (zero-score padding column), 6 (left filler: negative).
inline KParams make_kparams(int o_del, int e_del, int o_ins, int e_ins, int match, int mismatch) {
    KParams K;
    K.a = match; K.b = -mismatch; K.amb = -1;
    K.oe_del = o_del + e_del; K.e_del = e_del; K.oe_ins = o_ins + e_ins; K.e_ins = e_ins;
    K.qmax = match > K.amb ? match : K.amb;
    if (K.b > K.qmax) K.qmax = K.b;
    int mn = K.a < K.b ? K.a : K.b;
    if (K.amb < mn) mn = K.amb;
    K.shift = (uint8_t)(256 - (uint8_t)mn);
    const uint32_t A = (uint8_t)(int8_t)K.a, B = (uint8_t)(int8_t)K.b, N = (uint8_t)(int8_t)K.amb;
    K.lut_mis = B * 0x01010101u;
    K.lut_ab = A ^ B;
    K.lut_amb = N * 0x01010101u;
    K.lut_hi = N | (0u << 8) | (N << 16) | (N << 24);
    K.one = 1;
    return K;
}

struct Best { int32_t gmax, te, qe, rows; bool dead; };

constexpr int kMaxC32 = 8, kMaxC16 = 16, kMaxC8 = 20;     // columns per lane, by group width
// widest padded query a group of W lanes takes in one pass
__host__ __device__ constexpr int group_cols(int W) { return W == 32 ? 32 * kMaxC32 : (W == 16 ? 16 * kMaxC16 : 8 * kMaxC8); }

// The DP of tlen reference rows x q[0..qlen) (padded) on one group of W lanes.
//  lutw  : per reference row, the four scores of its base against query codes 0..3 (one byte each)
//  thr   : gmax >= thr ends the pair (kNoStop: never)
//  rowkey: row i's key (row maximum << 16 | 0xFFFF - its first column, fillers included) for rows 0 .. rows-1
//  bnd   : boundary column between passes (MP: queries above 256 columns, W = 32 only), tlen entries
// The padded columns are cut into strips of C consecutive columns, one per lane; lane k works on row s - k at step
// s. gmax / te / qe are taken from the stored keys after the loop: the lane that sees a finished row only stores its
// key and tests the stop threshold, because whatever a single lane does costs the whole warp an issue slot.
// SAT: the diagonal term can reach the 8-bit class's ceiling (255 - shift) and is clamped there by a VIADDMNMX.
// When min(tlen, qlen) * match stays below the ceiling (every pair bwa-mem2 puts in the 8-bit class: l_ms * a < 250)
// and in the 16-bit class, the sum is an IMAD by a run-time 1 instead -- the FMA pipe has room, the ALU pipe binds.
template <int W, int C, bool SAT, bool MP>
__device__ __noinline__ Best kswv_dp(const KParams &K, const uint32_t *__restrict__ lutw, int tlen,
                                     const uint8_t *__restrict__ q, int qlen, bool byte, int thr,
                                     uint32_t *rowkey, uint2 *bnd) {
    typedef Grp<W> G;
    const int k = G::lane();
    const int quantum = byte ? 16 : 8;
    int ncol = (qlen + quantum - 1) / quantum * quantum;
    if (ncol == 0) ncol = quantum;          // an empty query is all padding: every H stays 0
    const int npass = MP ? (ncol + kPassCols - 1) / kPassCols : 1;
    // scalars by value: K sits in the caller's local memory and would be re-read every step
    const int clamp = byte ? 255 - K.shift : 32767;
    const int noe_ins = -K.oe_ins, noe_del = -K.oe_del, e_ins = K.e_ins, e_del = K.e_del, one = K.one;
    const int x10000 = K.one << 16;         // a register, so that the key's IMAD can take the column as its immediate
    const uint32_t lut_hi = K.lut_hi;

    int rows = 0;
    bool dead = false;
    int pad = 0;
    for (int p = 0; p < npass; ++p) {
        const int c_lo = p * kPassCols;
        const int pcols = MP ? ((ncol - c_lo < kPassCols) ? ncol - c_lo : kPassCols) : ncol;
        const int nl = (pcols + C - 1) / C;
        pad = nl * C - pcols;               // only non-zero in a single-pass DP (C divides 256)
        const int last = nl - 1;
        const bool lastpass = !MP || p == npass - 1;
        int H[C], F[C];
        uint32_t sel[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int gc = c_lo + k * C + c - pad;
            uint32_t code;
            if (gc < 0 || k >= nl) code = 6u;
            else if (gc >= qlen) code = 5u;
            else {
                const uint32_t b = q[gc];
                code = b > 3u ? 4u : b;
            }
            sel[c] = code * 0x1111u + 0x8880u;
            H[c] = 0; F[c] = 0;
        }
        // a key's low half is 0xFFFF - column: the strip works with C - 1 - c and adds its own offset once per row
        const int kbase = 0xFFFF - (c_lo + k * C) - (C - 1);
        const int my_rows = k < nl ? tlen : 0;
        const bool keeper = k == last && lastpass;      // the lane that sees finished rows
        const bool feeder = MP && k == last && !lastpass;
        const uint32_t not0 = k == 0 ? 0u : 0xFFFFFFFFu;
        int hdiag = 0;
        uint32_t out_he = 0, out_key = 0;
        const int steps = tlen + nl - 1;
        uint32_t nxt4[4] = {0u, 0u, 0u, 0u}, out_lut = 0x80808080u;     // four scores of -128: nothing rises from 0
        // (16-byte loads: the scratch of a group starts on a 16-byte boundary and has kScratchSlack words past tlen)
        if (k == 0) {
            const uint4 v = *reinterpret_cast<const uint4 *>(lutw);
            nxt4[0] = v.x; nxt4[1] = v.y; nxt4[2] = v.z; nxt4[3] = v.w;
        }
        int s4 = 0;
        // four steps per stop test: a pair that has reached thr keeps going for at most three rows, which the
        // scan after the loop drops again
        for (; s4 < steps; s4 += 4) {
            if (G::any(dead)) break;
            // Lane 0 reads the rows' LUT words, one block of four steps ahead of their use; every other lane gets its
            // row's word from its left neighbour, which worked on that row one step earlier.
            uint32_t lut4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) lut4[u] = nxt4[u];
            if (k == 0) {
                const uint4 v = *reinterpret_cast<const uint4 *>(lutw + s4 + 4);
                nxt4[0] = v.x; nxt4[1] = v.y; nxt4[2] = v.z; nxt4[3] = v.w;
            }
            uint32_t key4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int s = s4 + u;
                uint32_t in_he = G::up1(out_he) & not0, in_key = G::up1(out_key) & not0;
                const uint32_t in_lut = G::up1(out_lut);
                const int i = s - k;
                const bool active = (unsigned)i < (unsigned)my_rows;
                if (MP && p > 0 && k == 0 && active) { const uint2 bv = bnd[i]; in_he = bv.x; in_key = bv.y; }
                // No branch around the cells: a lane computes every step, so the four steps are one straight line of
                // code (no register shuffling at block boundaries). Before its first row a lane sees all-negative
                // scores and zero inputs, which keeps its H, E and F at 0; past its last row it computes values that
                // only ever flow to lanes that are past their last row too.
                {
                    const uint32_t lut_lo = k == 0 ? lut4[u] : in_lut;
                    out_lut = lut_lo;
                    const int hl = (int)(in_he & 0xFFFFu);
                    int e = (int)(in_he >> 16);
                    int key = 0;
                    int diag = hdiag;
                    int kprev = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int sc = (int)prmt_sx(lut_lo, lut_hi, sel[c]);
                        const int x = SAT ? __viaddmin_s32(diag, sc, clamp) : diag * one + sc;
                        diag = H[c];
                        const int f = F[c];
                        const int h = __vimax3_s32_relu(x, e, f);
                        const int kcur = h * x10000 + (C - 1 - c);      // IMAD, FMA pipe
                        if (c & 1) key = __vimax3_s32(key, kprev, kcur);  // one VIMNMX3 per two columns
                        else if (c == C - 1) key = max(key, kcur);
                        kprev = kcur;
                        e = __viaddmax_s32_relu(h, noe_ins, e - e_ins);
                        F[c] = __viaddmax_s32_relu(h, noe_del, f - e_del);
                        H[c] = h;
                    }
                    hdiag = hl;
                    out_he = (uint32_t)H[C - 1] | ((uint32_t)e << 16);
                    key = max(key + kbase, (int)in_key);
                    out_key = (uint32_t)key;
                    key4[u] = active ? (uint32_t)key : 0u;
                    if (MP && feeder && active) bnd[i] = make_uint2(out_he, out_key);
                }
            }
            // the keeper's four finished rows: one address, up to four stores, one stop test (gmax >= thr first holds
            // on the first row that reaches thr; rows at or past tlen carry key 0)
            if (keeper) {
                const int i0 = s4 - k;
                uint32_t *rk = rowkey + i0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if ((unsigned)(i0 + u) < (unsigned)tlen) rk[u] = key4[u];
                const uint32_t m4 = max(max(key4[0], key4[1]), max(key4[2], key4[3]));
                dead |= i0 + 3 >= 0 && i0 < tlen && (int)(m4 >> 16) >= thr;   // only rows that exist can stop the pair
            }
        }
        if (MP && !lastpass) G::sync();     // the next pass's lane 0 reads what this pass's last lane wrote
        if (lastpass) {
            // the loop ran s4 steps; the keeper's last finished row is s4 - 1 - last
            rows = s4 - last;
            if (rows > tlen) rows = tlen;
            if (rows < 0) rows = 0;
        }
    }
    G::sync();                              // rowkey: one lane wrote, all lanes read
    // the stop row: the first row whose maximum reaches thr (Block II's exit, kswv.cpp:535-548); rows past it
    // were computed by the lanes that were ahead of the keeper and do not exist for the reference
    // (no H can exceed min(tlen, qlen) * match: a threshold above that is never reached and the scan is skipped --
    // phase 0 of every pair whose only threshold is the 8-bit ceiling)
    uint32_t best = 0;
    dead = false;
    if (thr <= (tlen < qlen ? tlen : qlen) * K.a) {
        uint32_t first = 0;
        for (int base = 0; base < rows; base += W) {
            const int i = base + k;
            if (i < rows && (int)(rowkey[i] >> 16) >= thr) { first = (uint32_t)(0xFFFF - i); break; }
        }
        first = G::maxu(first);
        dead = first != 0u;
        if (dead) {
            // every earlier row stayed below thr, so the stop row holds gmax and is its first row
            rows = 0xFFFF - (int)first + 1;
            best = (rowkey[rows - 1] & 0xFFFF0000u) | first;
        }
    }
    // Block II (kswv.cpp:526-548): gmax is the largest row maximum up to the stop row, te its FIRST row, qe that
    // row's first column
    if (!dead) {
        for (int base = 0; base < rows; base += W) {
            const int i = base + k;
            if (i < rows) {
                const uint32_t cand = (rowkey[i] & 0xFFFF0000u) | (uint32_t)(0xFFFF - i);
                best = cand > best ? cand : best;
            }
        }
        best = G::maxu(best);
    }
    Best B;
    B.gmax = (int)(best >> 16); B.rows = rows; B.dead = dead;
    B.te = -1; B.qe = 0;
    if (B.gmax > 0) {
        B.te = 0xFFFF - (int)(best & 0xFFFFu);
        B.qe = 0xFFFF - (int)(rowkey[B.te] & 0xFFFFu) - pad;
        if (byte) B.qe &= 255;              // the 8-bit kernel counts columns in a byte (l512, kswv.cpp:505)
    }
    return B;
}

// does the pair need the clamped (SAT) arithmetic? (host and device agree on this)
__host__ __device__ inline bool needs_sat(int a, int shift, int tlen, int qlen, bool byte) {
    return byte && (tlen < qlen ? tlen : qlen) * a >= 255 - shift;
}
__host__ __device__ inline int padded_cols(int qlen, bool byte) {
    const int quantum = byte ? 16 : 8;
    const int n = (qlen + quantum - 1) / quantum * quantum;
    return n == 0 ? quantum : n;
}

// Strip-width class of a padded query: pairs of one class run the same kswv_dp instance at every group width (the
// host's task order and the phase-1 key keep a warp's pairs in one class). Columns: <=32, 64, 80, 96, 112, 128, 160,
// then one class per further 32 columns.
__host__ __device__ inline int strip_bucket(int ncol) {
    return ncol <= 32 ? 0 : ncol <= 64 ? 1 : ncol <= 80 ? 2 : ncol <= 96 ? 3 : ncol <= 112 ? 4 : ncol <= 128 ? 5
         : ncol <= 160 ? 6 : 6 + (ncol - 160 + 31) / 32;
}

#define KSWV_DP_ARGS K, t, tlen, q, qlen, byte, thr, rowkey, bnd
template <int W>
__device__ inline Best kswv_dp_any(const KParams &K, const uint32_t *t, int tlen, const uint8_t *q, int qlen,
                                   bool byte, int thr, uint32_t *rowkey, uint2 *bnd);

// W = 32: any pair -- strips of 1..8 columns, the clamped variant where needed, several passes above 256 columns
template <>
__device__ inline Best kswv_dp_any<32>(const KParams &K, const uint32_t *t, int tlen, const uint8_t *q, int qlen,
                                       bool byte, int thr, uint32_t *rowkey, uint2 *bnd) {
    const int ncol = padded_cols(qlen, byte);
    if (ncol > kPassCols) return kswv_dp<32, 8, true, true>(KSWV_DP_ARGS);
    const bool sat = needs_sat(K.a, K.shift, tlen, qlen, byte);
#define KSWV_CASE32(CC) case CC: return sat ? kswv_dp<32, CC, true, false>(KSWV_DP_ARGS) : kswv_dp<32, CC, false, false>(KSWV_DP_ARGS);
    switch ((ncol + 31) / 32) {
        KSWV_CASE32(1) KSWV_CASE32(2) KSWV_CASE32(3) KSWV_CASE32(4) KSWV_CASE32(5) KSWV_CASE32(6) KSWV_CASE32(7)
        default: return sat ? kswv_dp<32, 8, true, false>(KSWV_DP_ARGS) : kswv_dp<32, 8, false, false>(KSWV_DP_ARGS);
    }
#undef KSWV_CASE32
}
// W = 16: pairs without clamping, up to 256 padded columns; strips of 2, 4, .. 16 columns
template <>
__device__ inline Best kswv_dp_any<16>(const KParams &K, const uint32_t *t, int tlen, const uint8_t *q, int qlen,
                                       bool byte, int thr, uint32_t *rowkey, uint2 *bnd) {
    const int ncol = padded_cols(qlen, byte);
    switch ((ncol + 31) / 32) {
        case 1: return kswv_dp<16, 2, false, false>(KSWV_DP_ARGS);
        case 2: return kswv_dp<16, 4, false, false>(KSWV_DP_ARGS);
        case 3: return kswv_dp<16, 6, false, false>(KSWV_DP_ARGS);
        case 4: return kswv_dp<16, 8, false, false>(KSWV_DP_ARGS);
        case 5: return kswv_dp<16, 10, false, false>(KSWV_DP_ARGS);
        case 6: return kswv_dp<16, 12, false, false>(KSWV_DP_ARGS);
        case 7: return kswv_dp<16, 14, false, false>(KSWV_DP_ARGS);
        default: return kswv_dp<16, 16, false, false>(KSWV_DP_ARGS);
    }
}
// W = 8: pairs without clamping, up to 160 padded columns; strips of 4, 8, 10, 12, 14, 16 or 20 columns (the multiples
// of four plus the widths of 76- and 100-bp reads; more instances cost more in instruction cache than they save in
// filler columns: with every even width the 151-bp workload ran 3 % slower)
template <>
__device__ inline Best kswv_dp_any<8>(const KParams &K, const uint32_t *t, int tlen, const uint8_t *q, int qlen,
                                      bool byte, int thr, uint32_t *rowkey, uint2 *bnd) {
    switch (strip_bucket(padded_cols(qlen, byte))) {
        case 0: return kswv_dp<8, 4, false, false>(KSWV_DP_ARGS);
        case 1: return kswv_dp<8, 8, false, false>(KSWV_DP_ARGS);
        case 2: return kswv_dp<8, 10, false, false>(KSWV_DP_ARGS);
        case 3: return kswv_dp<8, 12, false, false>(KSWV_DP_ARGS);
        case 4: return kswv_dp<8, 14, false, false>(KSWV_DP_ARGS);
        case 5: return kswv_dp<8, 16, false, false>(KSWV_DP_ARGS);
        default: return kswv_dp<8, 20, false, false>(KSWV_DP_ARGS);
    }
}
#undef KSWV_DP_ARGS

// Second best (kswv.cpp:589-703, :1139-1212) from the stored row maxima. Row i's maximum is kept when row i+1
// did not rise above it and row i-1 was not kept (Block I's mask, kswv.cpp:510-523), it reached minsc, and the
// lane was still live when the reference stored it. kept(i) = nr(i+1) & !kept(i-1) is a one-bit recurrence:
// W rows per step, the bits of one ballot word resolved in a short serial loop that every lane runs.
template <int W>
__device__ inline void kswv_second(const KParams &K, const uint32_t *rowkey, int tlen, const Best &B, bool byte,
                                   bool has_minsc, int minsc, int32_t *score2, int32_t *te2) {
    typedef Grp<W> G;
    const int k = G::lane();
    const int R = B.dead ? B.rows - 1 : tlen;           // rows whose stored maximum can be kept
    const int val = (B.gmax + K.qmax - 1) / K.qmax;
    const int low = (int16_t)(B.te - val), high = (int16_t)(B.te + val);
    const uint32_t off = byte ? 0u : 1u;
    uint32_t best = 0;
    if (has_minsc) {
        uint32_t carry = 0;
        for (int base = 0; base < R; base += W) {
            const int i = base + k;
            const int cur = i < R ? (int)(rowkey[i] >> 16) : 0;
            const int nxt = (i + 1 < B.rows && i < R) ? (int)(rowkey[i + 1] >> 16) : 0;   // past the last row: not rising
            const uint32_t N = G::ballot(i < R && !(nxt > cur));
            uint32_t X = 0, prev = carry;
            for (int b = 0; b < W; ++b) {
                const uint32_t bit = (N >> b) & 1u & (prev ^ 1u);
                X |= bit << b;
                prev = bit;
            }
            carry = prev;
            const bool kept = ((X >> k) & 1u) != 0u && cur >= minsc && (i < low || i > high);
            const uint32_t w = (uint32_t)cur + off;
            if (kept && w > 0u) {
                const uint32_t key = (w << 16) | (uint32_t)(0xFFFF - i);
                best = key > best ? key : best;
            }
        }
    }
    best = G::maxu(best);
    if (best == 0u) { *score2 = -1; *te2 = -1; }
    else { *score2 = (int)(best >> 16) - (int)off; *te2 = 0xFFFF - (int)(best & 0xFFFFu); }
}

// The two phases of a pair are separate kernels: phase 1's trip count (te + 1 rows, qe + 1 columns) is only known
// after phase 0, and pairs that share a warp should have similar trip counts, so the phase-1 tasks are re-ordered on
// the device in between (by strip width and te, a radix sort on p1key).

struct Thresholds { bool byte, has_minsc; int minsc, thr, sat; };
__device__ inline Thresholds kswv_thresholds(const KParams &K, int xtra) {
    Thresholds t;
    t.byte = (xtra & kXByte) != 0;
    const int lim = t.byte ? 255 : 32767;
    int v = (xtra & kXSubo) ? (xtra & 0xffff) : 0x10000;                 // kswv.cpp:422-437, :976-993
    t.has_minsc = v <= lim;
    t.minsc = v;
    v = (xtra & kXStop) ? (xtra & 0xffff) : 0x10000;
    t.thr = v <= lim ? v : kNoStop;
    t.sat = t.byte ? 255 - K.shift : kNoStop;                           // adds_epu8(gmax, shift) == 255, kswv.cpp:539-540
    if (t.sat < t.thr) t.thr = t.sat;
    return t;
}

// per reference row the four scores of its base (LUT word): built once per phase by all lanes of the group, so that
// a step of the row loop loads one word instead of a base and five instructions of LUT arithmetic. The first rt
// rows are read backwards (phase 1: revseq of the aligned prefix, bwamem_pair.cpp:673, :691; len1 is unchanged).
template <int W>
__device__ inline void row_luts(const KParams &K, const uint8_t *t, int tlen, int rt, uint32_t *lutw) {
    for (int i = Grp<W>::lane(); i < tlen; i += W) {
        const uint32_t b = t[i < rt ? rt - 1 - i : i];
        lutw[i] = b > 3u ? K.lut_amb : (K.lut_mis ^ (K.lut_ab << (8u * b)));
    }
}

// Phase 0 of one pair on one group of W lanes: score, te, qe, score2, te2 (tb = qb = -1). *p1key = 0 if the pair has
// no phase 1 (bwamem_pair.cpp:667, :685), else (strip-width bucket of phase 1 << 16) | rows of the reversed prefix.
template <int W>
__device__ inline Result kswv_phase0(const KParams &K, const Task &T, const uint8_t *ref, const uint8_t *qer,
                                     uint32_t *rowkey, uint2 *bnd, uint32_t *lutw, uint32_t *p1key) {
    const Thresholds th = kswv_thresholds(K, T.xtra);
    const uint8_t *t = ref + T.roff, *q = qer + T.qoff;
    row_luts<W>(K, t, T.tlen, 0, lutw);
    Grp<W>::sync();
    Result r;
    const Best B = kswv_dp_any<W>(K, lutw, T.tlen, q, T.qlen, th.byte, th.thr, rowkey, bnd);
    r.score = th.byte ? (B.gmax + K.shift < 255 ? B.gmax : 255) : B.gmax;   // kswv.cpp:568
    r.te = B.te; r.qe = B.qe;
    r.tb = r.qb = -1;
    if (th.byte && r.score == 255) { r.score2 = -1; r.te2 = -1; }
    else kswv_second<W>(K, rowkey, T.tlen, B, th.byte, th.has_minsc, th.minsc, &r.score2, &r.te2);
    const bool phase1 = (T.xtra & kXStart) && !((T.xtra & kXSubo) && r.score < (T.xtra & 0xffff));
    *p1key = phase1 ? ((uint32_t)(strip_bucket(padded_cols(r.qe + 1, th.byte)) + 1) << 16) | (uint32_t)(r.te + 1) : 0u;
    Grp<W>::sync();                                                      // the scratch is free for the next pair
    return r;
}

// Phase 1 (bwamem_pair.cpp:660-699): the prefixes ref[0..te], qer[0..qe] reversed, h0 = KSW_XSTOP | score, same class.
template <int W>
__device__ inline void kswv_phase1(const KParams &K, const Task &T, const uint8_t *ref, const uint8_t *qer,
                                   uint32_t *rowkey, uint2 *bnd, uint32_t *lutw, uint8_t *qbuf, Result *r) {
    const Thresholds th = kswv_thresholds(K, T.xtra);
    const uint8_t *t = ref + T.roff, *q = qer + T.qoff;
    const int rt = r->te + 1, q1 = r->qe + 1;
    row_luts<W>(K, t, T.tlen, rt, lutw);
    for (int j = Grp<W>::lane(); j < q1; j += W) qbuf[j] = q[q1 - 1 - j];
    Grp<W>::sync();
    int thr1 = r->score;
    if (th.sat < thr1) thr1 = th.sat;
    const Best V = kswv_dp_any<W>(K, lutw, T.tlen, qbuf, q1, th.byte, thr1, rowkey, bnd);
    if (r->score == V.gmax) { r->tb = r->te - V.te; r->qb = r->qe - V.qe; }
    Grp<W>::sync();                                                      // the scratch is free for the next pair
}

#ifndef BSW_HOST_EMUL
constexpr int kKswvWarps = 4;       // warps per block

// Persistent warps: each takes the next 32 / W tasks (the host orders them by decreasing size, equal strip widths
// together) until none is left. Scratch (row keys, LUT words, reversed query, boundary column) is per group.
template <int W>
__global__ void __launch_bounds__(kKswvWarps * 32, 4)
kswv_phase0_kernel(const KParams K, const Task *__restrict__ tasks, int ntasks, const uint8_t *__restrict__ ref,
                   const uint8_t *__restrict__ qer, Result *__restrict__ out, uint32_t *__restrict__ p1key,
                   uint32_t *__restrict__ p1val, uint32_t *rowkey_all, uint2 *bnd_all, uint32_t *lutw_all,
                   int scratch_rows, int *counter) {
    constexpr int kGroups = 32 / W;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int grp = warp * kGroups + (hw_lane() / W);
    uint32_t *rowkey = rowkey_all + (size_t)grp * scratch_rows;
    uint2 *bnd = bnd_all ? bnd_all + (size_t)grp * scratch_rows : nullptr;
    uint32_t *lutw = lutw_all + (size_t)grp * scratch_rows;
    for (;;) {
        int id = 0;
        if (hw_lane() == 0) id = atomicAdd(counter, kGroups);
        id = __shfl_sync(0xFFFFFFFFu, id, 0);
        if (id >= ntasks) break;
        const int mine = id + hw_lane() / W;
        if (mine < ntasks) {
            const Task T = tasks[mine];
            uint32_t key;
            const Result r = kswv_phase0<W>(K, T, ref, qer, rowkey, bnd, lutw, &key);
            if (Grp<W>::lane() == 0) { out[T.out] = r; p1key[mine] = key; p1val[mine] = (uint32_t)mine; }
        }
    }
}

// Phase 1 over the re-ordered tasks: order[j] is a task index, key[j] == 0 ends the list.
template <int W>
__global__ void __launch_bounds__(kKswvWarps * 32, 4)
kswv_phase1_kernel(const KParams K, const Task *__restrict__ tasks, int ntasks, const uint32_t *__restrict__ key,
                   const uint32_t *__restrict__ order, const uint8_t *__restrict__ ref,
                   const uint8_t *__restrict__ qer, Result *out, uint32_t *rowkey_all, uint2 *bnd_all,
                   uint32_t *lutw_all, uint8_t *qbuf_all, int scratch_rows, int scratch_q, int *counter) {
    constexpr int kGroups = 32 / W;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int grp = warp * kGroups + (hw_lane() / W);
    uint32_t *rowkey = rowkey_all + (size_t)grp * scratch_rows;
    uint2 *bnd = bnd_all ? bnd_all + (size_t)grp * scratch_rows : nullptr;
    uint32_t *lutw = lutw_all + (size_t)grp * scratch_rows;
    uint8_t *qbuf = qbuf_all + (size_t)grp * scratch_q;
    for (;;) {
        int id = 0;
        if (hw_lane() == 0) id = atomicAdd(counter, kGroups);
        id = __shfl_sync(0xFFFFFFFFu, id, 0);
        if (id >= ntasks || key[id] == 0u) break;
        const int mine = id + hw_lane() / W;
        if (mine < ntasks && key[mine] != 0u) {
            const Task T = tasks[order[mine]];
            Result r = out[T.out];
            kswv_phase1<W>(K, T, ref, qer, rowkey, bnd, lutw, qbuf, &r);
            if (Grp<W>::lane() == 0) { out[T.out].tb = r.tb; out[T.out].qb = r.qb; }
        }
    }
}
#endif

}  // namespace kswvk
