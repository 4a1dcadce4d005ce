// Host-side sequence packers of the bsw pipeline: 2 bits per base for plain pairs, 4 bits per base
// for pairs that hold an ambiguous base (code 4). Layout is what bswk::unpack_pair expects.
#pragma once
#include <stdint.h>
#include <string.h>

namespace bswk {

// Packs len base codes 2 bits each into dst (seq_bytes(len,false) bytes, zero padded).
// Returns true if a code >= 4 was seen (the caller then re-packs the pair wide).
inline bool pack2bit(const uint8_t *src, int len, uint8_t *dst) {
    const int nbytes = (int)seq_bytes((uint32_t)len, false);
    int i = 0, o = 0;
    uint64_t bad = 0;
    for (; i + 8 <= len; i += 8, o += 2) {
        uint64_t v;
        memcpy(&v, src + i, 8);
        bad |= v;
        uint64_t t = v | (v >> 6) | (v >> 12) | (v >> 18);
        dst[o] = (uint8_t)t;
        dst[o + 1] = (uint8_t)(t >> 32);
    }
    if (i < len) {
        uint8_t tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        memcpy(tmp, src + i, (size_t)(len - i));
        uint64_t v;
        memcpy(&v, tmp, 8);
        bad |= v;
        uint64_t t = v | (v >> 6) | (v >> 12) | (v >> 18);
        dst[o++] = (uint8_t)t;
        if (len - i > 4) dst[o++] = (uint8_t)(t >> 32);
    }
    for (; o < nbytes; ++o) dst[o] = 0;
    return (bad & 0xFCFCFCFCFCFCFCFCull) != 0;
}

// BMI2 variant: PEXT gathers the 2 low bits of 8 bases in one instruction. Selected at run time.
inline bool pack_have_pext() {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool v = __builtin_cpu_supports("bmi2");
    return v;
#else
    return false;
#endif
}
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi2")))
inline bool pack2bit_pext(const uint8_t *src, int len, uint8_t *dst) {
    const int nbytes = (int)seq_bytes((uint32_t)len, false);
    const uint64_t M = 0x0303030303030303ull;
    int i = 0, o = 0;
    uint64_t bad = 0;
    for (; i + 32 <= len; i += 32, o += 8) {
        uint64_t a, b, c, d;
        memcpy(&a, src + i, 8); memcpy(&b, src + i + 8, 8);
        memcpy(&c, src + i + 16, 8); memcpy(&d, src + i + 24, 8);
        bad |= a | b | c | d;
        const uint64_t r = __builtin_ia32_pext_di(a, M) | (__builtin_ia32_pext_di(b, M) << 16) |
                           (__builtin_ia32_pext_di(c, M) << 32) | (__builtin_ia32_pext_di(d, M) << 48);
        memcpy(dst + o, &r, 8);
    }
    for (; i + 8 <= len; i += 8, o += 2) {
        uint64_t v;
        memcpy(&v, src + i, 8);
        bad |= v;
        const uint16_t r = (uint16_t)__builtin_ia32_pext_di(v, M);
        memcpy(dst + o, &r, 2);
    }
    if (i < len) {
        uint8_t tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        memcpy(tmp, src + i, (size_t)(len - i));
        uint64_t v;
        memcpy(&v, tmp, 8);
        bad |= v;
        const uint16_t r = (uint16_t)__builtin_ia32_pext_di(v, M);
        dst[o++] = (uint8_t)r;
        if (len - i > 4) dst[o++] = (uint8_t)(r >> 8);
    }
    for (; o < nbytes; ++o) dst[o] = 0;
    return (bad & 0xFCFCFCFCFCFCFCFCull) != 0;
}
#else
inline bool pack2bit_pext(const uint8_t *src, int len, uint8_t *dst) { return pack2bit(src, len, dst); }
#endif

// AVX2 variant: 32 bases per step. maddubs folds base pairs (b0 + 4*b1), madd folds those again
// (+ 16*(b2 + 4*b3)), two saturating packs bring the eight 32-bit results down to eight bytes. The tail
// is loaded in one piece when that cannot cross a page, through a bounce buffer otherwise.
inline bool pack_have_avx2() {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
#else
    return false;
#endif
}
#if defined(__x86_64__) && defined(__GNUC__)
}  // namespace bswk
#include <immintrin.h>
namespace bswk {
__attribute__((target("avx2")))
inline uint64_t pack_fold32(__m256i v, __m256i &bad) {
    const __m256i w8 = _mm256_set1_epi16(0x0401), w16 = _mm256_set1_epi32(0x00100001);
    bad = _mm256_or_si256(bad, v);
    const __m256i u = _mm256_madd_epi16(_mm256_maddubs_epi16(v, w8), w16);
    const __m128i p16 = _mm_packus_epi32(_mm256_castsi256_si128(u), _mm256_extracti128_si256(u, 1));
    return (uint64_t)_mm_cvtsi128_si64(_mm_packus_epi16(p16, p16));
}
__attribute__((target("avx2")))
inline bool pack2bit_avx2(const uint8_t *src, int len, uint8_t *dst) {
    const int nbytes = (int)seq_bytes((uint32_t)len, false);
    __m256i bad = _mm256_setzero_si256();
    int i = 0, o = 0;
    for (; i + 32 <= len; i += 32, o += 8) {
        const uint64_t r = pack_fold32(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)), bad);
        memcpy(dst + o, &r, 8);
    }
    if (i < len) {
        const int rem = len - i;
        __m256i v;
        if (((uintptr_t)(src + i) & 4095u) <= 4096u - 32u) {
            const __m256i iota = _mm256_setr_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19,
                                                  20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31);
            const __m256i keep = _mm256_cmpgt_epi8(_mm256_set1_epi8((char)rem), iota);
            v = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)), keep);
        } else {
            alignas(32) uint8_t tmp[32] = {0};
            memcpy(tmp, src + i, (size_t)rem);
            v = _mm256_load_si256(reinterpret_cast<const __m256i *>(tmp));
        }
        const uint64_t r = pack_fold32(v, bad);
        if (nbytes - o >= 8) { memcpy(dst + o, &r, 8); o += 8; }
        else { const uint32_t r4 = (uint32_t)r; memcpy(dst + o, &r4, 4); o += 4; }
    }
    for (; o < nbytes; ++o) dst[o] = 0;
    return !_mm256_testz_si256(bad, _mm256_set1_epi8((char)0xFC));
}
// One pair in one call: query at dst, target at dst + qb (qb = the query's padded bytes). Every step stores
// eight bytes, so up to FOUR bytes past a sequence's padded length are written (zeros: the masked lanes
// pack to zero): the query's spill lands where the target is packed next, the target's in the slot's own
// padding or -- at most one word -- in what follows the slot, which the caller therefore must own (the
// thread's next slot, or the spare word it keeps at the end of its arena). No zero-fill loop, one
// ambiguity test per pair. Returns true when either sequence holds a base > 3.
__attribute__((target("avx2")))
inline void pack2bit_avx2_steps(const uint8_t *src, int len, uint8_t *dst, __m256i &bad) {
    int i = 0, o = 0;
    for (; i + 32 <= len; i += 32, o += 8) {
        const uint64_t r = pack_fold32(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)), bad);
        memcpy(dst + o, &r, 8);
    }
    if (i < len) {
        const int rem = len - i;
        __m256i v;
        if (((uintptr_t)(src + i) & 4095u) <= 4096u - 32u) {
            const __m256i iota = _mm256_setr_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19,
                                                  20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31);
            const __m256i keep = _mm256_cmpgt_epi8(_mm256_set1_epi8((char)rem), iota);
            v = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)), keep);
        } else {
            alignas(32) uint8_t tmp[32] = {0};
            memcpy(tmp, src + i, (size_t)rem);
            v = _mm256_load_si256(reinterpret_cast<const __m256i *>(tmp));
        }
        const uint64_t r = pack_fold32(v, bad);
        memcpy(dst + o, &r, 8);
    }
}
__attribute__((target("avx2")))
inline bool pack_pair_avx2(const uint8_t *q, int qlen, const uint8_t *t, int tlen, uint8_t *dst, uint32_t qb) {
    __m256i bad = _mm256_setzero_si256();
    pack2bit_avx2_steps(q, qlen, dst, bad);
    pack2bit_avx2_steps(t, tlen, dst + qb, bad);
    return !_mm256_testz_si256(bad, _mm256_set1_epi8((char)0xFC));
}
#else
inline bool pack2bit_avx2(const uint8_t *src, int len, uint8_t *dst) { return pack2bit(src, len, dst); }
inline bool pack_pair_avx2(const uint8_t *q, int qlen, const uint8_t *t, int tlen, uint8_t *dst, uint32_t qb) {
    const bool a = pack2bit(q, qlen, dst);
    const bool b = pack2bit(t, tlen, dst + qb);
    return a || b;
}
#endif

inline void pack4bit(const uint8_t *src, int len, uint8_t *dst) {
    const int nbytes = (int)seq_bytes((uint32_t)len, true);
    int o = 0;
    for (int i = 0; i < len; i += 2, ++o) {
        uint8_t a = src[i] > 4 ? 4 : src[i];
        uint8_t b = (i + 1 < len) ? (src[i + 1] > 4 ? 4 : src[i + 1]) : 0;
        dst[o] = (uint8_t)(a | (b << 4));
    }
    for (; o < nbytes; ++o) dst[o] = 0;
}


}  // namespace bswk
