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
