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
