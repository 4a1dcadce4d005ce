// TEST INFRASTRUCTURE: host emulation of the CUDA intrinsics used by the per-pair device code in
// genarchbench_b200/csrc/bsw_kernels.cuh, so that code can be compiled with g++ and compared with the
// oracle on a machine without a GPU. Semantics follow the CUDA Math API (SIMD intrinsics) and PTX prmt.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __align__(n) alignas(n)
#define __restrict__

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
using std::max;
using std::min;

namespace emul {
static inline int16_t lo(uint32_t v) { return (int16_t)(v & 0xFFFFu); }
static inline int16_t hi(uint32_t v) { return (int16_t)(v >> 16); }
static inline uint32_t pk(int lo_, int hi_) { return ((uint32_t)lo_ & 0xFFFFu) | (((uint32_t)hi_ & 0xFFFFu) << 16); }
static inline int wrap16(int v) { return (int16_t)(uint16_t)v; }
}  // namespace emul

static inline uint32_t __vmins2(uint32_t a, uint32_t b) {
    return emul::pk(std::min(emul::lo(a), emul::lo(b)), std::min(emul::hi(a), emul::hi(b)));
}
static inline uint32_t __vmaxs2(uint32_t a, uint32_t b) {
    return emul::pk(std::max(emul::lo(a), emul::lo(b)), std::max(emul::hi(a), emul::hi(b)));
}
static inline uint32_t __vadd2(uint32_t a, uint32_t b) {  // per-halfword wrapping add
    return emul::pk(emul::lo(a) + emul::lo(b), emul::hi(a) + emul::hi(b));
}
static inline uint32_t __viaddmax_s16x2(uint32_t a, uint32_t b, uint32_t c) {
    int l = std::max(emul::wrap16(emul::lo(a) + emul::lo(b)), (int)emul::lo(c));
    int h = std::max(emul::wrap16(emul::hi(a) + emul::hi(b)), (int)emul::hi(c));
    return emul::pk(l, h);
}
static inline uint32_t __viaddmax_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) {
    int l = std::max(std::max(emul::wrap16(emul::lo(a) + emul::lo(b)), (int)emul::lo(c)), 0);
    int h = std::max(std::max(emul::wrap16(emul::hi(a) + emul::hi(b)), (int)emul::hi(c)), 0);
    return emul::pk(l, h);
}
static inline uint32_t __viaddmin_s16x2(uint32_t a, uint32_t b, uint32_t c) {
    int l = std::min(emul::wrap16(emul::lo(a) + emul::lo(b)), (int)emul::lo(c));
    int h = std::min(emul::wrap16(emul::hi(a) + emul::hi(b)), (int)emul::hi(c));
    return emul::pk(l, h);
}
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b) {
    return std::max(a & 0xFFFFu, b & 0xFFFFu) | (std::max(a >> 16, b >> 16) << 16);
}
static inline uint32_t __vimax3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxu2(__vmaxu2(a, b), c); }
static inline uint32_t __viaddmax_u16x2(uint32_t a, uint32_t b, uint32_t c) {   // max(a + b, c), wrapping add
    const uint32_t s = (((a & 0xFFFFu) + (b & 0xFFFFu)) & 0xFFFFu) | ((((a >> 16) + (b >> 16)) & 0xFFFFu) << 16);
    return __vmaxu2(s, c);
}
static inline int __viaddmax_s32(int a, int b, int c) { return std::max((int)((uint32_t)a + (uint32_t)b), c); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __vimax3_s16x2(uint32_t a, uint32_t b, uint32_t c) {
    return emul::pk(std::max({emul::lo(a), emul::lo(b), emul::lo(c)}),
                    std::max({emul::hi(a), emul::hi(b), emul::hi(c)}));
}
// max(a, b) per halfword; *pred = (a >= b)
static inline uint32_t __vibmax_s16x2(uint32_t a, uint32_t b, bool *pred_hi, bool *pred_lo) {
    *pred_lo = emul::lo(a) >= emul::lo(b);
    *pred_hi = emul::hi(a) >= emul::hi(b);
    return emul::pk(std::max(emul::lo(a), emul::lo(b)), std::max(emul::hi(a), emul::hi(b)));
}
// PTX prmt.b32 default mode: byte i of the result is byte (sel nibble i & 7) of {y:x}; nibble bit 3
// replicates that byte's sign bit instead.
namespace emul {
static inline uint32_t prmt(uint32_t x, uint32_t y, uint32_t s) {
    uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t nib = (s >> (4 * i)) & 0xFu;
        uint32_t b = (uint32_t)(src >> (8 * (nib & 7u))) & 0xFFu;
        if (nib & 8u) b = (b & 0x80u) ? 0xFFu : 0x00u;
        r |= b << (8 * i);
    }
    return r;
}
}  // namespace emul
// CUDA's __byte_perm ignores bit 3 of every selector nibble (plain byte copy only)
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) { return emul::prmt(x, y, s & 0x7777u); }
static inline uint32_t __funnelshift_r(uint32_t lo_, uint32_t hi_, uint32_t sh) {
    uint64_t v = ((uint64_t)hi_ << 32) | lo_;
    return (uint32_t)(v >> (sh & 31u));
}
