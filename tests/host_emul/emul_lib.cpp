// TEST INFRASTRUCTURE: runs the product's per-pair device code (unpack_pair + extend_pair from
// genarchbench_b200/csrc/bsw_kernels.cuh) and host packers (bsw_pack.h) on the CPU through the
// intrinsic emulation in dpx_host_emul.h. Used only by tests/test_host_emulation.py to check the
// kernel's ALGORITHM against the oracle where no GPU exists; it is not a product path.
#define BSW_HOST_EMUL 1
#include "bsw_kernels.cuh"
#include "bsw_pack.h"
#include "bsw_types.h"
#include <vector>

using namespace bswk;

// Two pairs per thread (extend_duo2, bsw_duo.cuh): consecutive pairs of the caller's order
// share a "thread". key = true: threads whose scores and column indices fit the 16-bit key run the KEY variant
// (seqid = -1 marks those pairs).
extern "C" int bsw_emul_batch_duo2(const bsw_params *p, bsw_seqpair *pairs, const uint8_t *ref,
                                   const uint8_t *qer, int64_t n, int32_t w, int32_t key) {
    const KParams K0{p->o_del, p->e_del, p->o_ins, p->e_ins, p->zdrop, p->end_bonus, p->match, p->mismatch, p->ambig, w,
                     max_score_of(p->match, p->mismatch, p->ambig), 65536u, (uint32_t)(p->match + 1), 1u, 0u, 0u};
    const bool sym = p->o_del == p->o_ins && p->e_del == p->e_ins;
#pragma omp parallel
    {
        std::vector<uint4> he;
        std::vector<uint2> qs;
        std::vector<uint32_t> blob[2];
#pragma omp for schedule(dynamic, 128)
        for (int64_t k = 0; k < n; k += 2) {
            DuoIn L[2];
            bool anywide = false, fast = true;
            int qmax = 0, maxsc = 0;
            for (int a = 0; a < 2; ++a) {
                const bool has = k + a < n;
                bsw_seqpair sp = has ? pairs[k + a] : bsw_seqpair{};
                L[a].qlen = has ? sp.len2 : 0; L[a].tlen = has ? sp.len1 : 0; L[a].h0 = has ? sp.h0 : 0;
                blob[a].assign((size_t)(seq_bytes(sp.len2, true) + seq_bytes(sp.len1, true)) / 4 + 4, 0);
                uint8_t *b = reinterpret_cast<uint8_t *>(blob[a].data());
                bool wide = false;
                if (has) {
                    wide = pack2bit(qer + sp.idq, sp.len2, b);
                    wide |= pack2bit(ref + sp.idr, sp.len1, b + seq_bytes(sp.len2, false));
                    if (wide) {
                        pack4bit(qer + sp.idq, sp.len2, b);
                        pack4bit(ref + sp.idr, sp.len1, b + seq_bytes(sp.len2, true));
                    }
                    fast = fast && (int64_t)(sp.h0 + sp.len2 * p->match) * (p->match + 1) <= 32767;
                    if (sp.len1 > 0 && sp.len2 > 0) {
                        qmax = std::max(qmax, sp.len2);
                        maxsc = std::max(maxsc, sp.h0 + sp.len2 * p->match);
                    }
                }
                L[a].wide = wide;
                L[a].blob = blob[a].data();
                anywide |= wide;
            }
            if (getenv("BSW_EMUL_SLOWM")) fast = false;
            // exactly what the kernel allocates per thread; poisoned so that any read of an entry that was
            // never initialised shows up as a mismatch
            he.assign((size_t)2 * duo_blocks(qmax), uint4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
            qs.assign((size_t)duo_blocks(qmax), uint2{0xDEADBEEFu, 0xDEADBEEFu});
            RowsD R{he.data(), qs.data(), 1, 0u};
            PairResult r[2];
            KParams K = K0;
            const int kbits = bits_for((uint32_t)(4 * duo_blocks(qmax) - 1));   // masked blocks index up to the block end
            const bool keyed = key && fast && maxsc < (1 << (16 - kbits));
            if (keyed) { K.kbits = (uint32_t)kbits; K.kkey = 1u << kbits; }
#define ED2(F, S, KY) (anywide ? extend_duo2<F, S, true, KY>(R, L, K, r) : extend_duo2<F, S, false, KY>(R, L, K, r))
            if (keyed) { if (sym) ED2(true, true, true); else ED2(true, false, true); }
            else if (fast) { if (sym) ED2(true, true, false); else ED2(true, false, false); }
            else { if (sym) ED2(false, true, false); else ED2(false, false, false); }
#undef ED2
            for (int a = 0; a < 2 && k + a < n; ++a) {
                bsw_seqpair &sp = pairs[k + a];
                sp.seqid = keyed ? -1 : 0;
                if (sp.len1 == 0 || sp.len2 == 0) {
                    sp.score = sp.h0; sp.qle = sp.tle = sp.gtle = 0; sp.gscore = -1; sp.max_off = 0;
                    continue;
                }
                sp.score = r[a].score; sp.qle = r[a].qle; sp.tle = r[a].tle; sp.gtle = r[a].gtle;
                sp.gscore = r[a].gscore; sp.max_off = r[a].max_off;
            }
        }
    }
    return 0;
}

extern "C" int bsw_emul_batch_duo(const bsw_params *p, bsw_seqpair *pairs, const uint8_t *ref,
                                  const uint8_t *qer, int64_t n, int32_t w) {
    return bsw_emul_batch_duo2(p, pairs, ref, qer, n, w, 0);
}
extern "C" int bsw_emul_batch_duo_key(const bsw_params *p, bsw_seqpair *pairs, const uint8_t *ref,
                                      const uint8_t *qer, int64_t n, int32_t w) {
    return bsw_emul_batch_duo2(p, pairs, ref, qer, n, w, 1);
}

// key = true: pairs whose scores and group indices fit the 16-bit key run extend_pair<.., KEY> (with the
// tightest index width their own query allows, so the packing is exercised at its limits).
static int emul_batch_impl(const bsw_params *p, bsw_seqpair *pairs, const uint8_t *ref,
                           const uint8_t *qer, int64_t n, int32_t w, bool key) {
    const KParams K0{p->o_del, p->e_del, p->o_ins, p->e_ins, p->zdrop, p->end_bonus, p->match, p->mismatch, p->ambig, w,
                     max_score_of(p->match, p->mismatch, p->ambig), 65536u, (uint32_t)(p->match + 1), 1u, 0u, 0u};
    const bool sym = p->o_del == p->o_ins && p->e_del == p->e_ins;
#pragma omp parallel
    {
        std::vector<uint4> he;
        std::vector<uint32_t> qs, blob;
#pragma omp for schedule(dynamic, 256)
        for (int64_t k = 0; k < n; ++k) {
            bsw_seqpair &sp = pairs[k];
            if (sp.len1 == 0 || sp.len2 == 0) {
                sp.score = sp.h0; sp.qle = sp.tle = sp.gtle = 0; sp.gscore = -1; sp.max_off = 0;
                continue;
            }
            he.assign((size_t)row_elems(sp.len2), uint4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
            qs.assign((size_t)sel_words(sp.len2), 0xDEADBEEFu);
            blob.assign((size_t)(seq_bytes(sp.len2, true) + seq_bytes(sp.len1, true)) / 4 + 4, 0);
            uint8_t *b = reinterpret_cast<uint8_t *>(blob.data());
            bool wide = pack2bit(qer + sp.idq, sp.len2, b);
            wide |= pack2bit(ref + sp.idr, sp.len1, b + seq_bytes(sp.len2, false));
            if (wide) {
                pack4bit(qer + sp.idq, sp.len2, b);
                pack4bit(ref + sp.idr, sp.len1, b + seq_bytes(sp.len2, true));
            }
            Rows R{he.data(), qs.data(), nullptr, 1, -1, nullptr};
            if (wide) unpack_pair<true>(blob.data(), sp.len2, R);
            else unpack_pair<false>(blob.data(), sp.len2, R);
            PairResult r;
            const bool m1 = (int64_t)(sp.h0 + sp.len2 * p->match) * (p->match + 1) <= 32767 && !getenv("BSW_EMUL_SLOWM");
            KParams K = K0;
            const int kbits = bits_for((uint32_t)std::min<int64_t>((sp.len2 - 1) >> 1, BSW_KEY_REL ? (int64_t)w + 2 : INT32_MAX));
            if (key && m1 && sp.h0 + sp.len2 * p->match < (1 << (16 - kbits))) {
                K.kbits = (uint32_t)kbits; K.kkey = 1u << kbits;
#define EK(S) (wide ? extend_pair<true, S, false, true, false, 4, true>(R, sp.len2, sp.len1, sp.h0, K) \
               : extend_pair<true, S, false, false, false, 4, true>(R, sp.len2, sp.len1, sp.h0, K))
                r = sym ? EK(true) : EK(false);
#undef EK
                r.cells = 0xFFFFFFFFu;
            } else {
#define EP(F, S) (wide ? extend_pair<F, S, true, true>(R, sp.len2, sp.len1, sp.h0, K) \
                  : extend_pair<F, S, true, false>(R, sp.len2, sp.len1, sp.h0, K))
                if (m1) r = sym ? EP(true, true) : EP(true, false);
                else r = sym ? EP(false, true) : EP(false, false);
#undef EP
            }
            sp.score = r.score; sp.qle = r.qle; sp.tle = r.tle; sp.gtle = r.gtle;
            sp.gscore = r.gscore; sp.max_off = r.max_off;
            sp.seqid = (int32_t)r.cells;   // test hook: cell count of the COUNT variant
        }
    }
    return 0;
}

extern "C" int bsw_emul_batch(const bsw_params *p, bsw_seqpair *pairs, const uint8_t *ref,
                              const uint8_t *qer, int64_t n, int32_t w) {
    return emul_batch_impl(p, pairs, ref, qer, n, w, false);
}
// returns in seqid 0xFFFFFFFF (-1) for the pairs that took the keyed path
extern "C" int bsw_emul_batch_key(const bsw_params *p, bsw_seqpair *pairs, const uint8_t *ref,
                                  const uint8_t *qer, int64_t n, int32_t w) {
    return emul_batch_impl(p, pairs, ref, qer, n, w, true);
}

// Windowed rows (extend_pair<.., WIN>): every pair runs with the window a band of w needs.
extern "C" int bsw_emul_batch_win(const bsw_params *p, bsw_seqpair *pairs, const uint8_t *ref,
                              const uint8_t *qer, int64_t n, int32_t w) {
    KParams K{p->o_del, p->e_del, p->o_ins, p->e_ins, p->zdrop, p->end_bonus, p->match, p->mismatch, p->ambig, w,
              max_score_of(p->match, p->mismatch, p->ambig), 65536u, (uint32_t)(p->match + 1), 1u};
    const bool sym = p->o_del == p->o_ins && p->e_del == p->e_ins;
#pragma omp parallel
    {
        std::vector<uint4> he;
        std::vector<uint32_t> qs, blob;
#pragma omp for schedule(dynamic, 256)
        for (int64_t k = 0; k < n; ++k) {
            bsw_seqpair &sp = pairs[k];
            if (sp.len1 == 0 || sp.len2 == 0) {
                sp.score = sp.h0; sp.qle = sp.tle = sp.gtle = 0; sp.gscore = -1; sp.max_off = 0;
                continue;
            }
            he.assign((size_t)row_elems(sp.len2), uint4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
            qs.assign((size_t)sel_words(sp.len2), 0xDEADBEEFu);
            blob.assign((size_t)(seq_bytes(sp.len2, true) + seq_bytes(sp.len1, true)) / 4 + 4, 0);
            uint8_t *b = reinterpret_cast<uint8_t *>(blob.data());
            bool wide = pack2bit(qer + sp.idq, sp.len2, b);
            wide |= pack2bit(ref + sp.idr, sp.len1, b + seq_bytes(sp.len2, false));
            if (wide) {
                pack4bit(qer + sp.idq, sp.len2, b);
                pack4bit(ref + sp.idr, sp.len1, b + seq_bytes(sp.len2, true));
            }
            const int nk = window_elems(w);
            he.assign((size_t)nk, uint4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
            qs.assign((size_t)nk, 0xDEADBEEFu);
            Rows R{he.data(), qs.data(), nullptr, 1, nk - 1, nullptr};
            if (wide) unpack_pair<true>(blob.data(), sp.len2, R);
            else unpack_pair<false>(blob.data(), sp.len2, R);
            PairResult r;
            const bool m1 = (int64_t)(sp.h0 + sp.len2 * p->match) * (p->match + 1) <= 32767 && !getenv("BSW_EMUL_SLOWM");
#define EP(F, S) (wide ? extend_pair<F, S, true, true, true, 8>(R, sp.len2, sp.len1, sp.h0, K) \
                  : extend_pair<F, S, true, false, true, 8>(R, sp.len2, sp.len1, sp.h0, K))
            if (m1) r = sym ? EP(true, true) : EP(true, false);
            else r = sym ? EP(false, true) : EP(false, false);
#undef EP
            sp.score = r.score; sp.qle = r.qle; sp.tle = r.tle; sp.gtle = r.gtle;
            sp.gscore = r.gscore; sp.max_off = r.max_off;
            sp.seqid = (int32_t)r.cells;   // test hook: cell count of the COUNT variant
        }
    }
    return 0;
}

// The host packers against the scalar one: random lengths and contents (ambiguous bases included), sources
// placed so that tails end right at a page boundary, destination buffers of exactly slot + one spare word
// (what the product's arenas guarantee), so the AddressSanitizer build sees any further overrun.
// Returns the number of mismatching pairs.
#include <random>
#include <sys/mman.h>
extern "C" int64_t bsw_emul_pack_check(int64_t n, uint32_t seed) {
    if (!pack_have_avx2() || !pack_have_pext()) return 0;   // nothing to compare on such a host
    std::mt19937 rng(seed);
    const size_t page = 4096;
    uint8_t *area = static_cast<uint8_t *>(mmap(nullptr, 3 * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    if (area == MAP_FAILED) return -1;
    mprotect(area + 2 * page, page, PROT_NONE);   // reads past the end of the second page fault
    int64_t bad = 0;
    for (int64_t it = 0; it < n; ++it) {
        const int qlen = (int)(rng() % 300), tlen = (int)(rng() % 700);
        const bool amb = rng() % 4 == 0;
        // (the packers read a tail in one 32-byte piece when that cannot cross a page: sources may be over-READ
        // within their page by design, so the heap-allocated query gets that much slack; the target sits
        // against a protected page instead)
        std::vector<uint8_t> q((size_t)qlen + 32);
        for (auto &b : q) b = (uint8_t)(rng() & 3);
        if (amb && qlen) q[rng() % (size_t)qlen] = 4;
        // the target ends exactly at the protected page (or a few bytes before it)
        uint8_t *t = area + 2 * page - (size_t)tlen - (rng() % 3 ? 0 : rng() % 40);
        for (int i = 0; i < tlen; ++i) t[i] = (uint8_t)(rng() & 3);
        if (amb && tlen && (rng() & 1)) t[rng() % (size_t)tlen] = (uint8_t)(4 + rng() % 3);
        const uint32_t qb = seq_bytes((uint32_t)qlen, false), tb = seq_bytes((uint32_t)tlen, false);
        const size_t slot = (size_t)slot_words((uint32_t)qlen, (uint32_t)tlen) * 4;
        std::vector<uint8_t> want(slot + 4, 0xAB), got(slot + 4, 0xCD), got2(slot + 4, 0xEF);
        const bool w0a = pack2bit(q.data(), qlen, want.data());
        const bool w0b = pack2bit(t, tlen, want.data() + qb);
        const bool w0 = w0a || w0b;
        const bool w1 = pack_pair_avx2(q.data(), qlen, t, tlen, got.data(), qb);
        const bool w2a = pack2bit_avx2(q.data(), qlen, got2.data());
        const bool w2b = pack2bit_avx2(t, tlen, got2.data() + qb);
        const bool w3a = pack2bit_pext(q.data(), qlen, got2.data());   // overwrites with the same bytes
        const bool w3b = pack2bit_pext(t, tlen, got2.data() + qb);
        bool ok = w0 == w1 && w0 == (w2a || w2b) && w0 == (w3a || w3b);
        if (!w0) {   // packed bytes only matter for pairs without ambiguous bases (the others are re-packed in 4 bits)
            ok = ok && memcmp(want.data(), got.data(), qb + tb) == 0 && memcmp(want.data(), got2.data(), qb + tb) == 0;
        }
        if (!ok) ++bad;
    }
    munmap(area, 3 * page);
    return bad;
}
