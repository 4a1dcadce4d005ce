// TEST INFRASTRUCTURE: INTEGRATION.md section 4b, compiled. This program includes the REFERENCE's own headers
// (kswv.h -> ksw.h, bandedSWA.h, from where they lie under /root/reference; nothing is copied) for SeqPair, kswr_t and
// the KSW_X* flags, fills a batch the way mem_matesw_batch_pre does (bwamem_pair.cpp:1003-1086), and hands the
// reference-typed arrays to kswv_gpu_batch with the casts the binding uses. The same batch goes through the
// unmodified class (oracle/_ref/libkswv_ref_avx512.so: sort_classify + mem_sam_pe_batch's vector branch); the seven
// kswr_t fields must agree. Built by integration/Makefile (needs AVX512BW to run the reference side).
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <dlfcn.h>

// bandedSWA.h reaches utils.h, which re-declares __rdtsc() as a static function: keep it away from the compiler's own
#include <immintrin.h>
#define __rdtsc bwa_utils_rdtsc_shadow
#include "kswv.h"          // the reference's: SeqPair (bandedSWA.h:91-100), kswr_t (ksw.h:45-50), KSW_X* (ksw.h:31-34)
#undef __rdtsc
#include "kswv_gpu.h"      // ours

static_assert(sizeof(SeqPair) == sizeof(bsw_seqpair), "SeqPair layout");
static_assert(offsetof(SeqPair, idr) == offsetof(bsw_seqpair, idr) && offsetof(SeqPair, idq) == offsetof(bsw_seqpair, idq) &&
              offsetof(SeqPair, len1) == offsetof(bsw_seqpair, len1) && offsetof(SeqPair, len2) == offsetof(bsw_seqpair, len2) &&
              offsetof(SeqPair, h0) == offsetof(bsw_seqpair, h0) && offsetof(SeqPair, regid) == offsetof(bsw_seqpair, regid),
              "SeqPair fields");
static_assert(sizeof(kswr_t) == sizeof(kswv_result) && offsetof(kswr_t, score2) == offsetof(kswv_result, score2) &&
              offsetof(kswr_t, tb) == offsetof(kswv_result, tb) && offsetof(kswr_t, qb) == offsetof(kswv_result, qb), "kswr_t layout");
static_assert(KSW_XBYTE == KSWV_XBYTE && KSW_XSTOP == KSWV_XSTOP && KSW_XSUBO == KSWV_XSUBO && KSW_XSTART == KSWV_XSTART, "xtra flags");

typedef int (*ref_batch_fn)(const int32_t *, const SeqPair *, const uint8_t *, int64_t, const uint8_t *, int64_t, int32_t, kswr_t *);

static uint32_t rng_state = 12345;
static inline uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 20000;
    const char *reflib = argc > 2 ? argv[2] : nullptr;
    const int a = 1, b = 4, o = 6, e = 1, min_seed_len = 19;
    std::vector<SeqPair> pairs((size_t)n + 256);                  // the reference pads its array (kswv.cpp:193-199)
    std::vector<uint8_t> ref, qer;
    for (int i = 0; i < n; ++i) {
        const int l_ms = 60 + (int)(rnd() % 240);                 // both classes: l_ms * a < 250 is the 8-bit one
        const int l_ref = l_ms * 2 + (int)(rnd() % (3 * l_ms));
        SeqPair sp;
        memset(&sp, 0, sizeof sp);
        sp.idr = (int64_t)ref.size(); sp.idq = (int64_t)qer.size();
        sp.len1 = l_ref; sp.len2 = l_ms;
        sp.h0 = KSW_XSUBO | KSW_XSTART | (l_ms * a < 250 ? KSW_XBYTE : 0) | (min_seed_len * a);   // bwamem_pair.cpp:1003
        sp.regid = i;
        for (int k = 0; k < l_ref; ++k) ref.push_back((uint8_t)(rnd() & 3));
        const int at = (int)(rnd() % (uint32_t)(l_ref - l_ms + 1));
        for (int k = 0; k < l_ms; ++k) {
            uint8_t c = ref[(size_t)sp.idr + at + k];
            if (rnd() % 25 == 0) c = (uint8_t)(rnd() & 3);        // substitutions
            if (rnd() % 400 == 0) c = 4;                          // an ambiguous base now and then
            qer.push_back(c);
        }
        pairs[(size_t)i] = sp;
    }
    ref.resize(ref.size() + 64); qer.resize(qer.size() + 64);

    // ---- the binding of INTEGRATION.md 4b
    kswv_handle *h = nullptr;
    kswv_params kp = { o, e, o, e, a, b };
    int rc = kswv_gpu_init(&kp, 1, &h);
    if (rc != BSW_OK) { fprintf(stderr, "kswv_gpu_init: %d\n", rc); return 2; }
    std::vector<kswr_t> aln((size_t)n + 64);
    rc = kswv_gpu_batch(h, (const bsw_seqpair *)pairs.data(), ref.data(), qer.data(), n, (kswv_result *)aln.data());
    if (rc != BSW_OK) { fprintf(stderr, "kswv_gpu_batch: %s\n", kswv_gpu_last_error(h)); return 2; }
    kswv_gpu_free(h);

    // ---- the unmodified class on the same batch
    long mism = -1;
    if (reflib) {
        void *L = dlopen(reflib, RTLD_NOW);
        ref_batch_fn fn = L ? (ref_batch_fn)dlsym(L, "ref_kswv_batch") : nullptr;
        if (!fn) { fprintf(stderr, "cannot load %s\n", reflib); return 2; }
        const int32_t params[6] = { o, e, o, e, a, b };
        std::vector<kswr_t> want((size_t)n + 64);
        fn(params, pairs.data(), ref.data(), (int64_t)ref.size() - 64, qer.data(), (int64_t)qer.size() - 64, n, want.data());
        mism = 0;
        for (int i = 0; i < n; ++i) mism += memcmp(&want[(size_t)i], &aln[(size_t)i], sizeof(kswr_t)) != 0;
    }
    long s = 0, with_start = 0;
    for (int i = 0; i < n; ++i) { s += aln[(size_t)i].score; with_start += aln[(size_t)i].tb >= 0; }
    printf("kswv binding: %d pairs, score sum %ld, %ld with start positions, mismatches vs the reference class: %ld\n",
           n, s, with_start, mism);
    return mism > 0 ? 1 : 0;
}
