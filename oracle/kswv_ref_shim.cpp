// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// C shim around the UNMODIFIED bwa-mem2 mate-rescue Smith-Waterman (class kswv), compiled in place from
// /root/reference/benchmarks/fmi/bwa-mem2/x86_64/src/kswv.cpp by oracle/Makefile into
// oracle/_ref/libkswv_ref_avx512.so. No reference source is copied into this repository: the reference
// translation unit is pulled in by path below and only its public class (kswv.h:60-190) is *called*.
// The class only has an AVX512BW body (kswv.cpp:60,163,...), so this library exists only for hosts with
// AVX512BW; tests that need it skip elsewhere.
//
//   ref_kswv_batch -- what the production caller does with the class: sort_classify (bwamem.cpp:1136-1163,
//                     8-bit class = KSW_XBYTE set in h0) and the vector branch of mem_sam_pe_batch
//                     (bwamem_pair.cpp:634-704): phase 0 over both classes, the in-place reversal of the
//                     aligned prefixes, phase 1 with h0 = KSW_XSTOP | score.
//
// kswv.cpp reaches utils.h, which re-declares __rdtsc() as a static function; the macro below moves that
// declaration out of the way of the compiler's own intrinsic. Nothing else is changed.
#include <immintrin.h>
#define __rdtsc bwa_utils_rdtsc_shadow
#include "kswv.cpp"
#undef __rdtsc

#include <cstring>
#include <vector>

namespace {
inline void reverse_prefix(int l, uint8_t *s) {
    for (int i = 0; i < l >> 1; ++i) { uint8_t t = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = t; }
}
}  // namespace

extern "C" {

int ref_kswv_sizeof_seqpair(void) { return (int)sizeof(SeqPair); }
int ref_kswv_sizeof_kswr(void) { return (int)sizeof(kswr_t); }

// params: {o_del, e_del, o_ins, e_ins, match, mismatch(+ve)}; pairs[i].h0 = xtra, regid = index into aln.
// aln[0..n) is pre-set to {0,-1,-1,-1,-1,-1,-1}: the reference leaves score2/te2 of a 64-lane batch whose
// lanes all saturated unwritten (kswv.cpp:586), and tb/qb are reset to -1 by the caller (bwamem_pair.cpp:635).
// The sequence buffers are copied first: the reference reverses them in place.
int ref_kswv_batch(const int32_t *params, const SeqPair *pairs, const uint8_t *ref, int64_t ref_bytes,
                   const uint8_t *qer, int64_t qer_bytes, int32_t n, kswr_t *aln_out) {
    int maxRef = 0, maxQer = 0;
    for (int i = 0; i < n; ++i) {
        if (pairs[i].len1 > maxRef) maxRef = pairs[i].len1;
        if (pairs[i].len2 > maxQer) maxQer = pairs[i].len2;
    }
    std::vector<uint8_t> rbuf(ref, ref + ref_bytes), qbuf(qer, qer + qer_bytes);
    rbuf.resize(rbuf.size() + 64); qbuf.resize(qbuf.size() + 64);
    std::vector<SeqPair> arr((size_t)n + MAX_LINE_LEN + 2 * SIMD_WIDTH8);
    memset(arr.data(), 0, arr.size() * sizeof(SeqPair));
    std::vector<kswr_t> aln((size_t)n + SIMD_WIDTH8);
    for (auto &r : aln) r = g_defr;

    // sort_classify (bwamem.cpp:1136-1163): 8-bit class first, order kept inside each class
    int64_t pcnt = n, pcnt8 = 0;
    for (int i = 0; i < n; ++i) if (pairs[i].h0 & KSW_XBYTE) arr[pcnt8++] = pairs[i];
    {   int64_t k = pcnt8;
        for (int i = 0; i < n; ++i) if (!(pairs[i].h0 & KSW_XBYTE)) arr[k++] = pairs[i];
    }
    SeqPair *seqPairArray = arr.data();
    uint8_t *seqBufRef = rbuf.data(), *seqBufQer = qbuf.data();

    kswv *pwsw = new kswv(params[0], params[1], params[2], params[3], (int8_t)params[4], (int8_t)(-params[5]),
                          1, maxRef, maxQer);
    // bwamem_pair.cpp:646-653
    for (int64_t i = 0; i < pcnt - pcnt8; i++)
        seqPairArray[pcnt + MAX_LINE_LEN - 1 - i] = seqPairArray[pcnt - i - 1];
    pwsw->getScores8(seqPairArray, seqBufRef, seqBufQer, aln.data(), (int32_t)pcnt8, 1, 0);
    pwsw->getScores16(seqPairArray + pcnt8 + MAX_LINE_LEN, seqBufRef, seqBufQer, aln.data(),
                      (int32_t)(pcnt - pcnt8), 1, 0);
    // bwamem_pair.cpp:660-695
    int64_t pos = 0, pos8 = 0, pos16 = 0;
    for (int64_t i = 0; i < pcnt8; i++) {
        SeqPair sp = seqPairArray[i];
        kswr_t r = aln[sp.regid];
        int xtra = sp.h0;
        if ((xtra & KSW_XSTART) == 0 || ((xtra & KSW_XSUBO) && r.score < (xtra & 0xffff))) continue;
        sp.h0 = KSW_XSTOP | r.score;
        sp.len2 = r.qe + 1;
        reverse_prefix(r.qe + 1, seqBufQer + sp.idq);
        reverse_prefix(r.te + 1, seqBufRef + sp.idr);
        seqPairArray[pos++] = sp;
        pos8++;
    }
    const int64_t id = pcnt8 + MAX_LINE_LEN;
    for (int64_t i = 0; i < pcnt - pcnt8; i++) {
        SeqPair sp = seqPairArray[i + id];
        kswr_t r = aln[sp.regid];
        int xtra = sp.h0;
        if ((xtra & KSW_XSTART) == 0 || ((xtra & KSW_XSUBO) && r.score < (xtra & 0xffff))) continue;
        sp.h0 = KSW_XSTOP | r.score;
        sp.len2 = r.qe + 1;
        reverse_prefix(r.qe + 1, seqBufQer + sp.idq);
        reverse_prefix(r.te + 1, seqBufRef + sp.idr);
        seqPairArray[pos++] = sp;
        pos16++;
    }
    // bwamem_pair.cpp:697-699
    pwsw->getScores16(seqPairArray + pos8, seqBufRef, seqBufQer, aln.data(), (int32_t)pos16, 1, 1);
    pwsw->getScores8(seqPairArray, seqBufRef, seqBufQer, aln.data(), (int32_t)pos8, 1, 1);
    delete pwsw;
    memcpy(aln_out, aln.data(), sizeof(kswr_t) * (size_t)n);
    return 0;
}

}  // extern "C"
