/* kswv_gpu.h -- C ABI of the B200 replacement for bwa-mem2's batched mate-rescue Smith-Waterman (class kswv).
 *
 * SURVEY.md 8(f)3. The reference interface is the class in
 *   /root/reference/benchmarks/fmi/bwa-mem2/x86_64/src/kswv.h:60-190 (ctor kswv.cpp:117-160; getScores8 :165,
 *   getScores16 :719) as its production caller drives it: sort_classify (bwamem.cpp:1136-1163) and the vector
 *   branch of mem_sam_pe_batch (bwamem_pair.cpp:612-707).
 * One kswv_gpu_batch call replaces that whole branch: both score classes, phase 0 (score, te, qe, score2, te2)
 * and phase 1 (tb, qb from the reversed prefixes). The entry points live in libbsw_gpu.so next to the bsw ones;
 * error codes and kswv_gpu_strerror are those of bsw_gpu.h. No CPU fallback: without a CUDA device
 * kswv_gpu_init returns BSW_ERR_NO_DEVICE.
 *
 * Results are those of the reference's vector kernels, bit for bit, including what distinguishes them from
 * ksw_align2: the query padded with zero-score columns to a multiple of 16 (8-bit class) / 8 (16-bit class),
 * the rising-row filter behind score2/te2, score = 255 for a saturated 8-bit pair (no 16-bit rerun; score2 =
 * te2 = -1), len1 unchanged in phase 1. tests/test_kswv_gpu.py compares with oracle/kswv_oracle.c, which is
 * pinned to the compiled reference.
 */
#ifndef KSWV_GPU_H
#define KSWV_GPU_H

#include <stdint.h>
#include "bsw_gpu.h"   /* BSW_OK / BSW_ERR_* and bsw_gpu_strerror, bsw_gpu_host_alloc */
#include "bsw_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kswv_handle kswv_handle;

/* kswv::kswv(o_del, e_del, o_ins, e_ins, w_match, w_mismatch, ...) (kswv.cpp:117-124); mismatch is the positive
 * penalty (the caller passes -1 * opt->b, bwamem_pair.cpp:640). The ambiguous-base score is DEFAULT_AMBIG = -1
 * in the reference (kswv.cpp:131) and here. */
typedef struct kswv_params { int32_t o_del, e_del, o_ins, e_ins, match, mismatch; } kswv_params;
#define KSWV_DEFAULT_PARAMS { 6, 1, 6, 1, 1, 4 }

/* == kswr_t (ksw.h:45-50) */
typedef struct kswv_result { int32_t score, te, qe, score2, te2, tb, qb; } kswv_result;

/* xtra flags carried in bsw_seqpair.h0 (ksw.h:31-34; built at bwamem_pair.cpp:1003) */
#define KSWV_XBYTE  0x10000
#define KSWV_XSTOP  0x20000
#define KSWV_XSUBO  0x40000
#define KSWV_XSTART 0x80000

typedef struct kswv_gpu_stats {
    int32_t n_gpus;
    int32_t chunks;             /* last batch: pipeline chunks */
    int32_t lanes_per_pair;     /* last chunk: lanes per pair of its plain pairs (8, 16 or 32) */
    int32_t reserved;
    int64_t pairs;              /* last batch */
    int64_t pairs8;             /* last batch: pairs of the 8-bit class (KSWV_XBYTE) */
    int64_t cells;              /* last batch: phase-0 DP cells, len1 x padded query columns (the CUPS numerator) */
    int64_t h2d_bytes, d2h_bytes;
    int64_t kernel_launches;
    int64_t gathered;           /* last batch: chunks whose sequences were gathered on the host (not one dense range) */
    double kernel_ms;           /* last batch: sum over chunks of the kernel's event time (max over GPUs per chunk wave) */
    double wall_ms;             /* last batch: the call, host clock */
    double host_check_ms;       /* last batch: validation pass over the records */
    double host_prep_ms;        /* last batch: task ordering + task records (+ gather), summed over chunks */
    double host_wait_ms;        /* last batch: blocked on a slot's results (includes the scatter to aln) */
    int64_t staged;             /* last batch: dense chunks copied through page-locked staging (pageable caller buffers) */
} kswv_gpu_stats;

/* Scoring is fixed per handle, as in the reference's constructor. n_gpus <= 0: all visible devices.
 * BSW_ERR_ARG for parameters outside 0 < match <= 127, 0 < mismatch <= 127, 0 <= o, 0 < e, o + e <= 127. */
int kswv_gpu_init(const kswv_params *params, int n_gpus, kswv_handle **out);
void kswv_gpu_free(kswv_handle *h);

/* == sort_classify + mem_sam_pe_batch's vector branch (bwamem.cpp:1136-1163, bwamem_pair.cpp:634-704).
 * pairs[i]: idr / idq = byte offsets into ref / qer, len1 / len2, h0 = xtra (KSWV_X* | threshold), regid = the
 * slot of aln the result goes to (0 <= regid < n_pairs, as mem_matesw_batch_pre numbers them, bwamem_pair.cpp:1085).
 * Bases are 0..3, anything above is the ambiguous base. Pairs may come in any order and both classes mixed: the
 * class is KSWV_XBYTE in h0, which is all sort_classify looks at. The pair array and the sequence buffers are
 * not modified (the reference reverses the aligned prefixes in place and leaves them reversed).
 * Domain (BSW_ERR_RANGE, nothing computed): 0 <= len1, len2 <= 32767 (te and qe are int16 in the reference), and
 * for the 16-bit class min(len1, len2) * match <= 32767 (the reference's int16 lanes wrap above that).
 * Sequences are sent as one range per chunk when a chunk's pairs lie densely and in order in ref / qer (the
 * production layout); page-locked buffers (bsw_gpu_host_alloc / cudaHostRegister) then go to the device without any
 * host copy, pageable ones through the library's page-locked staging. */
int kswv_gpu_batch(kswv_handle *h, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                   int64_t n_pairs, kswv_result *aln);

int kswv_gpu_get_stats(const kswv_handle *h, kswv_gpu_stats *out);
const char *kswv_gpu_last_error(const kswv_handle *h);

#ifdef __cplusplus
}
#endif
#endif
