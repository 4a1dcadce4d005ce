/* Shared plain-C types of the bsw drop-in boundary.
 *
 * bsw_seqpair is layout-identical to the reference's `SeqPair`
 * (/root/reference/benchmarks/bsw/src/bandedSWA.h:104-113): 72 bytes,
 *   int64 idr@0, idq@8, id@16; int32 len1@24, len2@28, h0@32, seqid@36, regid@40,
 *   score@44, tle@48, gtle@52, qle@56, gscore@60, max_off@64 (+4 tail padding).
 * idr / idq are byte offsets of the target ("ref", len1 bytes) and the query (len2 bytes) inside the
 * caller's seqBufRef / seqBufQer (main_banded.cpp:192-193). Bases are codes 0..3, 4 = ambiguous.
 */
#ifndef BSW_TYPES_H
#define BSW_TYPES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bsw_seqpair {
    int64_t idr, idq, id;
    int32_t len1, len2;
    int32_t h0;
    int32_t seqid, regid;
    int32_t score, tle, gtle, qle;
    int32_t gscore, max_off;
} bsw_seqpair;

/* Scoring parameters fixed at construction in the reference
 * (BandedPairWiseSW::BandedPairWiseSW, bandedSWA.cpp:48-68; driver defaults main_banded.cpp:70-74,268).
 * mismatch is the positive penalty (the reference negates it, bandedSWA.cpp:64). */
typedef struct bsw_params {
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t zdrop, end_bonus;
    int32_t match, mismatch, ambig;
} bsw_params;

/* Packed pair input (include/bsw_pairio.h "BSWPAIR1", bsw_gpu_batch_packed): one 12-byte record per pair;
 * flags bit 0 = the pair's sequences are stored at 4 bits per base (it holds an ambiguous base), else 2 bits. */
typedef struct bsw_packed_rec { uint16_t len1, len2; int32_t h0; uint32_t flags; } bsw_packed_rec;
/* The six per-pair outputs of the reference (bandedSWA.cpp:3336-3362) as one 16-byte record. */
typedef struct bsw_result { int16_t score, qle, tle, gtle, gscore, max_off; uint32_t reserved; } bsw_result;

#define BSW_DEFAULT_PARAMS { 6, 1, 6, 1, 100, 5, 1, 4, -1 }
#define BSW_DEFAULT_BAND 100
#define BSW_MAX_SEQ_LEN 32767 /* MAX_SEQ_LEN16 - 1, bandedSWA.h:97 */

#ifdef __cplusplus
}
#endif
#endif
