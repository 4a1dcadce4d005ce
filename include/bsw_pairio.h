/* Host-only helpers around the reference driver's pair-file format: a synthetic pair generator
 * (SURVEY.md 8d), and a reader / writer for the 3-line text format of
 * /root/reference/benchmarks/bsw/src/main_banded.cpp:152-206 (h0 line, reference line, query line;
 * bases as the digits '0'..'4').  No CUDA. Built into libbsw_pairio.so.
 */
#ifndef BSW_PAIRIO_H
#define BSW_PAIRIO_H
#include "bsw_types.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { BSW_GEN_READ_FLANK = 0, BSW_GEN_UNIFORM = 1, BSW_GEN_LOGUNIFORM = 2 };

typedef struct bsw_gen_config {
    int32_t mode;              /* BSW_GEN_* */
    int32_t read_len;          /* READ_FLANK: read length (151) */
    int32_t seed_min, seed_max;/* READ_FLANK: seed length s ~ U[min,max], h0 = s */
    int32_t len2_min, len2_max;/* UNIFORM / LOGUNIFORM: query length range */
    int32_t h0_min, h0_max;    /* UNIFORM / LOGUNIFORM: h0 ~ U[min,max] */
    int32_t tail_cap;          /* READ_FLANK/UNIFORM: len1 = len2 + min(max(len2-5,1), tail_cap) */
    int32_t extra_max;         /* LOGUNIFORM: len1 = len2 + U[0,extra_max] */
    double  sub_rate, indel_rate;
    double  n_frac;            /* fraction of pairs that get one ambiguous base (code 4) */
    double  small_h0_frac;     /* fraction of pairs with h0 in {0,1} */
    double  random_frac;       /* fraction of pairs whose target is unrelated to the query */
    uint64_t seed;
} bsw_gen_config;

/* The five BASELINE.json configurations: 1 small, 2 16-bit, 3 large, 4 skewed, 5 scaling. */
int bsw_gen_preset(int config_id, bsw_gen_config *out);

/* Generates n pairs into `pairs` (caller-allocated, n entries) and two malloc'd, densely packed
 * sequence buffers returned through ref_out / qer_out (free with bsw_host_free). idr/idq are byte
 * offsets into them, id = index, outputs = -1 (main_banded.cpp:200-201).
 * Deterministic for a given (config, n) regardless of thread count. Returns 0 on success. */
int bsw_gen_pairs(const bsw_gen_config *cfg, int64_t n, bsw_seqpair *pairs, uint8_t **ref_out,
                  uint8_t **qer_out, int64_t *ref_bytes, int64_t *qer_bytes, int32_t nthreads);

void bsw_host_free(void *p);

/* Text pair file, the reference's format. */
int bsw_write_pairs_text(const char *path, const bsw_seqpair *pairs, const uint8_t *ref,
                         const uint8_t *qer, int64_t n);
/* Counts pairs (lines / 3, main_banded.cpp:237-253); returns -1 on error. */
int64_t bsw_count_pairs_text(const char *path);
/* Reads up to n pairs; allocates dense ref/qer buffers like bsw_gen_pairs. Unlike the reference
 * loader (fixed 2048/256-byte strides, main_banded.cpp:76-79,172-176) line length is unbounded
 * below BSW_MAX_SEQ_LEN, and the file is parsed by all host threads. Returns the number of pairs read, or -1. */
int64_t bsw_read_pairs_text(const char *path, int64_t n, bsw_seqpair *pairs, uint8_t **ref_out,
                            uint8_t **qer_out, int64_t *ref_bytes, int64_t *qer_bytes);

/* Packed binary pair file: "BSWPAIR1", u64 n, u64 data bytes, n records
 * { u16 len1, u16 len2, i32 h0, u32 flags }, then per pair query + target at 2 bits per base (4 if the pair
 * holds an ambiguous base), each padded to 4 bytes. ~3.5x smaller than the text format, read in parallel.
 * Same return conventions as the text functions. */
int bsw_write_pairs_packed(const char *path, const bsw_seqpair *pairs, const uint8_t *ref,
                           const uint8_t *qer, int64_t n);
int64_t bsw_count_pairs_packed(const char *path);
int64_t bsw_read_pairs_packed(const char *path, int64_t n, bsw_seqpair *pairs, uint8_t **ref_out,
                              uint8_t **qer_out, int64_t *ref_bytes, int64_t *qer_bytes);

/* The same packed form in memory, for bsw_gpu_batch_packed (include/bsw_gpu.h):
 *   bsw_packed_bytes    : bytes the packed sequences of pairs[0..n) take (-1: a length out of range);
 *   bsw_pack_pairs      : fills rec[0..n) and data[0..data_cap) (caller-allocated, e.g. page-locked); 0 on success;
 *   bsw_packed_file_info: n and the data bytes of a packed pair file (0 on success);
 *   bsw_read_packed_raw : reads the file's records and data as they are into caller-allocated arrays
 *                         (no unpacking to one byte per base); returns the pairs read or -1. */
int64_t bsw_packed_bytes(const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer, int64_t n);
int bsw_pack_pairs(const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer, int64_t n,
                   bsw_packed_rec *rec, uint8_t *data, int64_t data_cap);
int bsw_packed_file_info(const char *path, int64_t *n, int64_t *data_bytes);
int64_t bsw_read_packed_raw(const char *path, int64_t n, bsw_packed_rec *rec, uint8_t *data, int64_t data_cap);

#ifdef __cplusplus
}
#endif
#endif
