/* libbsw_gpu.so -- the B200 drop-in for GenArchBench's bsw hot path.
 *
 * Replaces, behind a plain C ABI, the reference's per-batch kernel entry point
 *     void BandedPairWiseSW::getScores16(SeqPair *pairArray, uint8_t *seqBufRef, uint8_t *seqBufQer,
 *                                        int32_t numPairs, uint16_t numThreads, int32_t w)
 *     (/root/reference/benchmarks/bsw/src/bandedSWA.h:300-305, bandedSWA.cpp:2679-2975),
 * as called from the driver's region of interest (main_banded.cpp:338-350), together with the
 * constructor that fixes the scoring parameters (bandedSWA.h:132-135, bandedSWA.cpp:48-97) and the
 * destructor (bandedSWA.cpp:100-103).
 *
 * Same inputs (the caller's SeqPair array and the two base-code buffers, untouched), same six
 * per-pair outputs written in place (score, tle, gtle, qle, gscore, max_off; bandedSWA.cpp:3336-3362),
 * caller's order preserved. Differences from the reference, all deliberate:
 *   - errors are returned, never exit()ed (reference: bandedSWA.cpp:91-96, 2720-2723);
 *   - entries pairArray[n ..] are never written (reference pads in place, bandedSWA.cpp:2726-2732);
 *   - call it with the WHOLE pair set (or multi-million chunks): binning, batching, packing, the
 *     stream pipeline and the multi-GPU split happen inside; `numThreads` has no equivalent.
 *
 * No torch / CUDA types in any signature. Every entry point returns BSW_OK (0) or a BSW_ERR_* code.
 */
#ifndef BSW_GPU_H
#define BSW_GPU_H
#include "bsw_types.h"
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bsw_handle bsw_handle;

enum {
    BSW_OK = 0,
    BSW_ERR_ARG = 1,        /* null pointer / negative count / bad parameter */
    BSW_ERR_NO_DEVICE = 2,  /* no usable CUDA device: the library has NO CPU fallback */
    BSW_ERR_CUDA = 3,       /* a CUDA runtime call failed; see bsw_gpu_last_error() */
    BSW_ERR_NOMEM = 4,      /* host or device allocation failed */
    BSW_ERR_RANGE = 5,      /* a pair is outside the reference's valid domain (see bsw_gpu_batch) */
    BSW_ERR_STATE = 6       /* staged-API call out of order */
};

/* == BandedPairWiseSW::BandedPairWiseSW (bandedSWA.cpp:48-97). Uses CUDA devices 0..n_gpus-1
 * (n_gpus <= 0: all visible devices). Allocates per-GPU streams, pinned staging and device arenas. */
int bsw_gpu_init(const bsw_params *params, int n_gpus, bsw_handle **out);
/* Same, on an explicit device list (one process per GPU under torchrun passes {LOCAL_RANK}). */
int bsw_gpu_init_devices(const bsw_params *params, int n_devices, const int *device_ids,
                         bsw_handle **out);
/* == BandedPairWiseSW::~BandedPairWiseSW (bandedSWA.cpp:100-103). */
void bsw_gpu_free(bsw_handle *h);

/* Optional: sizes the pinned staging rings and device arenas for calls of about n_pairs pairs holding
 * total_bases bases (len1 + len2 summed), so that the first bsw_gpu_batch does not pay for the
 * allocations (about 0.7 s for full-size slabs: page-locking host memory is slow). The reference
 * allocates its scratch in the constructor as well (bandedSWA.cpp:70-97); here the sizes depend on the
 * input, so it is a separate call. Buffers still grow on demand if a call needs more. */
int bsw_gpu_reserve(bsw_handle *h, int64_t n_pairs, int64_t total_bases);

/* == getScores16 (bandedSWA.cpp:2679-2703) over n pairs with band width w.
 * pairs[k].idr / .idq are byte offsets into ref / qer, .len1 / .len2 the lengths, .h0 the seed score.
 * Writes only score, tle, gtle, qle, gscore, max_off of pairs[0..n).
 * Domain: 0 <= len1, len2 <= BSW_MAX_SEQ_LEN and 0 <= h0.
 *   - h0 + min(len1, len2)*match <= 32767 (bwa-mem2's rule for its 8- and 16-bit classes, bwamem.cpp:2218-2228;
 *     SURVEY.md 8a note 4): the DPX kernels, results bit-identical to getScores16.
 *   - beyond that bound: the pair belongs to what bwa-mem2 calls the scalar class (bwamem.cpp:2218-2228) and is
 *     computed like there, by the rules of scalarBandedSWA (bandedSWA.cpp:132-253) in int32, in one extra launch at
 *     the end of the call (stats.pairs_scalar). The call never fails for such input.
 *   - a record outside the domain above (the reference asserts or reads out of bounds): its six outputs are set to
 *     -1, EVERY OTHER pair of the batch is computed as usual, and the call returns BSW_ERR_RANGE with
 *     stats.pairs_invalid / stats.first_invalid saying how many and where.
 * One call at a time per handle (same rule as one object per thread in the reference, bandedSWA.cpp:2771). */
int bsw_gpu_batch(bsw_handle *h, bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                  int64_t n, int32_t w);

/* == the a-priori classification bwa-mem2 applies before it calls the kernels (bwamem.cpp:2218-2228; the three
 * groups sortPairsLenExt then forms, :1846-1925): with minval = h0 + min(len1, len2) * match,
 *   class 0 (8-bit):  len1, len2, minval < 128;   class 1 (16-bit): all three < 32768;   class 2: scalar.
 * counts[c] = pairs per class; cls (nullable) receives the class of every pair. bsw_gpu_batch applies the same
 * rule itself (classes 0 and 1 share its exact int16 DPX kernels, class 2 runs its int32 kernel), so a caller
 * need not split its batch; the counts are what its own bookkeeping (numPairs128 / 16 / 1) expects. */
int bsw_gpu_classify(const bsw_seqpair *pairs, int64_t n, int32_t match, int64_t counts[3], uint8_t *cls);

/* == the band-doubling retry loop the production caller wraps around getScores16
 * (bwa-mem2, benchmarks/fmi/bwa-mem2/x86_64/src/bwamem.cpp:2448-2508, MAX_BAND_TRY at :51):
 *   for t in 0 .. max_tries-1:  run the pairs still active with band w << t;
 *   a pair is final after try t if its score equals its score of the previous try (-1 before the first),
 *   or max_off < (w<<t)/2 + (w<<t)/4, or t is the last try; the others are compacted and re-run.
 * On return every pair holds the six outputs of its LAST try; tries[k] (nullable) = tries pair k took.
 * Same domain / error rules as bsw_gpu_batch. */
int bsw_gpu_batch_retry(bsw_handle *h, bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                        int64_t n, int32_t w, int32_t max_tries, int32_t *tries);

/* ---- packed input: the pair-file ingest path ----
 * What a packed pair file (include/bsw_pairio.h, "BSWPAIR1") holds, handed over as is: n records and the pairs'
 * sequences at 2 bits per base (4 bits for a pair that holds an ambiguous base, flags bit 0), per pair the query
 * then the target, each padded to 4 bytes, pairs back to back in record order. Replaces, for such input, the
 * reference driver's loader + per-batch call (main_banded.cpp:164-206, 283-287, 338-350): no byte-per-base
 * buffers and no 72-byte records exist on this path. Results come back as 16-byte records in the same order.
 * `data` (4-byte aligned) and `out` are DMA'd in place when they are page-locked (bsw_gpu_host_alloc, or memory the
 * caller registered with CUDA); pageable memory works too and is staged through the library's pinned buffers.
 * Domain: lengths <= BSW_MAX_SEQ_LEN, h0 >= 0 and h0 + min(len1, len2)*match <= 32767 for every record (packed pair
 * files of this repo are written from such data); otherwise BSW_ERR_RANGE. Records are validated slab by slab
 * (~1 M pairs) as the call streams, so on that error slabs before the offending one may already hold results. */
/* (bsw_packed_rec, 12 bytes, and bsw_result, 16 bytes: include/bsw_types.h) */
int bsw_gpu_batch_packed(bsw_handle *h, const bsw_packed_rec *rec, const uint8_t *data, int64_t data_bytes,
                         int64_t n, int32_t w, bsw_result *out);
/* Page-locked host memory (portable across the handle's GPUs) for the packed arrays; free with bsw_gpu_host_free. */
void *bsw_gpu_host_alloc(size_t bytes);
void bsw_gpu_host_free(void *p);

/* ---- staged variant of the same path, for measurement (bench.py `value` vs `e2e`) ----
 * stage:  bin + pack + host->device; inputs stay resident in HBM.
 * run:    launches the DP kernels over the resident batch; *kernel_ms = CUDA-event time on the
 *         launching streams (max over GPUs). May be called repeatedly.
 * fetch:  device->host + scatter of the six outputs into pairs[0..n) (same array order as staged). */
int bsw_gpu_stage(bsw_handle *h, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                  int64_t n, int32_t w);
int bsw_gpu_run_staged(bsw_handle *h, float *kernel_ms);
int bsw_gpu_fetch_staged(bsw_handle *h, bsw_seqpair *pairs, int64_t n);
/* Unit of work of the GCUPS metric for the staged batch: the number of DP cells the reference's
 * scalar loop visits (bandedSWA.cpp:191-216, the commented SW_cells++ at :215), counted on the device
 * by a COUNT variant of the kernel that tracks the reference's exact beg/end. Not for timed regions. */
int bsw_gpu_count_staged(bsw_handle *h, int64_t *cells_visited);

typedef struct bsw_gpu_stats {
    int64_t pairs;              /* pairs processed by the last batch / staged run */
    int64_t kernel_launches;    /* launches of OUR kernels in the last batch / run_staged call */
    int64_t h2d_bytes;          /* bytes copied host->device by the last batch / stage */
    int64_t d2h_bytes;          /* bytes copied device->host by the last batch / fetch */
    int64_t pairs_short;        /* pairs routed to the thread-per-pair shared-memory kernel */
    int64_t pairs_long;         /* pairs routed to the long-pair kernel */
    double  host_bin_ms;        /* last batch: validation + key extraction */
    double  host_pack_ms;       /* last batch: 2-bit packing into pinned staging */
    double  host_scatter_ms;    /* last batch: result scatter into SeqPair */
    double  kernel_ms;          /* last batch: sum of CUDA-event kernel time, max over GPUs */
    double  wall_ms;            /* last batch: wall time of the whole call */
    int32_t n_gpus;
    int32_t reserved;
    double  host_sort_ms;       /* last batch: the two counting sorts + gather of sorted lengths */
    double  host_plan_ms;       /* last batch: launch planning + slot offsets */
    double  host_alloc_ms;      /* last batch: (re)allocation of pinned / device buffers */
    double  host_cut_ms;        /* last batch: slab cutting */
    double  host_wait_ms;       /* last batch: host blocked on the GPU (ring slot not yet drained) */
    int64_t pairs_keyed;        /* of pairs_short: launched with the keyed row argmax (scores and group
                                   indices of the launch share 16 bits) */
    int64_t pairs_duo;          /* of pairs_short: launched on the two-pairs-per-thread kernel */
    int64_t pairs_scalar;       /* last batch: pairs of the scalar class (score bound beyond int16; int32 kernel) */
    int64_t pairs_invalid;      /* last batch: invalid records (outputs set to -1; the call returned BSW_ERR_RANGE) */
    int64_t first_invalid;      /* index of the first of them, -1 if none */
} bsw_gpu_stats;
int bsw_gpu_get_stats(const bsw_handle *h, bsw_gpu_stats *out);

/* Integer-pipe microbenchmark on device `device`: packed s16x2 DPX instruction throughput in
 * giga thread-instructions per second (warp-instructions * 32), all SMs, `which`:
 *   0 VIADDMNMX.S16x2.RELU, 1 VIMNMX3.S16x2, 2 VIADD.16x2, 3 LOP3, 4 PRMT, 5 IMAD,
 *   6 SHF, 7 IMAD.HI, 8 VIADDMNMX + IMAD co-issue; 9 = the arithmetic of one inner-loop trip of the
 *   thread-per-pair kernel on registers only, reported in giga CELLS per second (the ceiling of that
 *   kernel if shared memory, row bookkeeping and divergence were free).
 *   Used by bench.py for the `dpx_peak` roofline denominator. */
int bsw_gpu_dpx_peak(int device, int which, double *ginstr_per_s, double *sm_mhz_est);

/* Developer probe: the register-only inner-loop trips (kind 0: one pair per thread, 1: two pairs per thread) at
 * warps_per_sm one-warp blocks per SM, in giga cells per second. */
int bsw_gpu_trip_probe(int device, int kind, int warps_per_sm, double *gcells_per_s);

const char *bsw_gpu_strerror(int code);
/* Text of the last CUDA error seen by this handle (empty string if none). */
const char *bsw_gpu_last_error(const bsw_handle *h);
/* Library / kernel ABI version, bumped when bsw_seqpair handling changes. */
int bsw_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
