// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// C shim around the UNMODIFIED reference kernel, compiled in place from
// /root/reference/benchmarks/bsw/src/bandedSWA.cpp by oracle/Makefile into oracle/_ref/.
// No reference source is copied into this repository: this file only *calls* the reference's
// public class (bandedSWA.h:127-412) the way its own driver does (main_banded.cpp:266-350).
//
// Exposed entry points (plain C ABI, loaded with ctypes by tests/ and bench.py's CPU arm):
//   ref_bsw_new / ref_bsw_free     -- one BandedPairWiseSW per worker thread (main_banded.cpp:271-276)
//   ref_bsw_getscores16            -- the ROI loop of main_banded.cpp:338-350 (omp dynamic,1 over
//                                     batches of `batch` pairs, each through getScores16)
//   ref_bsw_scalar                 -- scalarBandedSWAWrapper (bandedSWA.cpp:258-276)
//   ref_bsw_simd_width16           -- SIMD_WIDTH16 of this build (bandedSWA.h:66-92)
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <omp.h>
#include "bandedSWA.h"

// bandedSWA.cpp:41 expects the driver to define this (main_banded.cpp:92).
uint64_t prof[10][112];

namespace {

struct RefHandle {
    int8_t mat[25];
    int nthreads;
    std::vector<BandedPairWiseSW *> workers;
};

// same fill rule as the reference driver's scoring matrix (main_banded.cpp:94-102):
// a on the diagonal, -b elsewhere, ambig for any row/column 4.
void fill_mat(int a, int b, int ambig, int8_t mat[25]) {
    int k = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) mat[k++] = (i == j) ? a : -b;
        mat[k++] = ambig;
    }
    for (int j = 0; j < 5; ++j) mat[k++] = ambig;
}

}  // namespace

extern "C" {

// params: {o_del, e_del, o_ins, e_ins, zdrop, end_bonus, match, mismatch(+ve), ambig}
void *ref_bsw_new(const int32_t *params, int nthreads) {
    RefHandle *h = new RefHandle();
    fill_mat(params[6], params[7], params[8], h->mat);
    h->nthreads = nthreads < 1 ? 1 : nthreads;
    for (int t = 0; t < h->nthreads; ++t)
        h->workers.push_back(new BandedPairWiseSW(params[0], params[1], params[2], params[3],
                                                  params[4], params[5], h->mat,
                                                  (int8_t)params[6], (int8_t)params[7], 1));
    return h;
}

void ref_bsw_free(void *hv) {
    RefHandle *h = (RefHandle *)hv;
    if (!h) return;
    for (auto *w : h->workers) delete w;
    delete h;
}

int ref_bsw_simd_width16(void) { return SIMD_WIDTH16; }
int ref_bsw_sizeof_seqpair(void) { return (int)sizeof(SeqPair); }

// Runs getScores16 over `n` pairs exactly as the driver's ROI does. `pairs` uses GLOBAL byte offsets
// idr/idq into ref/qer. Works on a private, over-allocated copy because the reference pads
// pairArray[n..round) in place (bandedSWA.cpp:2726-2732), needs batch-local `id`
// (bandedSWA.cpp:1972-1986) and prefetches past the end (bandedSWA.cpp:2814).
// Returns 0; *roi_seconds = wall time of the parallel loop only.
int ref_bsw_getscores16(void *hv, SeqPair *pairs, uint8_t *ref, uint8_t *qer, int64_t n, int32_t w,
                        int32_t batch, double *roi_seconds) {
    RefHandle *h = (RefHandle *)hv;
    const int W = SIMD_WIDTH16;
    if (batch <= 0) batch = 512;
    batch = ((batch + W - 1) / W) * W;
    int64_t round = ((n + W - 1) / W) * W;
    SeqPair *work = (SeqPair *)_mm_malloc((round + 64) * sizeof(SeqPair), 64);
    if (!work) return 1;
    memset(work, 0, (round + 64) * sizeof(SeqPair));
    memcpy(work, pairs, n * sizeof(SeqPair));
    for (int64_t i = 0; i < n; ++i) work[i].id = i % batch;

    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(h->nthreads)
    {
        int tid = omp_get_thread_num();
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < round; i += batch) {
            int nb = (int)((n - i) >= batch ? batch : n - i);
            if (nb > 0) h->workers[tid]->getScores16(work + i, ref, qer, nb, 1, w);
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (roi_seconds) *roi_seconds = std::chrono::duration<double>(t1 - t0).count();

    for (int64_t i = 0; i < n; ++i) {
        pairs[i].score = work[i].score; pairs[i].tle = work[i].tle; pairs[i].gtle = work[i].gtle;
        pairs[i].qle = work[i].qle; pairs[i].gscore = work[i].gscore; pairs[i].max_off = work[i].max_off;
    }
    _mm_free(work);
    return 0;
}

int ref_bsw_scalar(void *hv, SeqPair *pairs, uint8_t *ref, uint8_t *qer, int64_t n, int32_t w) {
    RefHandle *h = (RefHandle *)hv;
#pragma omp parallel for num_threads(h->nthreads) schedule(dynamic, 512)
    for (int64_t i = 0; i < n; ++i)
        h->workers[omp_get_thread_num()]->scalarBandedSWAWrapper(pairs + i, ref, qer, 1, 1, w);
    return 0;
}

}  // extern "C"
