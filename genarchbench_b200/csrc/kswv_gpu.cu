// kswv_gpu.cu -- host side of the kswv path (include/kswv_gpu.h): chunks of pairs flow through a ring of three
// slots per GPU (tasks + sequences H2D, the two kernels, results D2H, each on its own stream; every GPU a contiguous
// share of the call; chunks dealt round by round by a coordinator, or by a host worker per GPU above four GPUs), the host orders each chunk's
// tasks by decreasing DP size and scatters finished results to aln[regid]. Kernels: kswv_kernels.cuh.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "bsw_gpu.h"
#include "kswv_gpu.h"
#include "kswv_kernels.cuh"

using namespace kswvk;

namespace {

constexpr int kRing = 3;
int64_t kChunkPairs = 32768;                    // pairs per chunk (about 20 MB of sequence at 150 bp reads); KSWV_CHUNK_PAIRS
constexpr int64_t kChunkBytes = 96ll << 20;     // and at most this many sequence bytes
constexpr int64_t kMinChunk = 16384;            // no chunk smaller than this unless the share is
constexpr int64_t kMaxChunk = 131072;           // chunks grow with the share up to this
constexpr int kBlocksPerSm = 4;                 // x kKswvWarps warps

struct KSlot {
    Task *h_tasks = nullptr, *d_tasks = nullptr;
    Result *h_out = nullptr, *d_out = nullptr;
    uint8_t *h_seq = nullptr;                   // gather staging (only when a chunk is not one dense range)
    uint8_t *d_ref = nullptr, *d_qer = nullptr;
    int *d_counter = nullptr;                   // four task counters: phase 0 / phase 1 x plain / special
    uint32_t *d_p1 = nullptr;                   // 4 x cap_pairs: phase-1 keys, task indices, and both sorted
    void *d_sort = nullptr;
    size_t cap_sort = 0;
    size_t cap_pairs = 0, cap_ref = 0, cap_qer = 0, cap_seq = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_done = nullptr, ev_in = nullptr, ev_k = nullptr;
    bool busy = false;
    int64_t first = 0, count = 0;               // the chunk in flight: pairs [first, first + count)
};

struct KDev {
    int id = 0, sms = 0, warps = 0, groups = 0;
    int occ[3] = {0, 0, 0};                      // resident blocks per SM of the 8-, 16- and 32-lane kernels
    cudaStream_t st = nullptr, st_in = nullptr, st_out = nullptr;   // kernels; H2D; D2H (copies overlap the kernels)
    KSlot slot[kRing];
    uint32_t *d_rowmx = nullptr;
    uint2 *d_bnd = nullptr;
    uint32_t *d_lutw = nullptr;                 // per warp: one LUT word per reference row
    uint8_t *d_qbuf = nullptr;                  // per warp: the reversed query prefix phase 1 reads
    size_t cap_rows = 0, cap_bnd = 0, cap_lutw = 0, cap_qbuf = 0;
    int next = 0;
};

}  // namespace

struct kswv_handle {
    kswv_params P;
    KParams K;
    std::vector<KDev> devs;
    kswv_gpu_stats stats;
    int force_width = 0;                        // KSWV_MIN_LANES: developer switch, at least this many lanes per pair
    char err[256];
};

namespace {

std::mutex g_err_mutex;

void set_err(kswv_handle *h, const char *fmt, ...) {
    std::lock_guard<std::mutex> lock(g_err_mutex);
    if (h->err[0]) return;                      // the first error of a call is the one reported
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof h->err, fmt, ap);
    va_end(ap);
}

#define KCU(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            set_err(h, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);        \
            return e_ == cudaErrorMemoryAllocation ? BSW_ERR_NOMEM : BSW_ERR_CUDA;                  \
        }                                                                                           \
    } while (0)

template <typename T>
int grow_dev(kswv_handle *h, T *&p, size_t &cap, size_t need) {
    if (need <= cap) return BSW_OK;
    if (p) KCU(cudaFree(p));
    p = nullptr; cap = 0;
    const size_t want = need + need / 4 + 256;
    KCU(cudaMalloc((void **)&p, want * sizeof(T)));
    cap = want;
    return BSW_OK;
}

template <typename T>
int grow_host(kswv_handle *h, T *&p, size_t have, size_t want) {
    (void)have;
    if (p) KCU(cudaFreeHost(p));
    p = nullptr;
    KCU(cudaHostAlloc((void **)&p, want * sizeof(T), cudaHostAllocDefault));
    return BSW_OK;
}

int ensure_dev(kswv_handle *h, KDev &d) {
    if (d.st) return BSW_OK;
    KCU(cudaSetDevice(d.id));
    KCU(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
    KCU(cudaStreamCreateWithFlags(&d.st_in, cudaStreamNonBlocking));
    KCU(cudaStreamCreateWithFlags(&d.st_out, cudaStreamNonBlocking));
    for (KSlot &s : d.slot) {
        KCU(cudaEventCreate(&s.ev_start));
        KCU(cudaEventCreate(&s.ev_stop));
        KCU(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        KCU(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        KCU(cudaEventCreateWithFlags(&s.ev_k, cudaEventDisableTiming));
        KCU(cudaMalloc((void **)&s.d_counter, 4 * sizeof(int)));
    }
    KCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ[0], kswv_phase0_kernel<8>, kKswvWarps * 32, 0));
    KCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ[1], kswv_phase0_kernel<16>, kKswvWarps * 32, 0));
    KCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ[2], kswv_phase0_kernel<32>, kKswvWarps * 32, 0));
    for (int &o : d.occ) o = std::max(1, std::min(o, kBlocksPerSm));
    d.warps = d.sms * kBlocksPerSm * kKswvWarps;
    d.groups = d.warps * 4;                     // 8 lanes per pair at the narrowest
    return BSW_OK;
}

int ensure_slot(kswv_handle *h, KSlot &s, size_t npairs, size_t ref_bytes, size_t qer_bytes, size_t gather_bytes) {
    if (npairs > s.cap_pairs) {
        const size_t want = npairs + npairs / 4 + 256;
        if (s.d_tasks) KCU(cudaFree(s.d_tasks));
        if (s.d_out) KCU(cudaFree(s.d_out));
        if (s.d_p1) KCU(cudaFree(s.d_p1));
        if (s.d_sort) KCU(cudaFree(s.d_sort));
        s.d_tasks = nullptr; s.d_out = nullptr; s.d_p1 = nullptr; s.d_sort = nullptr; s.cap_pairs = 0; s.cap_sort = 0;
        int rc = grow_host(h, s.h_tasks, 0, want);
        if (rc) return rc;
        rc = grow_host(h, s.h_out, 0, want);
        if (rc) return rc;
        KCU(cudaMalloc((void **)&s.d_tasks, want * sizeof(Task)));
        KCU(cudaMalloc((void **)&s.d_out, want * sizeof(Result)));
        KCU(cudaMalloc((void **)&s.d_p1, 4 * want * sizeof(uint32_t)));
        size_t bytes = 0;
        KCU(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, s.d_p1, s.d_p1, s.d_p1, s.d_p1, (int)want, 0, 28));
        KCU(cudaMalloc(&s.d_sort, bytes + 256));
        s.cap_sort = bytes + 256;
        s.cap_pairs = want;
    }
    int rc = grow_dev(h, s.d_ref, s.cap_ref, ref_bytes + 64);
    if (rc) return rc;
    rc = grow_dev(h, s.d_qer, s.cap_qer, qer_bytes + 64);
    if (rc) return rc;
    if (gather_bytes > s.cap_seq) {
        const size_t want = gather_bytes + gather_bytes / 4 + 4096;
        rc = grow_host(h, s.h_seq, 0, want);
        if (rc) return rc;
        s.cap_seq = want;
    }
    return BSW_OK;
}

void free_dev(KDev &d) {
    cudaSetDevice(d.id);
    for (KSlot &s : d.slot) {
        if (s.h_tasks) cudaFreeHost(s.h_tasks);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.h_seq) cudaFreeHost(s.h_seq);
        if (s.d_tasks) cudaFree(s.d_tasks);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_p1) cudaFree(s.d_p1);
        if (s.d_sort) cudaFree(s.d_sort);
        if (s.d_ref) cudaFree(s.d_ref);
        if (s.d_qer) cudaFree(s.d_qer);
        if (s.d_counter) cudaFree(s.d_counter);
        if (s.ev_start) cudaEventDestroy(s.ev_start);
        if (s.ev_stop) cudaEventDestroy(s.ev_stop);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_k) cudaEventDestroy(s.ev_k);
    }
    if (d.d_rowmx) cudaFree(d.d_rowmx);
    if (d.d_bnd) cudaFree(d.d_bnd);
    if (d.d_lutw) cudaFree(d.d_lutw);
    if (d.d_qbuf) cudaFree(d.d_qbuf);
    if (d.st) cudaStreamDestroy(d.st);
    if (d.st_in) cudaStreamDestroy(d.st_in);
    if (d.st_out) cudaStreamDestroy(d.st_out);
}

// waits for the slot's chunk, scatters its results to aln[regid], adds its kernel time
int drain_slot(kswv_handle *h, KDev &d, KSlot &s, const bsw_seqpair *pairs, kswv_result *aln, kswv_gpu_stats &S, int inner) {
    if (!s.busy) return BSW_OK;
    KCU(cudaSetDevice(d.id));
    KCU(cudaEventSynchronize(s.ev_done));
    float ms = 0;
    KCU(cudaEventElapsedTime(&ms, s.ev_start, s.ev_stop));
    S.kernel_ms += ms;
    const bsw_seqpair *p = pairs + s.first;
    const Result *r = s.h_out;
    static_assert(sizeof(Result) == sizeof(kswv_result), "Result is kswr_t");
#pragma omp parallel for schedule(static) num_threads(inner) if (s.count > 4096)
    for (int64_t i = 0; i < s.count; ++i) memcpy(&aln[p[i].regid], &r[i], sizeof(Result));
    s.busy = false;
    return BSW_OK;
}

}  // namespace

extern "C" {

int kswv_gpu_init(const kswv_params *params, int n_gpus, kswv_handle **out) {
    if (!params || !out) return BSW_ERR_ARG;
    *out = nullptr;
    const kswv_params &p = *params;
    if (p.match <= 0 || p.match > 127 || p.mismatch <= 0 || p.mismatch > 127 || p.o_del < 0 || p.o_ins < 0 ||
        p.e_del <= 0 || p.e_ins <= 0 || p.o_del + p.e_del > 127 || p.o_ins + p.e_ins > 127 ||
        p.match + p.mismatch > 254)
        return BSW_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return BSW_ERR_NO_DEVICE;
    const int want = n_gpus > 0 ? n_gpus : ndev;
    if (want > ndev) return BSW_ERR_NO_DEVICE;
    kswv_handle *h = new (std::nothrow) kswv_handle();
    if (!h) return BSW_ERR_NOMEM;
    h->P = p;
    h->K = make_kparams(p.o_del, p.e_del, p.o_ins, p.e_ins, p.match, p.mismatch);
    h->err[0] = 0;
    if (const char *e = getenv("KSWV_CHUNK_PAIRS")) {       // developer switch
        const long v = atol(e);
        if (v >= 1024 && v <= (1 << 22)) kChunkPairs = v;
    }
    if (const char *e = getenv("KSWV_MIN_LANES")) {
        const int w = atoi(e);
        if (w == 8 || w == 16 || w == 32) h->force_width = w;
    }
    memset(&h->stats, 0, sizeof h->stats);
    h->stats.n_gpus = want;
    h->devs.resize((size_t)want);
    for (int d = 0; d < want; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) != cudaSuccess || prop.major < 10) {
            delete h;
            return BSW_ERR_NO_DEVICE;           // kernels are sm_100a only
        }
        h->devs[(size_t)d].id = d;
        h->devs[(size_t)d].sms = prop.multiProcessorCount;
    }
    *out = h;
    return BSW_OK;
}

void kswv_gpu_free(kswv_handle *h) {
    if (!h) return;
    for (KDev &d : h->devs) free_dev(d);
    delete h;
}

int kswv_gpu_get_stats(const kswv_handle *h, kswv_gpu_stats *out) {
    if (!h || !out) return BSW_ERR_ARG;
    *out = h->stats;
    return BSW_OK;
}

const char *kswv_gpu_last_error(const kswv_handle *h) { return h ? h->err : "null handle"; }

static int kswv_batch_impl(kswv_handle *h, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                           int64_t n, kswv_result *aln);

int kswv_gpu_batch(kswv_handle *h, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                   int64_t n, kswv_result *aln) {
    const int rc = kswv_batch_impl(h, pairs, ref, qer, n, aln);
    if (rc != BSW_OK && h)      // nothing may stay in flight towards the caller's arrays after a failed call
        for (KDev &d : h->devs) {
            if (!d.st) continue;
            cudaSetDevice(d.id);
            cudaStreamSynchronize(d.st_in);
            cudaStreamSynchronize(d.st);
            cudaStreamSynchronize(d.st_out);
            for (KSlot &s : d.slot) s.busy = false;
        }
    return rc;
}

static int kswv_batch_impl(kswv_handle *h, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                           int64_t n, kswv_result *aln) {
    if (!h || n < 0 || (n > 0 && (!pairs || !ref || !qer || !aln))) return BSW_ERR_ARG;
    const auto t0 = std::chrono::steady_clock::now();
    kswv_gpu_stats &S = h->stats;
    S.chunks = 0; S.pairs = n; S.pairs8 = 0; S.cells = 0; S.h2d_bytes = 0; S.d2h_bytes = 0; S.kernel_launches = 0;
    S.gathered = 0; S.staged = 0; S.kernel_ms = 0; S.wall_ms = 0; S.lanes_per_pair = 0;
    S.host_check_ms = 0; S.host_prep_ms = 0; S.host_wait_ms = 0;
    auto ms_since = [](std::chrono::steady_clock::time_point a) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
    };
    h->err[0] = 0;
    if (n == 0) return BSW_OK;

    // ---- validation and totals (one host pass over the 72-byte records)
    int64_t bad = -1, n8 = 0, cells = 0;
    int bad_kind = 0;
    int maxT = 0, maxCols = 0, maxQ = 0;
#pragma omp parallel for schedule(static) reduction(+ : n8) reduction(+ : cells) reduction(max : maxT) reduction(max : maxCols) reduction(max : maxQ)
    for (int64_t i = 0; i < n; ++i) {
        const bsw_seqpair &sp = pairs[i];
        const bool byte = (sp.h0 & kXByte) != 0;
        int kind = 0;
        if (sp.len1 < 0 || sp.len2 < 0 || sp.len1 > 32767 || sp.len2 > 32767 || sp.idr < 0 || sp.idq < 0) kind = 1;
        else if (!byte && (int64_t)std::min(sp.len1, sp.len2) * h->P.match > 32767) kind = 1;
        else if (sp.regid < 0 || sp.regid >= n) kind = 2;
        if (kind) {
#pragma omp critical
            if (bad < 0 || i < bad) { bad = i; bad_kind = kind; }
            continue;
        }
        n8 += byte;
        const int nc = padded_cols(sp.len2, byte);
        cells += (int64_t)sp.len1 * nc;
        maxT = std::max(maxT, sp.len1);
        maxCols = std::max(maxCols, nc);
        maxQ = std::max(maxQ, sp.len2);
    }
    if (bad >= 0) {
        set_err(h, bad_kind == 2 ? "pair %lld: regid %d outside [0, n_pairs)" : "pair %lld outside the kswv domain (len1=%d len2=%d)",
                (long long)bad, bad_kind == 2 ? pairs[bad].regid : pairs[bad].len1, pairs[bad].len2);
        return bad_kind == 2 ? BSW_ERR_ARG : BSW_ERR_RANGE;
    }
    S.pairs8 = n8; S.cells = cells;
    S.host_check_ms = ms_since(t0);

    // ---- per-device scratch: one row-maximum column (and one boundary column for queries above 256 columns) per warp
    for (KDev &d : h->devs) {
        int rc = ensure_dev(h, d);
        if (rc) return rc;
        KCU(cudaSetDevice(d.id));
        const size_t rows = (size_t)((maxT + kScratchSlack + 3) & ~3);     // a multiple of 4 words: every group's scratch is 16-byte aligned
        rc = grow_dev(h, d.d_rowmx, d.cap_rows, rows * (size_t)d.groups);
        if (rc) return rc;
        if (maxCols > kPassCols) {
            rc = grow_dev(h, d.d_bnd, d.cap_bnd, rows * (size_t)d.warps);     // several passes: 32 lanes per pair only
            if (rc) return rc;
        }
        rc = grow_dev(h, d.d_lutw, d.cap_lutw, rows * (size_t)d.groups);
        if (rc) return rc;
        rc = grow_dev(h, d.d_qbuf, d.cap_qbuf, (size_t)(maxQ + 64) * (size_t)d.groups);
        if (rc) return rc;
    }
    const int scratch_rows = (maxT + kScratchSlack + 3) & ~3, scratch_q = maxQ + 64;

    // ---- chunks
    const int n_dev = (int)h->devs.size();
    // are the caller's sequence buffers page-locked (bsw_gpu_host_alloc / cudaHostRegister)? Then ranges are DMA'd in place.
    auto is_pinned = [](const void *p) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    const bool caller_pinned = is_pinned(ref) && is_pinned(qer);
    struct HostScratch { std::vector<uint32_t> order, bucket_start, bkt; };
    // one chunk of GPU d's share [lo, hi), starting at `first` (advanced past the chunk)
    auto one_chunk = [&](KDev &d, int64_t lo, int64_t hi, int64_t &first, kswv_gpu_stats &S, int inner, HostScratch &w) -> int {
    std::vector<uint32_t> &order = w.order, &bucket_start = w.bucket_start, &bkt = w.bkt;
    int rc = BSW_OK;
    for (int once = 0; once < 1 && first < hi; ++once) {
        // cut: at most kChunkPairs pairs / kChunkBytes sequence bytes
        // (a share smaller than three chunks is cut in three so that copies and kernels overlap, but never below
        // kMinChunk pairs: the GPU holds about 9500 pairs at a time at 8 lanes per pair, and chunks that do not fill it
        // only serialise their kernels -- 8000 pairs: 0.89 ms in two chunks, one launch is as long as one pair's rows)
        // Large shares take larger chunks (a sixth of the share, up to kMaxChunk): every chunk's kernels end with a
        // tail of idle SMs, and the GPU holds about 9500 pairs at once (400 000 pairs: 1556 GCUPS in chunks of 32 Ki,
        // 1632 in chunks of 64 Ki; four chunks of 128 Ki are faster still on the GPU but overlap the copies less).
        const int64_t big = std::min<int64_t>(kMaxChunk, std::max<int64_t>(kChunkPairs, (hi - lo) / 6));
        const int64_t target = std::max<int64_t>(kMinChunk, std::min<int64_t>(big, (hi - lo + 2) / 3));
        int64_t cnt = std::min<int64_t>(target, hi - first);
        int64_t rlo, rhi, qlo, qhi, rsum, qsum;
        bool ordered;
        const auto tc = std::chrono::steady_clock::now();
        for (;;) {      // extents of the chunk by a parallel reduction; halve it while its sequences exceed kChunkBytes
            int64_t a_rlo = INT64_MAX, a_rhi = 0, a_qlo = INT64_MAX, a_qhi = 0, a_rsum = 0, a_qsum = 0;
            int unordered = 0;
            const bsw_seqpair *cp0 = pairs + first;
#pragma omp parallel for schedule(static) reduction(min : a_rlo, a_qlo) reduction(max : a_rhi, a_qhi) \
    reduction(+ : a_rsum, a_qsum) reduction(| : unordered) num_threads(inner) if (cnt > 4096)
            for (int64_t i = 0; i < cnt; ++i) {
                const bsw_seqpair &sp = cp0[i];
                a_rlo = std::min<int64_t>(a_rlo, sp.idr); a_rhi = std::max<int64_t>(a_rhi, sp.idr + sp.len1);
                a_qlo = std::min<int64_t>(a_qlo, sp.idq); a_qhi = std::max<int64_t>(a_qhi, sp.idq + sp.len2);
                a_rsum += sp.len1; a_qsum += sp.len2;
                // in order: every pair starts at or after the end of the one before it, in both buffers
                if (i > 0 && (sp.idr < cp0[i - 1].idr + cp0[i - 1].len1 || sp.idq < cp0[i - 1].idq + cp0[i - 1].len2)) unordered = 1;
            }
            rlo = a_rlo; rhi = a_rhi; qlo = a_qlo; qhi = a_qhi; rsum = a_rsum; qsum = a_qsum; ordered = !unordered;
            if (rsum + qsum <= kChunkBytes || cnt == 1) break;
            cnt = (cnt + 1) / 2;
        }
        S.host_prep_ms += ms_since(tc);
        // one dense range per buffer (the production layout: mem_matesw_batch_pre appends, bwamem_pair.cpp:1006-1013)?
        const bool dense = ordered && (rhi - rlo) <= rsum + rsum / 8 + 4096 && (qhi - qlo) <= qsum + qsum / 8 + 4096 &&
                           (rhi - rlo) < (1ll << 32) && (qhi - qlo) < (1ll << 32);
        KCU(cudaSetDevice(d.id));
        KSlot &s = d.slot[d.next];
        d.next = (d.next + 1) % kRing;
        {
            const auto tw = std::chrono::steady_clock::now();
            rc = drain_slot(h, d, s, pairs, aln, S, inner);
            S.host_wait_ms += ms_since(tw);
        }
        if (rc) break;
        const auto tp = std::chrono::steady_clock::now();
        const size_t ref_bytes = dense ? (size_t)(rhi - rlo) : (size_t)rsum;
        const size_t qer_bytes = dense ? (size_t)(qhi - qlo) : (size_t)qsum;
        // a dense chunk in pageable caller memory is copied into the slot's page-locked staging by this thread(s): the
        // driver's own staging of a pageable source blocks the call for the whole copy at a third of the rate
        const bool stage = dense && !caller_pinned;
        rc = ensure_slot(h, s, (size_t)cnt, ref_bytes, qer_bytes, (dense && !stage) ? 0 : ref_bytes + qer_bytes + 64);
        if (rc) break;

        // Task order. Plain pairs first: no clamped arithmetic and at most 256 padded columns, the ones a group of
        // fewer than 32 lanes can take. Inside each class by decreasing strip-width class (strip_bucket) and
        // decreasing reference length (8-row bins): the pairs that share a warp get the same code and nearly the same
        // trip counts, and the largest DPs start first. One counting sort.
        const bsw_seqpair *cp = pairs + first;
        order.resize((size_t)cnt);
        int64_t n_plain = 0;
        int plain_cols = 0;
        {
            constexpr int kLenBins = 4096, kColBins = 17, kBuckets = 2 * kColBins * kLenBins;
            bucket_start.assign((size_t)kBuckets + 1, 0);
            auto bucket_of = [&](int64_t i, bool *plain, int *cols) {
                const bool byte = (cp[i].h0 & kXByte) != 0;
                const int nc = padded_cols(cp[i].len2, byte);
                const bool special = nc > kPassCols || needs_sat(h->K.a, h->K.shift, cp[i].len1, cp[i].len2, byte);
                const int cb = std::min(strip_bucket(nc), kColBins - 1);
                if (plain) { *plain = !special; *cols = nc; }
                return ((special ? 1 : 0) * kColBins + (kColBins - 1 - cb)) * kLenBins + (kLenBins - 1 - (cp[i].len1 >> 3));
            };
            bkt.resize((size_t)cnt);
            int64_t np = 0;
            int pc = 0;
#pragma omp parallel for schedule(static) reduction(+ : np) reduction(max : pc) num_threads(inner) if (cnt > 4096)
            for (int64_t i = 0; i < cnt; ++i) {      // the only pass over the 72-byte records; the rest works on 4-byte keys
                bool plain; int cols;
                bkt[(size_t)i] = (uint32_t)bucket_of(i, &plain, &cols);
                if (plain) { ++np; pc = std::max(pc, cols); }
            }
            n_plain = np; plain_cols = pc;
            for (int64_t i = 0; i < cnt; ++i) ++bucket_start[(size_t)bkt[(size_t)i] + 1];
            for (int b = 0; b < kBuckets; ++b) bucket_start[(size_t)b + 1] += bucket_start[(size_t)b];
            for (int64_t i = 0; i < cnt; ++i) order[bucket_start[(size_t)bkt[(size_t)i]]++] = (uint32_t)i;
        }
        if (dense) {
#pragma omp parallel for schedule(static) num_threads(inner) if (cnt > 4096)
            for (int64_t j = 0; j < cnt; ++j) {
                const uint32_t i = order[(size_t)j];
                const bsw_seqpair &sp = cp[i];
                s.h_tasks[j] = Task{(uint32_t)(sp.idr - rlo), (uint32_t)(sp.idq - qlo), sp.len1, sp.len2, sp.h0, (int32_t)i};
            }
        } else {
            // gather: offsets by a prefix sum in the caller's order, then the copies in parallel
            std::vector<uint32_t> roff((size_t)cnt), qoff((size_t)cnt);
            uint32_t ro = 0, qo = 0;
            for (int64_t i = 0; i < cnt; ++i) { roff[(size_t)i] = ro; qoff[(size_t)i] = qo; ro += (uint32_t)cp[i].len1; qo += (uint32_t)cp[i].len2; }
            uint8_t *gr = s.h_seq, *gq = s.h_seq + rsum;
#pragma omp parallel for schedule(static) num_threads(inner) if (cnt > 1024)
            for (int64_t i = 0; i < cnt; ++i) {
                memcpy(gr + roff[(size_t)i], ref + cp[i].idr, (size_t)cp[i].len1);
                memcpy(gq + qoff[(size_t)i], qer + cp[i].idq, (size_t)cp[i].len2);
            }
            for (int64_t j = 0; j < cnt; ++j) {
                const uint32_t i = order[(size_t)j];
                s.h_tasks[j] = Task{roff[i], qoff[i], cp[i].len1, cp[i].len2, cp[i].h0, (int32_t)i};
            }
            ++S.gathered;
        }
        S.host_prep_ms += ms_since(tp);
        // ---- enqueue
        KCU(cudaMemcpyAsync(s.d_tasks, s.h_tasks, sizeof(Task) * (size_t)cnt, cudaMemcpyHostToDevice, d.st_in));
        if (stage) {
            const size_t total = ref_bytes + qer_bytes, piece = 1u << 20;
            const int64_t npieces = (int64_t)((total + piece - 1) / piece);
#pragma omp parallel for schedule(static) num_threads(inner) if (npieces > 4)
            for (int64_t pc = 0; pc < npieces; ++pc) {
                // piece pc of the concatenation [ref range | qer range]
                size_t lo_b = (size_t)pc * piece, hi_b = std::min(total, lo_b + piece);
                if (lo_b < ref_bytes) {
                    const size_t e = std::min(hi_b, ref_bytes);
                    memcpy(s.h_seq + lo_b, ref + rlo + lo_b, e - lo_b);
                    lo_b = e;
                }
                if (lo_b < hi_b) memcpy(s.h_seq + lo_b, qer + qlo + (lo_b - ref_bytes), hi_b - lo_b);
            }
            if (ref_bytes) KCU(cudaMemcpyAsync(s.d_ref, s.h_seq, ref_bytes, cudaMemcpyHostToDevice, d.st_in));
            if (qer_bytes) KCU(cudaMemcpyAsync(s.d_qer, s.h_seq + ref_bytes, qer_bytes, cudaMemcpyHostToDevice, d.st_in));
            ++S.staged;
        } else if (dense) {
            if (ref_bytes) KCU(cudaMemcpyAsync(s.d_ref, ref + rlo, ref_bytes, cudaMemcpyHostToDevice, d.st_in));
            if (qer_bytes) KCU(cudaMemcpyAsync(s.d_qer, qer + qlo, qer_bytes, cudaMemcpyHostToDevice, d.st_in));
        } else {
            if (ref_bytes) KCU(cudaMemcpyAsync(s.d_ref, s.h_seq, ref_bytes, cudaMemcpyHostToDevice, d.st_in));
            if (qer_bytes) KCU(cudaMemcpyAsync(s.d_qer, s.h_seq + rsum, qer_bytes, cudaMemcpyHostToDevice, d.st_in));
        }
        KCU(cudaEventRecord(s.ev_in, d.st_in));
        KCU(cudaStreamWaitEvent(d.st, s.ev_in, 0));
        KCU(cudaMemsetAsync(s.d_counter, 0, 4 * sizeof(int), d.st));
        KCU(cudaEventRecord(s.ev_start, d.st));
        // lanes per pair for the plain pairs: 8 up to 160 padded columns (strips of up to 20), 16 up to 256, else 32
        int width = plain_cols <= group_cols(8) ? 8 : (plain_cols <= group_cols(16) ? 16 : 32);
        // a chunk that would leave most warps of the GPU idle at that width takes more lanes per pair instead: the
        // steps get shorter (fewer columns per lane), and a small batch is as long as its longest pair's row loops
        const int64_t warps = (int64_t)d.sms * kBlocksPerSm * kKswvWarps;
        while (width < 32 && n_plain * (width * 2) / 32 <= warps) width *= 2;
        if (h->force_width) width = std::max(width, h->force_width);
        // one class of tasks [off, off + nt): phase 0, the phase-1 tasks ordered by (strip width, te) so that the pairs
        // of a warp have similar trip counts, phase 1
        auto run_class = [&](int W, int64_t off, int64_t nt, int *counters) -> cudaError_t {
            if (nt <= 0) return cudaSuccess;
            const int per_block = kKswvWarps * (32 / W);
            const int blocks = (int)std::min<int64_t>((nt + per_block - 1) / per_block,
                                                      (int64_t)d.sms * d.occ[W == 8 ? 0 : (W == 16 ? 1 : 2)]);
            uint2 *bnd = (W == 32 && maxCols > kPassCols) ? d.d_bnd : nullptr;
            const Task *tasks = s.d_tasks + off;
            uint32_t *key = s.d_p1 + off, *val = s.d_p1 + s.cap_pairs + off;
            uint32_t *key_s = s.d_p1 + 2 * s.cap_pairs + off, *val_s = s.d_p1 + 3 * s.cap_pairs + off;
#define KSWV_P0(WW) kswv_phase0_kernel<WW><<<blocks, kKswvWarps * 32, 0, d.st>>>(h->K, tasks, (int)nt, s.d_ref, s.d_qer, s.d_out, \
                        key, val, d.d_rowmx, bnd, d.d_lutw, scratch_rows, counters)
#define KSWV_P1(WW) kswv_phase1_kernel<WW><<<blocks, kKswvWarps * 32, 0, d.st>>>(h->K, tasks, (int)nt, key_s, val_s, s.d_ref, s.d_qer, \
                        s.d_out, d.d_rowmx, bnd, d.d_lutw, d.d_qbuf, scratch_rows, scratch_q, counters + 1)
            if (W == 8) KSWV_P0(8); else if (W == 16) KSWV_P0(16); else KSWV_P0(32);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            size_t bytes = s.cap_sort;
            e = cub::DeviceRadixSort::SortPairsDescending(s.d_sort, bytes, key, key_s, val, val_s, (int)nt, 0, 28, d.st);
            if (e != cudaSuccess) return e;
            if (W == 8) KSWV_P1(8); else if (W == 16) KSWV_P1(16); else KSWV_P1(32);
#undef KSWV_P0
#undef KSWV_P1
            S.kernel_launches += 2;
            return cudaGetLastError();
        };
        KCU(run_class(width, 0, n_plain, s.d_counter));
        KCU(run_class(32, n_plain, cnt - n_plain, s.d_counter + 2));
        S.lanes_per_pair = width;
        KCU(cudaEventRecord(s.ev_stop, d.st));
        KCU(cudaEventRecord(s.ev_k, d.st));
        KCU(cudaStreamWaitEvent(d.st_out, s.ev_k, 0));
        KCU(cudaMemcpyAsync(s.h_out, s.d_out, sizeof(Result) * (size_t)cnt, cudaMemcpyDeviceToHost, d.st_out));
        KCU(cudaEventRecord(s.ev_done, d.st_out));
        s.busy = true; s.first = first; s.count = cnt;
        S.h2d_bytes += (int64_t)(sizeof(Task) * (size_t)cnt + ref_bytes + qer_bytes);
        S.d2h_bytes += (int64_t)(sizeof(Result) * (size_t)cnt);
        ++S.chunks;
        first += cnt;
    }
    return rc;
    };  // one_chunk

    // Up to four GPUs: one coordinator thread deals the chunks round by round (every GPU a contiguous share of the
    // pairs, no collective: pairs are independent); each chunk's passes over its records run on all host threads.
    // More GPUs: one host worker per GPU (an OpenMP team, kept between calls) with serial passes -- the coordinator's
    // own time per chunk becomes the limit there (N = 8 end to end: 8005 GCUPS coordinator, 9621 workers; N = 4: 5329
    // against 5097; N = 2: 2913 against 2761). KSWV_HOST_WORKERS=0/1 forces either.
    int rc = BSW_OK;
    const int max_threads = std::max(1, omp_get_max_threads());
    bool workers = n_dev > 4;
    if (const char *e = getenv("KSWV_HOST_WORKERS")) workers = e[0] == '1' && n_dev > 1;
    std::vector<int64_t> cur((size_t)n_dev), lo_((size_t)n_dev), hi_((size_t)n_dev);
    for (int g = 0; g < n_dev; ++g) { lo_[(size_t)g] = cur[(size_t)g] = n * g / n_dev; hi_[(size_t)g] = n * (g + 1) / n_dev; }
    auto drain_dev = [&](KDev &d, kswv_gpu_stats &St, int inner) -> int {
        int r = BSW_OK;
        const auto tw = std::chrono::steady_clock::now();
        for (int j = 0; j < kRing; ++j) {
            KSlot &s = d.slot[(d.next + j) % kRing];
            const int rc2 = drain_slot(h, d, s, pairs, aln, St, inner);
            if (r == BSW_OK) r = rc2;
        }
        St.host_wait_ms += ms_since(tw);
        return r;
    };
    if (!workers) {
        HostScratch w;
        for (bool more = true; more && rc == BSW_OK;) {
            more = false;
            for (int g = 0; g < n_dev && rc == BSW_OK; ++g)
                if (cur[(size_t)g] < hi_[(size_t)g]) {
                    rc = one_chunk(h->devs[(size_t)g], lo_[(size_t)g], hi_[(size_t)g], cur[(size_t)g], S, max_threads, w);
                    more = true;
                }
        }
        for (KDev &d : h->devs) {           // every GPU's slots, oldest first
            const int rc2 = drain_dev(d, S, max_threads);
            if (rc == BSW_OK) rc = rc2;
        }
    } else {
        std::vector<kswv_gpu_stats> part((size_t)n_dev);
        std::vector<int> rcs((size_t)n_dev, BSW_OK);
#pragma omp parallel for schedule(static, 1) num_threads(n_dev)
        for (int g = 0; g < n_dev; ++g) {
            HostScratch w;
            kswv_gpu_stats &P = part[(size_t)g];
            memset(&P, 0, sizeof P);
            int r = BSW_OK;
            while (r == BSW_OK && cur[(size_t)g] < hi_[(size_t)g])
                r = one_chunk(h->devs[(size_t)g], lo_[(size_t)g], hi_[(size_t)g], cur[(size_t)g], P, 1, w);
            const int r2 = drain_dev(h->devs[(size_t)g], P, 1);
            rcs[(size_t)g] = r != BSW_OK ? r : r2;
        }
        for (int g = 0; g < n_dev; ++g) {
            const kswv_gpu_stats &P = part[(size_t)g];
            S.chunks += P.chunks; S.h2d_bytes += P.h2d_bytes; S.d2h_bytes += P.d2h_bytes; S.kernel_launches += P.kernel_launches;
            S.gathered += P.gathered; S.staged += P.staged; S.kernel_ms += P.kernel_ms;
            S.host_prep_ms = std::max(S.host_prep_ms, P.host_prep_ms);      // the workers run side by side
            S.host_wait_ms = std::max(S.host_wait_ms, P.host_wait_ms);
            if (P.lanes_per_pair) S.lanes_per_pair = P.lanes_per_pair;
            if (rc == BSW_OK) rc = rcs[(size_t)g];
        }
    }
    S.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

}  // extern "C"
