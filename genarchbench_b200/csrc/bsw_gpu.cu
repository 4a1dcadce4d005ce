// libbsw_gpu.so: host pipeline + C ABI of the B200 bsw drop-in (include/bsw_gpu.h).
//
// Replaces BandedPairWiseSW::getScores16 / smithWatermanBatchWrapper16
// (/root/reference/benchmarks/bsw/src/bandedSWA.cpp:2679-2975) as called from the driver's ROI
// (main_banded.cpp:338-350). Where the reference pads, counting-sorts by len1, transposes AoS->SoA
// per 32 lanes and unsorts (bandedSWA.cpp:2726-2761, 2811-2880, 2940-2960), this library
//   1. cuts the caller's pair array into slabs,
//   2. per slab, in ONE pass over the caller's data on all host cores: validates, packs the bases
//      2 bits each (4 bits for pairs holding an ambiguous base) into pinned memory, writes a 16-byte
//      record per pair in the caller's order and histograms the query lengths (= the launch plan),
//   3. streams slabs through a ring of buffers per GPU: H2D, length binning ON THE DEVICE (a key
//      kernel + a radix sort of pair indices), one DP launch per length bin, D2H of 16-byte result
//      records already in the caller's order,
//   4. scatters the six outputs into the caller's SeqPair array.
// There is NO CPU implementation of the DP in here: without a CUDA device init fails.
#include "bsw_gpu.h"
#include "bsw_kernels.cuh"
#include "bsw_pack.h"

#include <omp.h>
#include <cub/device/device_radix_sort.cuh>
#include <nvtx3/nvToolsExt.h>   // header-only; ranges show up in Nsight Systems / ncu --nvtx (no-ops otherwise)

#include <atomic>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace bswk;

namespace {

constexpr int kRing = 3;                       // slabs in flight per GPU
constexpr int kBaseSlots = 8192;               // chunk bases of a packed slab (4096-pair chunks: up to 32 Mi pairs)
// pairs per slab (upper bound); BSW_SLAB_PAIRS overrides (tuning). The streaming path wants small slabs
// (host packing, GPU and scatter overlap slab by slab: 1 Mi measured best end to end), the resident path
// larger ones (more blocks per launch, shorter tails: 4 Mi is 4 % faster than 1 Mi on the device).
constexpr int64_t kSlabPairsBatch = 1 << 20, kSlabPairsStaged = 4 << 20;
inline int64_t slab_pairs(bool staged) {
    static const int64_t v = [] {
        const char *e = getenv("BSW_SLAB_PAIRS");
        const int64_t x = e ? atoll(e) : 0;
        return x >= 65536 ? x : 0;
    }();
    return v ? v : (staged ? kSlabPairsStaged : kSlabPairsBatch);
}
constexpr int64_t kSlabBases = 512ll << 20;    // bases per slab (upper bound)
constexpr size_t kMaxSmem = 227 * 1024 - 64;   // opt-in shared memory per block on sm_100, minus the kernel's static bytes
constexpr int kBinCols = 16;                   // query-length granularity of a launch bin
constexpr int kVersion = 3;
// device sort key of a pair (64 bits), descending order = launch order:
//   launch bin (len2 - 1) / 16 | holds an ambiguous base | (len2 - 1) % 16 | len1 (b1 bits) | h0 (b0 bits)
// with b1, b0 sized per slab (bins merged into the windowed launch: 1 | wide | len2 - 1 on top instead);
// at most 17 + 15 + 15 bits
constexpr int kKeyBits = 47;
constexpr int kMaxBins = BSW_MAX_SEQ_LEN / kBinCols + 2;

// NVTX range for the scope: the profiler-side view of the reference's ROI markers (main_banded.cpp:290-333 brackets
// its kernel loop for perf / VTune / FAPP / DynamoRIO; here the call and its per-slab host steps are named ranges)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

using Clock = std::chrono::steady_clock;
inline double ms_since(Clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

struct Launch {
    int first;       // first sorted position
    int n;           // pairs (the n_wide pairs holding an ambiguous base come first)
    int n_wide;
    int row_el, qs_words;   // per pair: uint4 row elements, u32 selector words
    int duo2_blk;           // > 0: extend_duo2 launch, 4-column blocks per thread
    int win_nk;             // > 0: windowed rows of win_nk elements (long query, narrow band)
    size_t smem;     // 0 => long kernel
    int64_t work;    // sum len1*len2, for ordering
};

// One slab = the unit that travels through a stream. Host side pinned, device side plain.
struct Slab {
    // capacity
    int64_t cap_pairs = 0;
    size_t cap_blob = 0, cap_scratch = 0;
    size_t cap_dblob = 0;            // capacity of d_blob (>= cap_blob: packed input with page-locked data grows it alone)
    // pinned host
    PairMeta *h_meta = nullptr;
    uint32_t *h_blob = nullptr;
    PairOut *h_out = nullptr;
    SlabStatsDev *h_stats = nullptr; // packed input: the statistics block bsw_rec_meta_kernel fills
    // device
    uint32_t *d_base = nullptr;      // packed input: first word of every 4096-pair chunk (kBaseSlots entries)
    uint32_t *d_chunk_total = nullptr;
    SlabStatsDev *d_stats = nullptr;
    cudaEvent_t ev_stats = nullptr;
    bool use_base = false;           // the current contents' PairMeta offsets are chunk-relative
    PairMeta *d_meta = nullptr;      // caller order
    uint32_t *d_blob = nullptr;
    PairOut *d_out = nullptr;
    unsigned char *d_scratch = nullptr;
    uint64_t *d_keys = nullptr;      // [2][cap_pairs] sort keys (in / out)
    uint32_t *d_ord = nullptr;       // [2][cap_pairs] pair indices (in / out): out = the binned order
    void *d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    // current contents
    int64_t lo = 0;          // first pair (caller order) of the slab
    int n = 0;               // pairs in the slab
    int n_dev = 0;           // pairs that go to the device (non-empty sequences)
    size_t blob_bytes = 0;
    std::vector<Launch> launches;
    std::vector<uint32_t> trivial;  // slab-local ids answered on the host (empty sequence)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_done = nullptr;
    bool busy = false;
    int key_b1 = 15, key_b0 = 16, key_bits = 48;   // sort key layout of the current contents (see sort_key)
    int long_bin0 = 0x7FFFFFFF;                    // launch bins >= this one are merged (windowed rows)
    bool fastm = false;      // every score of the slab times (match+1) fits int16: one-instruction M
    int max_sc = 0;          // largest h0 + len2 * match of the slab (bounds every H of its pairs)
    bool pinned = true;      // false for staged slabs (host side borrowed)
};

#ifndef BSW_AUX
#define BSW_AUX 8
#endif
constexpr int kAux = BSW_AUX;   // launch streams per GPU: length bins of a slab run concurrently

struct Device {
    int id = 0;
    Slab ring[kRing];
    std::vector<Slab *> staged;  // device-resident slabs of the staged API
    cudaStream_t aux[kAux] = {};
    cudaEvent_t aux_ev[kAux] = {};
    cudaEvent_t fork_ev = nullptr;
    bool attr_set[8] = {false, false, false, false, false, false, false, false};
    bool attr_set_key[2] = {false, false};
    bool attr_set_long[8] = {false, false, false, false, false, false, false, false};
    bool attr_set_win[8] = {false, false, false, false, false, false, false, false};
    bool attr_set_duo2[8] = {false, false, false, false, false, false, false, false};
};

}  // namespace

struct bsw_handle {
    bsw_params P;
    KParams K;
    bool sym = false;
    std::vector<Device> devs;
    bsw_gpu_stats stats;
    std::string err;
    // staged state
    int64_t staged_n = -1;
    int32_t staged_w = 0;
    // host scratch reused across slabs: per-thread (wide, bin) histograms
    std::vector<uint32_t> hist;
    // per call: pairs of the scalar class (score bound beyond int16) and invalid records, by index into the caller's
    // array; both are left out of the slabs and settled at the end of the call
    std::vector<int64_t> big, invalid;
};

namespace {

#define CU(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return BSW_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

inline int cuda_rc(bsw_handle *h, cudaError_t e, const char *what);

void free_slab(Slab &s) {
    if (s.pinned) {
        if (s.h_meta) cudaFreeHost(s.h_meta);
        if (s.h_blob) cudaFreeHost(s.h_blob);
        if (s.h_out) cudaFreeHost(s.h_out);
    }
    if (s.h_stats) cudaFreeHost(s.h_stats);
    if (s.d_base) cudaFree(s.d_base);
    if (s.d_chunk_total) cudaFree(s.d_chunk_total);
    if (s.d_stats) cudaFree(s.d_stats);
    if (s.ev_stats) cudaEventDestroy(s.ev_stats);
    if (s.d_meta) cudaFree(s.d_meta);
    if (s.d_blob) cudaFree(s.d_blob);
    if (s.d_out) cudaFree(s.d_out);
    if (s.d_scratch) cudaFree(s.d_scratch);
    if (s.d_keys) cudaFree(s.d_keys);
    if (s.d_ord) cudaFree(s.d_ord);
    if (s.d_sort_tmp) cudaFree(s.d_sort_tmp);
    if (s.ev_k0) cudaEventDestroy(s.ev_k0);
    if (s.ev_k1) cudaEventDestroy(s.ev_k1);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slab();
}

int ensure_slab(bsw_handle *h, Slab &s, int64_t pairs, size_t blob_bytes) {
    if (!s.stream) {
        // (highest priority: a slab's small copies, its record / key kernels and its sort get SM and copy-engine
        // slots ahead of the DP launches of the slabs before it, which run on the default-priority aux streams)
        int prio_lo = 0, prio_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CU(cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prio_hi));
        CU(cudaEventCreate(&s.ev_k0));
        CU(cudaEventCreate(&s.ev_k1));
        CU(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_stats, cudaEventDisableTiming));
        CU(cudaHostAlloc((void **)&s.h_stats, sizeof(SlabStatsDev), cudaHostAllocDefault));
        CU(cudaMalloc((void **)&s.d_base, sizeof(uint32_t) * kBaseSlots));
        CU(cudaMalloc((void **)&s.d_chunk_total, sizeof(uint32_t) * kBaseSlots));
        CU(cudaMalloc((void **)&s.d_stats, sizeof(SlabStatsDev)));
    }
    if (pairs > s.cap_pairs) {
        int64_t cap = std::max<int64_t>(pairs, 1024);
        if (s.pinned) {
            if (s.h_meta) cudaFreeHost(s.h_meta);
            if (s.h_out) cudaFreeHost(s.h_out);
            s.h_meta = nullptr; s.h_out = nullptr;
            CU(cudaHostAlloc((void **)&s.h_meta, sizeof(PairMeta) * cap, cudaHostAllocDefault));
            CU(cudaHostAlloc((void **)&s.h_out, sizeof(PairOut) * cap, cudaHostAllocDefault));
        }
        if (s.d_meta) cudaFree(s.d_meta);
        if (s.d_out) cudaFree(s.d_out);
        s.d_meta = nullptr; s.d_out = nullptr;
        CU(cudaMalloc((void **)&s.d_meta, sizeof(PairMeta) * cap));
        CU(cudaMalloc((void **)&s.d_out, sizeof(PairOut) * cap));
        if (s.d_keys) cudaFree(s.d_keys);
        if (s.d_ord) cudaFree(s.d_ord);
        if (s.d_sort_tmp) cudaFree(s.d_sort_tmp);
        s.d_keys = nullptr; s.d_ord = nullptr; s.d_sort_tmp = nullptr;
        CU(cudaMalloc((void **)&s.d_keys, sizeof(uint64_t) * 2 * cap));
        CU(cudaMalloc((void **)&s.d_ord, sizeof(uint32_t) * 2 * cap));
        s.sort_tmp_bytes = 0;
        CU(cub::DeviceRadixSort::SortPairsDescending(nullptr, s.sort_tmp_bytes, s.d_keys, s.d_keys + cap, s.d_ord,
                                                     s.d_ord + cap, (int)cap, 0, kKeyBits));
        CU(cudaMalloc(&s.d_sort_tmp, s.sort_tmp_bytes + 16));
        s.cap_pairs = cap;
    }
    if (blob_bytes > s.cap_blob) {
        size_t cap = std::max<size_t>(blob_bytes, 1 << 16);
        if (s.pinned) {
            if (s.h_blob) cudaFreeHost(s.h_blob);
            s.h_blob = nullptr;
            CU(cudaHostAlloc((void **)&s.h_blob, cap, cudaHostAllocDefault));
        }
        if (cap > s.cap_dblob) {
            if (s.d_blob) cudaFree(s.d_blob);
            s.d_blob = nullptr; s.cap_dblob = 0;
            CU(cudaMalloc((void **)&s.d_blob, cap));
            s.cap_dblob = cap;
        }
        s.cap_blob = cap;
    }
    return BSW_OK;
}

// device blob only (packed input read in place from page-locked memory needs no pinned staging copy)
int ensure_dblob(bsw_handle *h, Slab &s, size_t bytes) {
    if (bytes <= s.cap_dblob) return BSW_OK;
    if (s.d_blob) cudaFree(s.d_blob);
    s.d_blob = nullptr; s.cap_dblob = 0;
    const size_t cap = bytes + (bytes >> 4) + 4096;
    CU(cudaMalloc((void **)&s.d_blob, cap));
    s.cap_dblob = cap;
    return BSW_OK;
}

inline int cuda_rc(bsw_handle *h, cudaError_t e, const char *what) {
    if (e == cudaSuccess) return BSW_OK;
    h->err = std::string(what) + ": " + cudaGetErrorString(e);
    return BSW_ERR_CUDA;
}

// ---- host: per-slab preparation ----------------------------------------------------------------

inline size_t smem_need(int row_el, int qs_words) {
    return ((size_t)16 * row_el + (size_t)4 * qs_words) * kBlockPairs;
}
// BSW_WINDOW=0 sends every long pair to the warp-per-pair kernel (A/B against the windowed rows)
inline bool use_window() {
    static const bool v = !(getenv("BSW_WINDOW") && getenv("BSW_WINDOW")[0] == '0');
    return v;
}
// BSW_KEY=0: the short kernel's general argmax bookkeeping everywhere (A/B switch of extend_pair<.., KEY>)
inline bool use_key() {
    static const bool v = !(getenv("BSW_KEY") && getenv("BSW_KEY")[0] == '0');
    return v;
}
// extend_duo2 (two pairs per thread, bsw_duo.cuh) for the bins whose rows leave at least kDuo2MinWarps warps of it
// on an SM; BSW_DUO2=0 turns it off (A/B against the one-pair-per-thread kernel), BSW_DUO2_MINWARPS overrides
inline int duo2_min_warps() {
    static const int v = [] {
        if (!(getenv("BSW_DUO2") && getenv("BSW_DUO2")[0] == '1')) return 1 << 30;   // opt-in while it is being tuned
        const char *e = getenv("BSW_DUO2_MINWARPS");
        return e ? atoi(e) : 4;
    }();
    return v;
}

// Validates and packs slab [lo, lo+n) of the caller's arrays into s (host side), in one parallel pass
// over chunks of kChunk pairs:
//   loop A  (the chunk's SeqPair records, 72 B each): validate, size the chunk's blob;
//   reserve the chunk's words in the pinned blob with one atomic add (chunks land in any order);
//   loop B  (records now in cache): 2-bit pack query and target of each pair into a 16-byte aligned slot,
//           4-bit pack the pair again into an extra reservation if it holds an ambiguous base (the slot's
//           first word then holds that offset), write the pair's 16-byte PairMeta in the CALLER's order,
//           count it in the thread's (wide, len2 bin) histogram.
// The histogram is the launch plan: the device sorts pair indices by (bin, wide, len2, len1) itself.
// Returns BSW_OK, an error, or kRetry with s.blob_bytes = the capacity the slab really needs.
constexpr int kRetry = -1;
constexpr uint64_t kArenaWords = 32 << 10;   // 128 KiB of blob per reservation

// Results of an older slab that still have to go into the caller's SeqPair array. prepare_slab works
// them off inside its own parallel loop, interleaved with the packing chunks: the scatter is pure memory
// traffic (a read-for-ownership of every 72-byte record), the packer is half arithmetic, and together
// they fill the cores better than one after the other.
struct ScatterJob {
    const PairOut *out = nullptr;
    bsw_seqpair *dst = nullptr;
    int n = 0;
};
constexpr int kScatterChunk = 8192;

struct ScatterJobs {
    ScatterJob j[4];
    int first_chunk[5] = {0, 0, 0, 0, 0};   // chunk index ranges of the jobs
    int count = 0;
    void add(const ScatterJob &x) {
        if (count >= 4 || x.n <= 0) return;
        j[count] = x;
        first_chunk[count + 1] = first_chunk[count] + (x.n + kScatterChunk - 1) / kScatterChunk;
        ++count;
    }
    int chunks() const { return first_chunk[count]; }
};

inline void scatter_chunk(const ScatterJob &j, int c) {
    const int k1 = std::min(j.n, (c + 1) * kScatterChunk);
    for (int k = c * kScatterChunk; k < k1; ++k) {
        const PairOut &o = j.out[k];
        bsw_seqpair &p = j.dst[k];
        p.score = o.score; p.qle = o.qle; p.tle = o.tle;
        p.gtle = o.gtle; p.gscore = o.gscore; p.max_off = o.max_off;
    }
}

void plan_slab(bsw_handle *h, Slab &s, int T, int maxq);

int prepare_slab(bsw_handle *h, Slab &s, const bsw_seqpair *pairs, const uint8_t *ref,
                 const uint8_t *qer, int64_t lo, int n, const ScatterJobs &jobs) {
    constexpr int kChunk = 2048;
    const bsw_seqpair *pp = pairs + lo;
    bsw_gpu_stats &st = h->stats;
    s.lo = lo; s.n = n;
    s.use_base = false;
    s.launches.clear();
    s.trivial.clear();
    auto t0 = Clock::now();

    const int nchunks = (n + kChunk - 1) / kChunk;
    const int nscat = jobs.chunks();
    const int nitems = nchunks + nscat;
    const int T = omp_get_max_threads();
    const int match = h->P.match;
    const size_t cap_words = s.cap_blob / 4 > 16 ? s.cap_blob / 4 - 16 : 0;
    uint8_t *blob = reinterpret_cast<uint8_t *>(s.h_blob);
    std::atomic<uint64_t> cursor{0};
    int bad = 0, maxq = 0, maxsc = 0, maxt = 0, maxh = 0, overflow = 0;
    // exact blob need of the slab in 4-byte words (slots + 4-bit copies), summed over every pair whether it
    // fitted or not: what a capacity retry asks for
    uint64_t need_words = 0;
    h->hist.assign((size_t)T * 2 * kMaxBins, 0);
    std::vector<std::vector<uint32_t>> triv((size_t)T);
    std::vector<std::vector<int64_t>> bigv((size_t)T), badv((size_t)T);
    const int packer = pack_have_avx2() ? 2 : (pack_have_pext() ? 1 : 0);

#pragma omp parallel num_threads(T) reduction(| : bad) reduction(| : overflow) reduction(max : maxq) reduction(max : maxsc) reduction(max : maxt) reduction(max : maxh) reduction(+ : need_words)
    {
        const int t = omp_get_thread_num();
        bool full = false;             // the blob ran out: this thread only sizes its remaining pairs
        uint32_t *hist = h->hist.data() + (size_t)t * 2 * kMaxBins;
        uint64_t aoff = 0, aend = 0;   // this thread's current arena of the blob, in 4-byte words
        // work items: the packing chunks, with the scatter chunks of `job` spread evenly between them
#pragma omp for schedule(dynamic, 2)
        for (int it = 0; it < nitems; ++it) {
            const int sc_before = (int)((int64_t)it * nscat / nitems);
            if ((int)((int64_t)(it + 1) * nscat / nitems) > sc_before) {
                int jj = 0;
                while (sc_before >= jobs.first_chunk[jj + 1]) ++jj;
                scatter_chunk(jobs.j[jj], sc_before - jobs.first_chunk[jj]);
                continue;
            }
            const int c = it - sc_before;
            const int k0 = c * kChunk, k1 = std::min(n, k0 + kChunk);
            // The thread packs into its own arena of the pinned blob and reserves the next one with a single
            // atomic add when a slot does not fit (the unused tail of an arena, less than one slot, is simply
            // uploaded with the rest): no sizing pass over the records.
            for (int k = k0; k < k1; ++k) {
                const bsw_seqpair &sp = pp[k];
                const bool invalid = sp.len1 < 0 || sp.len2 < 0 || sp.len1 > BSW_MAX_SEQ_LEN || sp.len2 > BSW_MAX_SEQ_LEN || sp.h0 < 0;
                // (the class rule of bwa-mem2, bwamem.cpp:2218-2228: minval = h0 + min(len1, len2) * a bounds every H)
                if (invalid || (int64_t)sp.h0 + (int64_t)std::min(sp.len1, sp.len2) * match > 32767) {
                    // not for the int16 kernels: the slab carries an empty pair in its place (answered by the key
                    // kernel, overwritten when the call settles its scalar class / invalid records)
                    (invalid ? badv : bigv)[(size_t)t].push_back(lo + k);
                    PairMeta &m = s.h_meta[k];
                    m.off = 0; m.id = (uint32_t)k; m.len2 = 0; m.len1 = 0; m.h0 = 0; m.flags = 0;
                    triv[(size_t)t].push_back((uint32_t)k);
                    continue;
                }
                if (k + 4 < k1) {   // the records are read in order; pull the next sequences in early
                    __builtin_prefetch(qer + pp[k + 4].idq);
                    __builtin_prefetch(ref + pp[k + 4].idr);
                    __builtin_prefetch(ref + pp[k + 4].idr + 64);
                }
                const uint32_t qb = seq_bytes((uint32_t)sp.len2, false);
                const uint32_t sw = slot_words((uint32_t)sp.len2, (uint32_t)sp.len1);
                need_words += sw + 1;
                if (full) {
                    // sizing only: a pair holding an ambiguous base needs its 4-bit copy as well
                    bool amb = false;
                    for (int32_t x = 0; x < sp.len2 && !amb; ++x) amb = qer[sp.idq + x] > 3;
                    for (int32_t x = 0; x < sp.len1 && !amb; ++x) amb = ref[sp.idr + x] > 3;
                    if (amb) need_words += ((uint64_t)seq_bytes((uint32_t)sp.len2, true) + seq_bytes((uint32_t)sp.len1, true) + 15) / 16 * 4;
                    continue;
                }
                // (one spare word behind the slot: pack_pair_avx2 may write four zero bytes past it)
                if (aoff + sw + 1 > aend) {
                    const uint64_t want = std::max<uint64_t>(kArenaWords, sw + 1);
                    aoff = cursor.fetch_add(want, std::memory_order_relaxed);
                    aend = aoff + want;
                    if (aend > cap_words) {   // no further reservations from this thread (they would only inflate the cursor)
                        overflow |= 1; full = true; aend = aoff;
                        --k;                  // size this pair again in `full` mode (its 4-bit copy included)
                        need_words -= sw + 1;
                        continue;
                    }
                }
                const uint64_t off = aoff;
                aoff += sw;
                uint8_t *dst = blob + (size_t)off * 4;
                bool w1, w2 = false;
                if (packer == 2) {
                    w1 = pack_pair_avx2(qer + sp.idq, sp.len2, ref + sp.idr, sp.len1, dst, qb);
                } else if (packer == 1) {
                    w1 = pack2bit_pext(qer + sp.idq, sp.len2, dst);
                    w2 = pack2bit_pext(ref + sp.idr, sp.len1, dst + qb);
                } else {
                    w1 = pack2bit(qer + sp.idq, sp.len2, dst);
                    w2 = pack2bit(ref + sp.idr, sp.len1, dst + qb);
                }
                uint32_t wide = 0;
                if (w1 || w2) {
                    wide = 1;
                    const uint32_t wq = seq_bytes((uint32_t)sp.len2, true), wt = seq_bytes((uint32_t)sp.len1, true);
                    const uint64_t ww = ((uint64_t)wq + wt + 15) / 16 * 4;
                    need_words += ww;
                    const uint64_t woff = cursor.fetch_add(ww, std::memory_order_relaxed);
                    if (woff + ww > cap_words) {
                        overflow |= 1;
                    } else {
                        uint8_t *wd = blob + (size_t)woff * 4;
                        pack4bit(qer + sp.idq, sp.len2, wd);
                        pack4bit(ref + sp.idr, sp.len1, wd + wq);
                        const uint32_t w32 = (uint32_t)woff;
                        memcpy(dst, &w32, 4);
                    }
                }
                PairMeta &m = s.h_meta[k];
                m.off = (uint32_t)off;
                m.id = (uint32_t)k;
                m.len2 = (uint16_t)sp.len2; m.len1 = (uint16_t)sp.len1;
                m.h0 = (int16_t)sp.h0;
                m.flags = (uint16_t)wide;
                if (sp.len1 == 0 || sp.len2 == 0) {
                    triv[(size_t)t].push_back((uint32_t)k);
                } else {
                    hist[wide * kMaxBins + (uint32_t)(sp.len2 - 1) / kBinCols] += 1;
                    maxq = std::max(maxq, sp.len2);
                    maxsc = std::max(maxsc, sp.h0 + std::min(sp.len1, sp.len2) * match);
                    maxt = std::max(maxt, sp.len1);
                    maxh = std::max(maxh, sp.h0);
                }
            }
        }
    }
    if (bad) return BSW_ERR_RANGE;
    if (overflow) {
        // the exact need plus what the threads' partly used arenas may waste; checked against the 32-bit word
        // addressing of a slab blob only now (the cursor itself is meaningless after an overflow)
        const uint64_t want = need_words + (uint64_t)T * kArenaWords + 64;
        if (want > 0xFFFFFF00ull) return BSW_ERR_RANGE;
        s.blob_bytes = (size_t)want * 4 + 16;
        return kRetry;
    }
    const uint64_t total_words = cursor.load();
    if (total_words > 0xFFFFFF00ull) return BSW_ERR_RANGE;   // slab blobs are addressed in 32-bit words
    s.blob_bytes = (size_t)total_words * 4 + 16;
    memset(blob + (size_t)total_words * 4, 0, 16);
    s.fastm = (int64_t)maxsc * (match + 1) <= 32767;
    s.max_sc = (int)std::min<int64_t>(maxsc, INT32_MAX);
    s.key_b1 = bits_for((uint32_t)maxt);
    s.key_b0 = bits_for((uint32_t)maxh);
    for (int t = 0; t < T; ++t) {
        s.trivial.insert(s.trivial.end(), triv[(size_t)t].begin(), triv[(size_t)t].end());
        h->big.insert(h->big.end(), bigv[(size_t)t].begin(), bigv[(size_t)t].end());
        h->invalid.insert(h->invalid.end(), badv[(size_t)t].begin(), badv[(size_t)t].end());
    }
    s.n_dev = n - (int)s.trivial.size();
    st.host_pack_ms += ms_since(t0);
    t0 = Clock::now();
    plan_slab(h, s, T, maxq);
    st.host_plan_ms += ms_since(t0);
    return BSW_OK;
}

// The launch plan of a slab from the per-thread (wide, len2 bin) histograms in h->hist (T threads) and the slab's
// longest query; also fixes the sort key layout (s.key_b1 / key_b0 must be set).
void plan_slab(bsw_handle *h, Slab &s, int T, int maxq) {
    s.launches.clear();
    // ---- plan: one launch per query-length bin, longest first, the wide pairs of a bin in front
    // (== the device sort order: key descending). The bins that run on windowed rows need the same
    // shared memory whatever their query length, so they are merged into ONE launch at the front (the sort
    // key puts all of their wide pairs first): a slab of mixed lengths then issues a few hundred blocks
    // at once instead of dozens of 30-block launches that each last as long as their slowest pair.
    s.long_bin0 = 0x7FFFFFFF;
    if (s.n_dev > 0) {
        const int nbins = maxq / kBinCols + 1;
        const int nk = window_elems(h->K.w);
        const size_t ws = (size_t)20 * nk * kWinBlockPairs;
        // (bands whose window would leave fewer than four warps on an SM go to the warp-per-pair kernel)
        const bool can_window = (size_t)20 * nk * 128 <= kMaxSmem && use_window();
        auto bin_smem = [&](int b) {
            const int q_hi = std::min(maxq, (b + 1) * kBinCols);
            return smem_need(row_elems(q_hi), sel_words(q_hi));
        };
        // windowed rows from the first bin whose whole rows do not fit (BSW_WINDOW_EARLY=1, experiment: from
        // the first bin whose whole rows need more shared memory per pair than the window -- config 2
        // 37.4 ms against 34.6 ms, config 4 38.0 against 38.7)
        static const bool early = getenv("BSW_WINDOW_EARLY") && getenv("BSW_WINDOW_EARLY")[0] == '1';
        if (can_window)
            for (int b = 0; b < nbins; ++b)
                if (bin_smem(b) > (early ? (size_t)20 * nk * kBlockPairs : kMaxSmem)) { s.long_bin0 = b; break; }
        int p = 0;
        int nw_long = 0, nn_long = 0;
        for (int b = nbins - 1; b >= 0; --b) {
            int nw = 0, nn = 0;
            for (int t = 0; t < T; ++t) {
                nn += (int)h->hist[(size_t)t * 2 * kMaxBins + (size_t)b];
                nw += (int)h->hist[(size_t)t * 2 * kMaxBins + kMaxBins + (size_t)b];
            }
            if (b >= s.long_bin0) {
                nw_long += nw; nn_long += nn;
                if (b == s.long_bin0 && nw_long + nn_long > 0) {
                    Launch L;
                    L.first = 0; L.n = nw_long + nn_long; L.n_wide = nw_long; L.work = (int64_t)L.n * maxq * 2 * h->K.w;
                    L.row_el = L.qs_words = L.duo2_blk = 0;
                    L.smem = ws; L.win_nk = nk;
                    s.launches.push_back(L);
                    p = L.n;
                }
                continue;
            }
            if (nw + nn == 0) continue;
            const int q_hi = std::min(maxq, (b + 1) * kBinCols);
            Launch L;
            L.first = p; L.n = nw + nn; L.n_wide = nw; L.work = (int64_t)(nw + nn) * q_hi * q_hi;
            L.row_el = row_elems(q_hi);
            L.qs_words = sel_words(q_hi);
            L.smem = bin_smem(b);
            L.win_nk = 0;
            L.duo2_blk = 0;
            {
                const size_t d2 = (size_t)duo2_thread_bytes(q_hi) * kDuo2Threads;
                if (d2 <= kMaxSmem && (int)(kMaxSmem / d2) * (kDuo2Threads / 32) >= duo2_min_warps()) {
                    L.duo2_blk = duo_blocks(q_hi);
                    L.smem = d2;
                }
            }
            if (L.smem > kMaxSmem) {
                // whole rows do not fit and the band is too wide for a window: one warp per pair. Its shared
                // memory per pair is small against the SM's, so neighbouring bins share a launch as long as
                // the longest query of the launch is at most twice theirs (one launch per length octave).
                L.smem = 0;
                if (!s.launches.empty()) {
                    Launch &prev = s.launches.back();
                    if (prev.smem == 0 && prev.win_nk == 0 && prev.first + prev.n == p && prev.row_el <= 2 * L.row_el) {
                        prev.n += L.n; prev.n_wide += L.n_wide;
                        p += nw + nn;
                        continue;
                    }
                }
            }
            s.launches.push_back(L);
            p += nw + nn;
        }
    }
    // sort key layout (see sort_key)
    s.key_bits = s.key_b1 + s.key_b0 + (s.long_bin0 != 0x7FFFFFFF ? 17 : bits_for(((((uint32_t)std::max(maxq, 1) - 1) >> 4) << 5) | 31u));
}

// prepare_slab with the (rare) capacity retry: the first pass over a slab whose blob does not fit the
// ring slot reports the exact size, the slot grows, the pass runs again.
int prepare_slab_fit(bsw_handle *h, Slab &s, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                     int64_t lo, int n, size_t blob_guess, ScatterJobs jobs = ScatterJobs()) {
    auto t0 = Clock::now();
    blob_guess += (size_t)omp_get_max_threads() * kArenaWords * 4;   // every thread may leave an arena partly unused
    if (n > s.cap_pairs && jobs.count > 0) {
        // the slot's pair-sized buffers are about to be reallocated, and a queued scatter job may still read
        // this slot's h_out: work the jobs off first
        const int nc = jobs.chunks();
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < nc; ++c) {
            int jj = 0;
            while (c >= jobs.first_chunk[jj + 1]) ++jj;
            scatter_chunk(jobs.j[jj], c - jobs.first_chunk[jj]);
        }
        jobs = ScatterJobs();
    }
    int rc = ensure_slab(h, s, n, blob_guess);
    h->stats.host_alloc_ms += ms_since(t0);
    if (rc) return rc;
    for (int attempt = 0; attempt < 3; ++attempt) {
        rc = prepare_slab(h, s, pairs, ref, qer, lo, n, jobs);
        if (rc != kRetry) return rc;
        jobs = ScatterJobs();   // done (a scatter is idempotent anyway)
        t0 = Clock::now();
        rc = ensure_slab(h, s, n, s.blob_bytes + (s.blob_bytes >> 4) + 4096);
        h->stats.host_alloc_ms += ms_since(t0);
        if (rc) return rc;
    }
    return BSW_ERR_NOMEM;
}

// ---- device side of a slab ---------------------------------------------------------------------

int ensure_aux(bsw_handle *h, Device &dev) {
    if (dev.aux[0]) return BSW_OK;
    for (int j = 0; j < kAux; ++j) {
        CU(cudaStreamCreateWithFlags(&dev.aux[j], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&dev.aux_ev[j], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&dev.fork_ev, cudaEventDisableTiming));
    return BSW_OK;
}

// The kernel instantiations, indexed [fastm][sym][count].
typedef void (*ShortFn)(const PairMeta *, const uint32_t *, const uint32_t *, PairOut *, int, int, KParams, int, int);
typedef void (*WinFn)(const PairMeta *, const uint32_t *, const uint32_t *, PairOut *, int, int, KParams, int);
typedef void (*LongFn)(const PairMeta *, const uint32_t *, const uint32_t *, PairOut *, int, int, KParams, int, int);
typedef void (*Duo2Fn)(const PairMeta *, const uint32_t *, const uint32_t *, PairOut *, int, int, KParams, int);
template <int I> struct KernelTable {
    static void fill(ShortFn *sf, LongFn *lf, WinFn *wf) {
        wf[I] = bsw_win_kernel<(I & 4) != 0, (I & 2) != 0, (I & 1) != 0>;
        sf[I] = bsw_short_kernel<(I & 4) != 0, (I & 2) != 0, (I & 1) != 0>;
        lf[I] = bsw_long_kernel<(I & 4) != 0, (I & 2) != 0, (I & 1) != 0>;
        KernelTable<I - 1>::fill(sf, lf, wf);
    }
};
template <> struct KernelTable<-1> { static void fill(ShortFn *, LongFn *, WinFn *) {} };
inline int kernel_index(bool fastm, bool sym, bool count) {
    return (fastm ? 4 : 0) | (sym ? 2 : 0) | (count ? 1 : 0);
}

// Enqueues every launch of the given slabs: work forks from `main` onto the device's aux streams
// (bins of different lengths overlap; the few-block long bins no longer leave the GPU idle) and
// joins back into `main`. Long-kernel launches share a per-slab scratch and stay on aux[0].
int launch_slabs(bsw_handle *h, Device &dev, cudaStream_t main, Slab *const *slabs, int nslabs, bool count = false) {
    static ShortFn short_fn[8];
    static LongFn long_fn[8];
    static WinFn win_fn[8];
    static const ShortFn short_key_fn[2] = {bsw_short_kernel<true, false, false, true>,
                                            bsw_short_kernel<true, true, false, true>};
    // [fastm * 2 + sym] and, keyed, [sym]
    static const Duo2Fn duo2_fn[4] = {bsw_duo2_kernel<false, false, false>, bsw_duo2_kernel<false, true, false>,
                                      bsw_duo2_kernel<true, false, false>, bsw_duo2_kernel<true, true, false>};
    static const Duo2Fn duo2_key_fn[2] = {bsw_duo2_kernel<true, false, true>, bsw_duo2_kernel<true, true, true>};
    // (a function-local static with an initialiser is filled exactly once, also with two handles on two threads)
    static const bool filled = [] { KernelTable<7>::fill(short_fn, long_fn, win_fn); return true; }();
    (void)filled;
    int rc = ensure_aux(h, dev);
    if (rc) return rc;
    CU(cudaEventRecord(dev.fork_ev, main));
    for (int j = 0; j < kAux; ++j) CU(cudaStreamWaitEvent(dev.aux[j], dev.fork_ev, 0));
    int rr = 0;
    for (int i = 0; i < nslabs; ++i) {
        Slab &s = *slabs[i];
        for (const Launch &L : s.launches) {
            const int grid = (launch_threads(L.n_wide, L.n - L.n_wide) + kBlockPairs - 1) / kBlockPairs;
            const int ki = kernel_index(s.fastm, h->sym, count);
            if (L.win_nk) {
                const int wgrid = (launch_threads(L.n_wide, L.n - L.n_wide) + kWinBlockPairs - 1) / kWinBlockPairs;
                if (!dev.attr_set_win[ki]) {
                    CU(cudaFuncSetAttribute(win_fn[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
                    dev.attr_set_win[ki] = true;
                }
                cudaStream_t st = dev.aux[rr++ % kAux];
                win_fn[ki]<<<wgrid, kWinBlockPairs, L.smem, st>>>(s.d_meta, s.d_ord + s.cap_pairs + L.first, s.d_blob,
                                                              s.d_out, L.n_wide, L.n - L.n_wide, h->K, L.win_nk);
            } else if (L.smem && L.duo2_blk && !count) {
                // keyed row argmax when the launch's scores and column indices (up to the end of the last block)
                // share 16 bits
                const int kbits = bits_for((uint32_t)(4 * L.duo2_blk - 1));
                const bool keyed = s.fastm && use_key() && kbits < 16 && s.max_sc < (1 << (16 - kbits));
                const int fi = keyed ? (h->sym ? 1 : 0) : ((s.fastm ? 2 : 0) | (h->sym ? 1 : 0));
                Duo2Fn fn = keyed ? duo2_key_fn[fi] : duo2_fn[fi];
                bool &attr = dev.attr_set_duo2[keyed ? 4 + fi : fi];
                if (!attr) {
                    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
                    attr = true;
                }
                KParams K = h->K;
                if (keyed) { K.kbits = (uint32_t)kbits; K.kkey = 1u << kbits; }
                cudaStream_t st = dev.aux[rr++ % kAux];
                const int threads = (L.n + 1) / 2;
                fn<<<(threads + kDuo2Threads - 1) / kDuo2Threads, kDuo2Threads, L.smem, st>>>(
                    s.d_meta, s.d_ord + s.cap_pairs + L.first, s.d_blob, s.d_out, L.n_wide, L.n - L.n_wide, K, L.duo2_blk);
                if (keyed) h->stats.pairs_keyed += L.n;
                h->stats.pairs_duo += L.n;
            } else if (L.smem) {
                // keyed row argmax when the launch's scores and group indices share 16 bits
                // (a query of this launch has at most 4 * row_el - 1 bases: groups 0 .. 2 * row_el - 1)
                // and a row never spans more than band + 2 <= w + 2 groups (beg >= i - band, end <= i + band + 1,
                // the first group rounded down to a 4-column boundary)
                const int kbits = bits_for((uint32_t)std::min<int64_t>(2 * L.row_el - 1, BSW_KEY_REL ? (int64_t)h->K.w + 2 : INT32_MAX));
                const bool keyed = s.fastm && !count && use_key() && kbits < 16 && s.max_sc < (1 << (16 - kbits));
                ShortFn fn = keyed ? short_key_fn[h->sym ? 1 : 0] : short_fn[ki];
                const size_t smem1 = L.duo2_blk ? smem_need(L.row_el, L.qs_words) : L.smem;   // (a COUNT run of a duo2 bin)
                bool &attr = keyed ? dev.attr_set_key[h->sym ? 1 : 0] : dev.attr_set[ki];
                if (!attr) {
                    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
                    attr = true;
                }
                KParams K = h->K;
                if (keyed) { K.kbits = (uint32_t)kbits; K.kkey = 1u << kbits; }
                cudaStream_t st = dev.aux[rr++ % kAux];
                fn<<<grid, kBlockPairs, smem1, st>>>(s.d_meta, s.d_ord + s.cap_pairs + L.first, s.d_blob,
                                                      s.d_out, L.n_wide, L.n - L.n_wide, K, L.row_el, L.qs_words);
                if (keyed) h->stats.pairs_keyed += L.n;
            } else {
                // one warp per pair; as many pairs per block as fit (at most 8)
                const int q_hi = 4 * L.row_el - 4;     // row_el = (qlen + 4) >> 2 was sized from the bin's longest query
                const int pb = (int)warp_pair_bytes(q_hi + 3);
                const int wpb = std::max(1, std::min(8, (int)(kMaxSmem / (size_t)pb)));
                if (!dev.attr_set_long[ki]) {
                    CU(cudaFuncSetAttribute(long_fn[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
                    dev.attr_set_long[ki] = true;
                }
                cudaStream_t st = dev.aux[rr++ % kAux];
                long_fn[ki]<<<(L.n + wpb - 1) / wpb, 32 * wpb, (size_t)wpb * pb, st>>>(
                    s.d_meta, s.d_ord + s.cap_pairs + L.first, s.d_blob, s.d_out, L.n_wide, L.n - L.n_wide, h->K,
                    L.row_el, pb);
            }
            CU(cudaGetLastError());
            h->stats.kernel_launches++;
            if (L.smem && !L.win_nk) h->stats.pairs_short += L.n; else h->stats.pairs_long += L.n;
        }
    }
    for (int j = 0; j < kAux; ++j) {
        CU(cudaEventRecord(dev.aux_ev[j], dev.aux[j]));
        CU(cudaStreamWaitEvent(main, dev.aux_ev[j], 0));
    }
    return BSW_OK;
}

// Length binning of a slab on its device: keys + identity, then the radix sort (descending).
int bin_slab(bsw_handle *h, Slab &s, cudaStream_t st) {
    if (s.n_dev == 0) return BSW_OK;
    const int n = s.n;
    bsw_key_kernel<<<(n + 255) / 256, 256, 0, st>>>(s.d_meta, n, s.d_keys, s.d_ord, s.key_b1, s.key_b0, s.long_bin0,
                                                    s.d_out, s.use_base ? s.d_base : nullptr);
    CU(cudaGetLastError());
    h->stats.kernel_launches++;
    size_t tmp = s.sort_tmp_bytes;
    CU(cub::DeviceRadixSort::SortPairsDescending(s.d_sort_tmp, tmp, s.d_keys, s.d_keys + s.cap_pairs, s.d_ord,
                                                 s.d_ord + s.cap_pairs, n, 0, std::min(s.key_bits, kKeyBits), st));
    return BSW_OK;
}

int launch_slab(bsw_handle *h, Device &dev, Slab &s, bool count = false) {
    Slab *one = &s;
    return launch_slabs(h, dev, s.stream, &one, 1, count);
}

// long launches of one slab must not share the scratch concurrently: they are on one stream -> serial.

int upload_slab(bsw_handle *h, Slab &s) {
    if (s.n_dev == 0) return BSW_OK;
    CU(cudaMemcpyAsync(s.d_meta, s.h_meta, sizeof(PairMeta) * (size_t)s.n, cudaMemcpyHostToDevice, s.stream));
    CU(cudaMemcpyAsync(s.d_blob, s.h_blob, s.blob_bytes, cudaMemcpyHostToDevice, s.stream));
    h->stats.h2d_bytes += (int64_t)(sizeof(PairMeta) * (size_t)s.n + s.blob_bytes);
    return BSW_OK;
}

int download_slab(bsw_handle *h, Slab &s) {
    if (s.n_dev == 0) return BSW_OK;
    CU(cudaMemcpyAsync(s.h_out, s.d_out, sizeof(PairOut) * (size_t)s.n, cudaMemcpyDeviceToHost, s.stream));
    h->stats.d2h_bytes += (int64_t)(sizeof(PairOut) * (size_t)s.n);
    return BSW_OK;
}

// writes the six outputs of slab s into the caller's array
void scatter_slab(bsw_handle *h, const Slab &s, const PairOut *out, bsw_seqpair *pairs) {
    bsw_seqpair *pp = pairs + s.lo;
    const int n = s.n;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; ++k) {
        const PairOut &o = out[k];
        bsw_seqpair &p = pp[k];
        p.score = o.score; p.qle = o.qle; p.tle = o.tle;
        p.gtle = o.gtle; p.gscore = o.gscore; p.max_off = o.max_off;
    }
    // empty target or query: the DP loop never runs (bandedSWA.cpp:181 with tlen == 0 / end == 0)
    for (uint32_t k : s.trivial) {
        bsw_seqpair &p = pp[k];
        p.score = p.h0; p.qle = 0; p.tle = 0; p.gtle = 0; p.gscore = -1; p.max_off = 0;
    }
    (void)h;
}

// Slab boundaries over [0, n): by pair count and by (estimated) bases, at a granularity of 64 Ki
// pairs. The lengths are SAMPLED (every 64th record: one cache line in 72 instead of a sweep over the
// whole array); blob_guess[i] is the pinned-blob capacity to try first for slab i -- prepare_slab
// reports the exact need if the estimate was short.
inline bool use_taper() {
    static const bool v = !(getenv("BSW_TAPER") && getenv("BSW_TAPER")[0] == '0');
    return v;
}
constexpr int64_t kCutGroup = 1 << 16;   // slab boundaries fall on multiples of this many pairs
// pairs of the slab that starts with `rem` pairs still to go (streaming calls; see cut_slabs)
inline int64_t slab_target(int64_t rem, int64_t full) {
    if (!use_taper() || rem > full + full / 2) return full;
    return std::max<int64_t>(2 * kCutGroup, (rem / 2 + kCutGroup - 1) / kCutGroup * kCutGroup);
}
void cut_slabs(const bsw_seqpair *pairs, int64_t n, std::vector<int64_t> &cuts, std::vector<size_t> &blob_guess,
               bool staged, std::vector<int64_t> *cuts_flat = nullptr, std::vector<size_t> *guess_flat = nullptr,
               int n_gpus = 1) {
    constexpr int64_t G = kCutGroup, S = 64;
    const int64_t ng = (n + G - 1) / G;
    std::vector<int64_t> bases((size_t)ng, 0);
    std::vector<double> work((size_t)ng, 0.0);   // estimated DP cells (len1 * len2) per group
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ng; ++c) {
        int64_t b = 0, cnt = 0;
        double wk = 0;
        const int64_t hi = std::min(n, (c + 1) * G);
        for (int64_t k = c * G; k < hi; k += S, ++cnt) {
            const int64_t l1 = std::min<int64_t>(std::max<int64_t>(pairs[k].len1, 0), BSW_MAX_SEQ_LEN);
            const int64_t l2 = std::min<int64_t>(std::max<int64_t>(pairs[k].len2, 0), BSW_MAX_SEQ_LEN);
            b += l1 + l2;
            wk += (double)l1 * (double)l2;
        }
        bases[(size_t)c] = cnt ? b * (hi - c * G) / cnt : 0;
        work[(size_t)c] = cnt ? wk * (double)(hi - c * G) / (double)cnt : 0.0;
    }
    // Streaming calls end with a taper: what follows the last slab's packing -- its upload, kernels, download
    // and scatter -- is not overlapped with anything, so the last 1.5 slabs' worth of pairs is cut into
    // halves, quarters, eighths (BSW_TAPER=0: off). That only pays while the host is the slower side, so the
    // untapered plan is returned as well (cuts_flat) and bsw_gpu_batch picks where the two diverge.
    const int64_t full = slab_pairs(staged);
    auto plan = [&](bool taper, std::vector<int64_t> &cv, std::vector<size_t> &gv) {
        cv.clear();
        gv.clear();
        cv.push_back(0);
        int64_t acc = 0, cnt = 0, target = full;
        auto close = [&](int64_t hi) {
            cv.push_back(hi);
            // 2 bits per base + up to 19 bytes of padding per pair + 15 % for sampling error and wide pairs
            gv.push_back((size_t)((acc / 4 + 20 * cnt) * 115 / 100) + 65536);
            acc = 0; cnt = 0;
        };
        for (int64_t c = 0; c < ng; ++c) {
            const int64_t hi = std::min(n, (c + 1) * G);
            if (taper && cnt == 0) target = slab_target(n - c * G, full);
            acc += bases[(size_t)c];
            cnt += hi - c * G;
            if (cnt >= target || acc >= kSlabBases) close(hi);
        }
        if (cv.back() != n) close(n);
    };
    if (staged && n_gpus > 1) {
        // Resident batch over several GPUs: contiguous slabs of (estimated) equal DP work, a multiple of
        // the GPU count of them, dealt round-robin -- every GPU gets the same number of slabs and the same work.
        double total = 0;
        for (double x : work) total += x;
        int64_t nsl = std::max<int64_t>((n + full - 1) / full, 1);
        nsl = (nsl + n_gpus - 1) / n_gpus * n_gpus;
        cuts.clear(); blob_guess.clear();
        cuts.push_back(0);
        int64_t acc = 0, cnt = 0;
        double wacc = 0;
        auto close = [&](int64_t hi) {
            cuts.push_back(hi);
            blob_guess.push_back((size_t)((acc / 4 + 20 * cnt) * 115 / 100) + 65536);
            acc = 0; cnt = 0;
        };
        for (int64_t c = 0; c < ng; ++c) {
            const int64_t hi = std::min(n, (c + 1) * G);
            acc += bases[(size_t)c];
            cnt += hi - c * G;
            wacc += work[(size_t)c];
            const int64_t k = (int64_t)cuts.size();     // closing would end slab k - 1
            if ((k < nsl && wacc >= total * (double)k / (double)nsl) || acc >= kSlabBases || cnt >= 2 * full) close(hi);
        }
        if (cuts.back() != n) close(n);
        return;
    }
    plan(!staged && use_taper(), cuts, blob_guess);
    if (cuts_flat && guess_flat) plan(false, *cuts_flat, *guess_flat);
}

// Waits for a busy slab, books its kernel time, answers its trivial pairs and hands its result records
// over as a scatter job (the slab is free afterwards: only h_out is still read, and nothing writes it
// before the job ran).
int claim_slab(bsw_handle *h, Slab &s, bsw_seqpair *pairs, double *kernel_ms_acc, ScatterJobs &jobs) {
    if (!s.busy) return BSW_OK;
    {
        auto tw = Clock::now();
        CU(cudaEventSynchronize(s.ev_done));
        h->stats.host_wait_ms += ms_since(tw);
    }
    if (s.n_dev) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
        *kernel_ms_acc += ms;
    }
    bsw_seqpair *pp = pairs + s.lo;
    ScatterJob j;
    j.out = s.h_out; j.dst = pp; j.n = s.n_dev ? s.n : 0;
    jobs.add(j);
    for (uint32_t k : s.trivial) {   // empty target or query: the DP loop never runs (bandedSWA.cpp:181)
        bsw_seqpair &p = pp[k];
        p.score = p.h0; p.qle = 0; p.tle = 0; p.gtle = 0; p.gscore = -1; p.max_off = 0;
    }
    s.busy = false;
    return BSW_OK;
}

int finish_slab(bsw_handle *h, Slab &s, bsw_seqpair *pairs, double *kernel_ms_acc) {
    if (!s.busy) return BSW_OK;
    {
        auto tw = Clock::now();
        CU(cudaEventSynchronize(s.ev_done));
        h->stats.host_wait_ms += ms_since(tw);
    }
    if (s.n_dev) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
        *kernel_ms_acc += ms;
    }
    auto t0 = Clock::now();
    scatter_slab(h, s, s.h_out, pairs);
    h->stats.host_scatter_ms += ms_since(t0);
    s.busy = false;
    return BSW_OK;
}

}  // namespace

// End of a bsw_gpu_batch call: the pairs its slabs left out.
//   * scalar class (score bound beyond int16; bwamem.cpp:2218-2228, 2384-2390): one launch of bsw_big_kernel over
//     byte-per-base copies of their sequences, int32 outputs;
//   * invalid records (negative length, length > BSW_MAX_SEQ_LEN, negative h0): the six outputs are set to -1,
//     every other pair of the batch is computed, the call reports BSW_ERR_RANGE.
static int settle_classes(bsw_handle *h, bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer) {
    bsw_gpu_stats &st = h->stats;
    int rc = BSW_OK;
    if (!h->big.empty()) {
        std::sort(h->big.begin(), h->big.end());
        const size_t nb = h->big.size();
        std::vector<BigMeta> meta(nb);
        uint64_t sbytes = 0, rows = 0;
        for (size_t a = 0; a < nb; ++a) {
            const bsw_seqpair &sp = pairs[h->big[a]];
            BigMeta &m = meta[a];
            m.toff = sbytes; sbytes += (uint64_t)sp.len1;
            m.qoff = sbytes; sbytes += (uint64_t)sp.len2;
            m.soff = rows; rows += 2ull * ((uint64_t)sp.len2 + 2);
            m.len1 = sp.len1; m.len2 = sp.len2; m.h0 = sp.h0; m.pad = 0;
        }
        std::vector<uint8_t> seq(sbytes + 16);
        for (size_t a = 0; a < nb; ++a) {
            const bsw_seqpair &sp = pairs[h->big[a]];
            memcpy(seq.data() + meta[a].toff, ref + sp.idr, (size_t)sp.len1);
            memcpy(seq.data() + meta[a].qoff, qer + sp.idq, (size_t)sp.len2);
        }
        std::vector<BigOut> out(nb);
        BigMeta *d_meta = nullptr; uint8_t *d_seq = nullptr; int32_t *d_rows = nullptr; BigOut *d_out = nullptr;
        cudaSetDevice(h->devs[0].id);
        cudaError_t e = cudaMalloc((void **)&d_meta, sizeof(BigMeta) * nb);
        if (e == cudaSuccess) e = cudaMalloc((void **)&d_seq, seq.size());
        if (e == cudaSuccess) e = cudaMalloc((void **)&d_rows, sizeof(int32_t) * (size_t)rows + 16);
        if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, sizeof(BigOut) * nb);
        if (e == cudaSuccess) e = cudaMemcpy(d_meta, meta.data(), sizeof(BigMeta) * nb, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_seq, seq.data(), seq.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            bsw_big_kernel<<<(int)((nb + 63) / 64), 64>>>(d_meta, (int)nb, d_seq, d_rows, d_out, h->K);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpy(out.data(), d_out, sizeof(BigOut) * nb, cudaMemcpyDeviceToHost);
        cudaFree(d_meta); cudaFree(d_seq); cudaFree(d_rows); cudaFree(d_out);
        if (e != cudaSuccess) return cuda_rc(h, e, "scalar-class pass");
        for (size_t a = 0; a < nb; ++a) {
            bsw_seqpair &sp = pairs[h->big[a]];
            const BigOut &o = out[a];
            sp.score = o.score; sp.qle = o.qle; sp.tle = o.tle; sp.gtle = o.gtle; sp.gscore = o.gscore; sp.max_off = o.max_off;
        }
        st.kernel_launches++;
        st.pairs_scalar = (int64_t)nb;
    }
    if (!h->invalid.empty()) {
        for (int64_t k : h->invalid) {
            bsw_seqpair &sp = pairs[k];
            sp.score = sp.qle = sp.tle = sp.gtle = sp.gscore = sp.max_off = -1;
        }
        st.pairs_invalid = (int64_t)h->invalid.size();
        st.first_invalid = *std::min_element(h->invalid.begin(), h->invalid.end());
        rc = BSW_ERR_RANGE;
    }
    return rc;
}

template <int W>
static int run_peak(int iters, int blocks, int threads, uint32_t *sink, float *ms) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    dpx_peak_kernel<W><<<blocks, threads>>>(sink, 16, 3u);  // warm-up
    cudaEventRecord(a);
    dpx_peak_kernel<W><<<blocks, threads>>>(sink, iters, 3u);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    cudaEventElapsedTime(ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return e == cudaSuccess ? 0 : 1;
}


static void drop_staged(bsw_handle *h) {
    for (Device &d : h->devs) {
        cudaSetDevice(d.id);
        for (Slab *s : d.staged) { free_slab(*s); delete s; }
        d.staged.clear();
    }
    h->staged_n = -1;
}


// ================================================================================================
extern "C" {

int bsw_gpu_version(void) { return kVersion; }

const char *bsw_gpu_strerror(int code) {
    switch (code) {
        case BSW_OK: return "ok";
        case BSW_ERR_ARG: return "invalid argument";
        case BSW_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case BSW_ERR_CUDA: return "CUDA runtime error";
        case BSW_ERR_NOMEM: return "out of memory";
        case BSW_ERR_RANGE: return "pair outside the int16 kernel's valid domain";
        case BSW_ERR_STATE: return "staged API called out of order";
        default: return "unknown error";
    }
}

const char *bsw_gpu_last_error(const bsw_handle *h) { return h ? h->err.c_str() : ""; }

int bsw_gpu_init_devices(const bsw_params *params, int n_devices, const int *device_ids,
                         bsw_handle **out) {
    if (!params || !out || n_devices < 0) return BSW_ERR_ARG;
    *out = nullptr;
    const bsw_params &p = *params;
    if (p.e_del <= 0 || p.e_ins <= 0 || p.o_del < 0 || p.o_ins < 0 || p.match <= 0 || p.match > 127 ||
        p.mismatch < 0 || p.mismatch > 128 || p.ambig < -128 || p.ambig > 127 || p.zdrop < 0 ||
        p.zdrop > 32767 || p.o_del + p.e_del > 16383 || p.o_ins + p.e_ins > 16383 || p.e_del > 8191 ||
        p.e_ins > 8191)
        return BSW_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return BSW_ERR_NO_DEVICE;
    std::vector<int> ids;
    if (device_ids && n_devices > 0) {
        for (int i = 0; i < n_devices; ++i) {
            if (device_ids[i] < 0 || device_ids[i] >= ndev) return BSW_ERR_ARG;
            ids.push_back(device_ids[i]);
        }
    } else {
        int want = n_devices > 0 ? n_devices : ndev;
        if (want > ndev) return BSW_ERR_NO_DEVICE;
        for (int i = 0; i < want; ++i) ids.push_back(i);
    }
    bsw_handle *h = new (std::nothrow) bsw_handle();
    if (!h) return BSW_ERR_NOMEM;
    h->P = p;
    h->K = KParams{p.o_del, p.e_del, p.o_ins, p.e_ins, p.zdrop, p.end_bonus, p.match, p.mismatch, p.ambig, 0,
                   max_score_of(p.match, p.mismatch, p.ambig), 65536u, (uint32_t)(p.match + 1), 1u};
    h->sym = (p.o_del == p.o_ins && p.e_del == p.e_ins);
    memset(&h->stats, 0, sizeof h->stats);
    h->stats.n_gpus = (int)ids.size();
    h->devs.resize(ids.size());
    for (size_t d = 0; d < ids.size(); ++d) {
        h->devs[d].id = ids[d];
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ids[d]) != cudaSuccess || prop.major < 10) {
            delete h;
            return BSW_ERR_NO_DEVICE;  // kernels are sm_100a only
        }
    }
    *out = h;
    return BSW_OK;
}

int bsw_gpu_init(const bsw_params *params, int n_gpus, bsw_handle **out) {
    return bsw_gpu_init_devices(params, n_gpus > 0 ? n_gpus : 0, nullptr, out);
}

void bsw_gpu_free(bsw_handle *h) {
    if (!h) return;
    for (Device &d : h->devs) {
        cudaSetDevice(d.id);
        for (Slab &s : d.ring) free_slab(s);
        for (Slab *s : d.staged) { free_slab(*s); delete s; }
        for (int j = 0; j < kAux; ++j) {
            if (d.aux[j]) cudaStreamDestroy(d.aux[j]);
            if (d.aux_ev[j]) cudaEventDestroy(d.aux_ev[j]);
        }
        if (d.fork_ev) cudaEventDestroy(d.fork_ev);
    }
    delete h;
}

// == the a-priori classification of bwa-mem2 (bwamem.cpp:2218-2228; the three groups sortPairsLenExt forms, :1846-1925)
int bsw_gpu_classify(const bsw_seqpair *pairs, int64_t n, int32_t match, int64_t counts[3], uint8_t *cls) {
    if (n < 0 || (n > 0 && !pairs) || !counts || match <= 0) return BSW_ERR_ARG;
    int64_t c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel for schedule(static) reduction(+ : c0) reduction(+ : c1) reduction(+ : c2)
    for (int64_t k = 0; k < n; ++k) {
        const bsw_seqpair &sp = pairs[k];
        const int64_t minval = (int64_t)sp.h0 + (int64_t)std::min(sp.len1, sp.len2) * match;
        int c;
        if (sp.len1 < 128 && sp.len2 < 128 && minval < 128) { c = 0; ++c0; }                  // MAX_SEQ_LEN8
        else if (sp.len1 < 32768 && sp.len2 < 32768 && minval < 32768) { c = 1; ++c1; }        // MAX_SEQ_LEN16
        else { c = 2; ++c2; }
        if (cls) cls[k] = (uint8_t)c;
    }
    counts[0] = c0; counts[1] = c1; counts[2] = c2;
    return BSW_OK;
}

int bsw_gpu_get_stats(const bsw_handle *h, bsw_gpu_stats *out) {
    if (!h || !out) return BSW_ERR_ARG;
    *out = h->stats;
    return BSW_OK;
}

int bsw_gpu_reserve(bsw_handle *h, int64_t n_pairs, int64_t total_bases) {
    if (!h || n_pairs < 0 || total_bases < 0) return BSW_ERR_ARG;
    if (n_pairs == 0) return BSW_OK;
    // one ring slot holds a slab: at most slab_pairs(false) pairs and their share of the bases
    const int64_t slab = std::min<int64_t>(n_pairs, slab_pairs(false));
    const double share = (double)slab / (double)n_pairs;
    const size_t blob = (size_t)(((double)total_bases * share / 4 + 20.0 * (double)slab) * 1.15) + 65536 +
                        (size_t)omp_get_max_threads() * kArenaWords * 4;
    int64_t nslabs = 0;   // as cut_slabs will cut them (by pair count; its base-count limit only adds slabs)
    for (int64_t done = 0; done < n_pairs; ++nslabs) {
        const int64_t t = slab_target(n_pairs - done, slab_pairs(false));
        done += (std::min(n_pairs - done, t) + kCutGroup - 1) / kCutGroup * kCutGroup;
    }
    for (size_t d = 0; d < h->devs.size(); ++d) {
        Device &dev = h->devs[d];
        CU(cudaSetDevice(dev.id));
        int rc = ensure_aux(h, dev);
        if (rc) return rc;
        const int64_t mine = (nslabs + (int64_t)h->devs.size() - 1 - (int64_t)d) / (int64_t)h->devs.size();
        for (int r = 0; r < kRing && r < mine; ++r)
            if ((rc = ensure_slab(h, dev.ring[r], slab, blob))) return rc;
    }
    return BSW_OK;
}

int bsw_gpu_batch(bsw_handle *h, bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                  int64_t n, int32_t w) {
    if (!h || n < 0 || (n > 0 && (!pairs || !ref || !qer)) || w < 0) return BSW_ERR_ARG;
    NvtxRange nvtx_call("bsw_gpu_batch");
    auto t_all = Clock::now();
    const int ng = (int)h->devs.size();
    bsw_gpu_stats &st = h->stats;
    st.pairs = n; st.kernel_launches = 0; st.h2d_bytes = 0; st.d2h_bytes = 0;
    st.pairs_short = 0; st.pairs_long = 0; st.pairs_keyed = 0; st.pairs_duo = 0;
    st.host_bin_ms = st.host_pack_ms = st.host_scatter_ms = st.kernel_ms = 0;
    st.host_sort_ms = st.host_plan_ms = st.host_alloc_ms = st.host_cut_ms = st.host_wait_ms = 0;
    h->K.w = w;
    h->big.clear(); h->invalid.clear();
    st.pairs_scalar = st.pairs_invalid = 0; st.first_invalid = -1;
    if (n == 0) { st.wall_ms = 0; return BSW_OK; }

    std::vector<int64_t> cuts, cuts_flat;
    std::vector<size_t> guess, guess_flat;
    {
        auto t0 = Clock::now();
        cut_slabs(pairs, n, cuts, guess, false, &cuts_flat, &guess_flat);
        st.host_cut_ms = ms_since(t0);
    }
    // the tapered and the untapered plan share their leading full slabs; where they part, the call keeps
    // the taper only if the host has been the slower side so far (it hardly ever waited for a ring slot)
    size_t common = 0;
    while (common + 1 < cuts.size() && common + 1 < cuts_flat.size() && cuts[common + 1] == cuts_flat[common + 1]) ++common;
    int nslabs = (int)cuts.size() - 1;
    std::vector<double> kms((size_t)ng, 0.0);
    int rc = BSW_OK;

    for (int sidx = 0; sidx < nslabs && rc == BSW_OK; ++sidx) {
        if ((size_t)sidx == common && common > 0 && cuts != cuts_flat &&
            st.host_wait_ms > 0.2 * ms_since(t_all)) {
            cuts = cuts_flat;
            guess = guess_flat;
            nslabs = (int)cuts.size() - 1;
        }
        const int d = sidx % ng, r = (sidx / ng) % kRing;
        Device &dev = h->devs[(size_t)d];
        Slab &s = dev.ring[r];
        if ((rc = cuda_rc(h, cudaSetDevice(dev.id), "cudaSetDevice"))) break;
        // results that are ready go into the caller's array while this slab is packed: the ring slot's
        // own older slab (wait for it if need be) and any other slab of this GPU that has finished
        ScatterJobs jobs;
        rc = claim_slab(h, s, pairs, &kms[(size_t)d], jobs);
        if (rc) break;
        for (int r2 = 0; r2 < kRing && rc == BSW_OK; ++r2) {
            Slab &o = dev.ring[r2];
            if (&o != &s && o.busy && cudaEventQuery(o.ev_done) == cudaSuccess)
                rc = claim_slab(h, o, pairs, &kms[(size_t)d], jobs);
        }
        if (rc) break;
        {
            NvtxRange nvtx_pack("slab: validate + 2-bit pack + scatter of older results");
            rc = prepare_slab_fit(h, s, pairs, ref, qer, cuts[(size_t)sidx], (int)(cuts[(size_t)sidx + 1] - cuts[(size_t)sidx]),
                                  guess[(size_t)sidx], jobs);
        }
        if (rc) break;
        NvtxRange nvtx_enq("slab: enqueue H2D, binning, DP launches, D2H");
        if (s.n_dev) {
            // (an error after the first enqueue leaves work in flight on the slab's stream: the slab is marked
            // busy all the same, so that the drain below waits for it before anything is freed or reused)
            s.busy = true;
            if ((rc = upload_slab(h, s))) break;
            if ((rc = cuda_rc(h, cudaEventRecord(s.ev_k0, s.stream), "cudaEventRecord"))) break;
            if ((rc = bin_slab(h, s, s.stream))) break;
            if ((rc = launch_slab(h, dev, s))) break;
            if ((rc = cuda_rc(h, cudaEventRecord(s.ev_k1, s.stream), "cudaEventRecord"))) break;
            if ((rc = download_slab(h, s))) break;
        }
        s.busy = true;
        if ((rc = cuda_rc(h, cudaEventRecord(s.ev_done, s.stream), "cudaEventRecord"))) break;
    }
    // drain (also on error, so that no stream still writes pinned memory we may free later)
    for (int d = 0; d < ng; ++d) {
        Device &dev = h->devs[(size_t)d];
        cudaSetDevice(dev.id);
        for (int r = 0; r < kRing; ++r) {
            Slab &s = dev.ring[r];
            if (!s.busy) continue;
            if (rc == BSW_OK) rc = finish_slab(h, s, pairs, &kms[(size_t)d]);
            else { cudaStreamSynchronize(s.stream); s.busy = false; }
            if (rc != BSW_OK && s.busy) { cudaStreamSynchronize(s.stream); s.busy = false; }
        }
    }
    st.kernel_ms = *std::max_element(kms.begin(), kms.end());
    if (rc == BSW_OK) rc = settle_classes(h, pairs, ref, qer);
    st.wall_ms = ms_since(t_all);
    return rc;
}

// ---- the production caller's band-doubling retry (bwa-mem2 bwamem.cpp:2448-2508) --------------------

int bsw_gpu_batch_retry(bsw_handle *h, bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                        int64_t n, int32_t w, int32_t max_tries, int32_t *tries) {
    if (!h || n < 0 || (n > 0 && (!pairs || !ref || !qer)) || w < 0 || max_tries < 1 || max_tries > 16 ||
        ((int64_t)w << (max_tries - 1)) > 0x3FFFFFFF)
        return BSW_ERR_ARG;
    // try 0 over everything, in place
    int rc = bsw_gpu_batch(h, pairs, ref, qer, n, w);
    if (rc) return rc;
    bsw_gpu_stats total = h->stats;
    if (tries) for (int64_t k = 0; k < n; ++k) tries[k] = 1;
    std::vector<int64_t> act;          // caller indices still being retried
    std::vector<bsw_seqpair> aux;      // their records, compacted (bwamem.cpp: pair_ar_aux)
    std::vector<int32_t> prev;
    for (int t = 0; t + 1 < max_tries; ++t) {
        const int32_t wt = w << t;
        const int32_t lim = (wt >> 1) + (wt >> 2);
        // a pair is final if its score did not change or its alignment stayed within 3/4 of the band
        // (bwamem.cpp:2479-2480; the score before the first try is -1, as in bwa's mem_chain2aln)
        if (t == 0) {
            for (int64_t k = 0; k < n; ++k)
                if (pairs[k].max_off >= lim) act.push_back(k);
        } else {
            std::vector<int64_t> next;
            std::vector<int32_t> nprev;
            for (size_t a = 0; a < act.size(); ++a) {
                const bsw_seqpair &sp = aux[a];
                if (!(sp.score == prev[a] || sp.max_off < lim)) next.push_back(act[a]);
            }
            act.swap(next);
        }
        if (act.empty()) break;
        aux.resize(act.size());
        prev.resize(act.size());
        for (size_t a = 0; a < act.size(); ++a) { aux[a] = pairs[act[a]]; prev[a] = aux[a].score; }
        rc = bsw_gpu_batch(h, aux.data(), ref, qer, (int64_t)aux.size(), w << (t + 1));
        if (rc) return rc;
        total.kernel_launches += h->stats.kernel_launches;
        total.h2d_bytes += h->stats.h2d_bytes; total.d2h_bytes += h->stats.d2h_bytes;
        total.kernel_ms += h->stats.kernel_ms; total.wall_ms += h->stats.wall_ms;
        for (size_t a = 0; a < act.size(); ++a) {
            bsw_seqpair &dst = pairs[act[a]];
            const bsw_seqpair &sp = aux[a];
            dst.score = sp.score; dst.tle = sp.tle; dst.gtle = sp.gtle; dst.qle = sp.qle;
            dst.gscore = sp.gscore; dst.max_off = sp.max_off;
            if (tries) tries[act[a]] = t + 2;
        }
    }
    total.pairs = n;
    h->stats = total;
    return BSW_OK;
}

// ---- packed input (the pair-file ingest path) -----------------------------------------
// The caller hands over what a BSWPAIR1 file holds (include/bsw_pairio.h) -- 12-byte records and the sequences at
// 2 bits per base (4 for a pair holding an ambiguous base), query then target, each padded to 4 bytes -- and gets
// 16-byte result records back. The packed data IS the device blob: the host only turns the records into PairMeta
// (a prefix sum of the pairs' sizes, the launch histogram) and starts the copies. When `data` / `out` are
// page-locked (bsw_gpu_host_alloc, or registered by the caller) they are DMA'd in place; otherwise they go
// through the slab's pinned staging buffers.

namespace {

inline bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

inline uint32_t rec_words(const bsw_packed_rec &r) {
    const bool wide = r.flags & 1u;
    return (seq_bytes(r.len2, wide) + seq_bytes(r.len1, wide)) >> 2;
}

struct PackedSlabOut {       // where a slab's results go once its stream is done
    bsw_result *dst = nullptr;   // null: the device wrote straight into the caller's (pinned) array
    int n = 0;
};

}  // namespace

void *bsw_gpu_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void bsw_gpu_host_free(void *p) { if (p) cudaFreeHost(p); }

int bsw_gpu_batch_packed(bsw_handle *h, const bsw_packed_rec *rec, const uint8_t *data, int64_t data_bytes,
                         int64_t n, int32_t w, bsw_result *out) {
    static_assert(sizeof(bsw_packed_rec) == 12 && sizeof(bsw_result) == sizeof(PairOut), "packed layouts");
    if (!h || n < 0 || w < 0 || data_bytes < 0 || (n > 0 && (!rec || !out || (!data && data_bytes > 0)))) return BSW_ERR_ARG;
    if (((uintptr_t)data & 3u) != 0) return BSW_ERR_ARG;
    NvtxRange nvtx_call("bsw_gpu_batch_packed");
    auto t_all = Clock::now();
    const int ng = (int)h->devs.size();
    bsw_gpu_stats &st = h->stats;
    st.pairs = n; st.kernel_launches = 0; st.h2d_bytes = 0; st.d2h_bytes = 0;
    st.pairs_short = 0; st.pairs_long = 0; st.pairs_keyed = 0; st.pairs_duo = 0;
    st.host_bin_ms = st.host_pack_ms = st.host_scatter_ms = st.kernel_ms = 0;
    st.host_sort_ms = st.host_plan_ms = st.host_alloc_ms = st.host_cut_ms = st.host_wait_ms = 0;
    h->K.w = w;
    if (n == 0) { st.wall_ms = 0; return BSW_OK; }
    const int T = omp_get_max_threads();
    const int match = h->P.match;

    // Slabs are cut, sized and validated one at a time (whole chunks of 4096 pairs, at most slab_pairs(false) pairs and
    // kSlabBases / 4 packed bytes), so the first slab is on its way before the rest of the records has been looked at.
    constexpr int64_t kChunk = 4096;
    const int64_t full = std::max<int64_t>(slab_pairs(false) / kChunk * kChunk, kChunk);
    std::vector<uint64_t> cw((size_t)(full / kChunk) + 1, 0);
    auto t0 = Clock::now();
    const bool data_pinned = is_pinned(data), out_pinned = is_pinned(out), rec_pinned = is_pinned(rec);
    std::vector<double> kms((size_t)ng, 0.0);
    std::vector<PackedSlabOut> pending((size_t)ng * kRing);
    int rc = BSW_OK;

    // waits for a ring slot's previous slab and delivers its results
    auto finish = [&](int d, int r) -> int {
        Slab &s = h->devs[(size_t)d].ring[r];
        if (!s.busy) return BSW_OK;
        auto tw = Clock::now();
        const int e = cuda_rc(h, cudaEventSynchronize(s.ev_done), "cudaEventSynchronize");
        st.host_wait_ms += ms_since(tw);
        s.busy = false;
        if (e) return e;
        if (s.n_dev) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1) == cudaSuccess) kms[(size_t)d] += ms;
        }
        PackedSlabOut &po = pending[(size_t)d * kRing + r];
        if (po.dst) {
            auto ts = Clock::now();
            const char *src = reinterpret_cast<const char *>(s.h_out);
            char *dst = reinterpret_cast<char *>(po.dst);
            const size_t bytes = sizeof(PairOut) * (size_t)po.n;
#pragma omp parallel for schedule(static)
            for (int t = 0; t < T; ++t) {
                const size_t a = bytes * (size_t)t / T & ~(size_t)63, b = t + 1 == T ? bytes : (bytes * (size_t)(t + 1) / T & ~(size_t)63);
                memcpy(dst + a, src + a, b - a);
            }
            st.host_scatter_ms += ms_since(ts);
            po.dst = nullptr;
        }
        return BSW_OK;
    };

    // Two steps per slab, pipelined one slab apart: PREP (records -> device, bsw_rec_meta_kernel, statistics back) is
    // issued for slab s + 1 before the host waits for the statistics of slab s, plans its launches and enqueues them.
    struct PSlab { int d = 0, r = 0, ns = 0; int64_t lo = 0, nch = 0; bool valid = false; };
    int prep_idx = 0;          // next slab to prepare
    int64_t prep_lo = 0;       // its first pair
    auto prep = [&](PSlab &ps) -> int {
        ps.valid = false;
        if (prep_lo >= n) return BSW_OK;
        const int sidx = prep_idx;
        const int d = sidx % ng, r = (sidx / ng) % kRing;
        Device &dev = h->devs[(size_t)d];
        Slab &s = dev.ring[r];
        int e = cuda_rc(h, cudaSetDevice(dev.id), "cudaSetDevice");
        if (e) return e;
        if ((e = finish(d, r))) return e;
        // slab sizes ramp up from 128 Ki pairs (the GPU starts after a fraction of a millisecond of copying instead of
        // a whole slab's) and taper off at the end (the last download and hand-over are short)
        int64_t target = std::min<int64_t>(full, (int64_t)131072 << std::min(sidx / ng, 6));
        if (use_taper() && n - prep_lo <= target + target / 2 && n - prep_lo > 2 * 131072)
            target = std::max<int64_t>(131072, ((n - prep_lo) / 2 + kChunk - 1) / kChunk * kChunk);
        const int64_t nch = (std::min(n - prep_lo, target) + kChunk - 1) / kChunk;
        const int ns = (int)(std::min(n, prep_lo + nch * kChunk) - prep_lo);
        if (nch > kBaseSlots) return BSW_ERR_RANGE;
        auto t1 = Clock::now();
        e = ensure_slab(h, s, ns, 0);
        st.host_alloc_ms += ms_since(t1);
        if (e) return e;
        // ---- records -> PairMeta ON THE DEVICE (bsw_rec_meta_kernel: sizes, chunk-relative offsets, validation, launch
        // histogram, maxima; bsw_chunk_scan_kernel: chunk bases); the host reads the 16 KB statistics block only
        t1 = Clock::now();
        s.lo = prep_lo; s.n = ns; s.trivial.clear();
        s.busy = true;      // (from here on work may be in flight on the slab's stream: see bsw_gpu_batch)
        const bsw_packed_rec *rsrc = rec + prep_lo;
        if (!rec_pinned) {   // 12 bytes per pair through the (otherwise unused) pinned meta buffer
            char *dst = reinterpret_cast<char *>(s.h_meta);
            const char *src = reinterpret_cast<const char *>(rsrc);
            const size_t bytes = sizeof(bsw_packed_rec) * (size_t)ns;
#pragma omp parallel for schedule(static)
            for (int t = 0; t < T; ++t) {
                const size_t a = bytes * (size_t)t / T & ~(size_t)63, b = t + 1 == T ? bytes : (bytes * (size_t)(t + 1) / T & ~(size_t)63);
                memcpy(dst + a, src + a, b - a);
            }
            rsrc = reinterpret_cast<const bsw_packed_rec *>(s.h_meta);
        }
        PackedRecDev *d_rec = reinterpret_cast<PackedRecDev *>(s.d_keys);   // 16 bytes per pair there; free until the key kernel
        if ((e = cuda_rc(h, cudaMemcpyAsync(d_rec, rsrc, sizeof(bsw_packed_rec) * (size_t)ns, cudaMemcpyHostToDevice, s.stream), "H2D records"))) return e;
        if ((e = cuda_rc(h, cudaMemsetAsync(s.d_stats, 0, sizeof(SlabStatsDev), s.stream), "memset"))) return e;
        bsw_rec_meta_kernel<<<(int)nch, 256, 0, s.stream>>>(d_rec, ns, match, s.d_meta, s.d_chunk_total, s.d_stats);
        bsw_chunk_scan_kernel<<<1, 256, 0, s.stream>>>(s.d_chunk_total, (int)nch, s.d_base, s.d_stats);
        if ((e = cuda_rc(h, cudaGetLastError(), "record kernels"))) return e;
        st.kernel_launches += 2;
        if ((e = cuda_rc(h, cudaMemcpyAsync(s.h_stats, s.d_stats, sizeof(SlabStatsDev), cudaMemcpyDeviceToHost, s.stream), "D2H stats"))) return e;
        if ((e = cuda_rc(h, cudaEventRecord(s.ev_stats, s.stream), "cudaEventRecord"))) return e;
        st.host_plan_ms += ms_since(t1);
        ps.d = d; ps.r = r; ps.ns = ns; ps.lo = prep_lo; ps.nch = nch; ps.valid = true;
        prep_lo += ns;
        ++prep_idx;
        return BSW_OK;
    };

    uint64_t w_lo = 0;       // first word in `data` of the slab being launched
    auto launch = [&](const PSlab &ps) -> int {
        const int d = ps.d, r = ps.r, ns = ps.ns;
        const int64_t lo = ps.lo;
        Device &dev = h->devs[(size_t)d];
        Slab &s = dev.ring[r];
        int rc = cuda_rc(h, cudaSetDevice(dev.id), "cudaSetDevice");
        if (rc) return rc;
        {
            auto tw = Clock::now();
            rc = cuda_rc(h, cudaEventSynchronize(s.ev_stats), "cudaEventSynchronize");
            st.host_wait_ms += ms_since(tw);
            if (rc) return rc;
        }
        t0 = Clock::now();
        const SlabStatsDev &sd = *s.h_stats;
        if (sd.bad) return BSW_ERR_RANGE;
        const uint64_t words = sd.total_words;
        if ((w_lo + words) * 4 > (uint64_t)data_bytes) return BSW_ERR_ARG;
        if (words > 0xFFFFFF00ull) return BSW_ERR_RANGE;
        h->hist.assign((size_t)2 * kMaxBins, 0);
        static_assert(kRecBins == kMaxBins, "device and host launch bins");
        memcpy(h->hist.data(), &sd.hist[0][0], sizeof(uint32_t) * 2 * kMaxBins);
        s.use_base = true;
        s.fastm = (int64_t)sd.maxsc * (match + 1) <= 32767;
        s.max_sc = sd.maxsc;
        s.key_b1 = bits_for((uint32_t)sd.maxt);
        s.key_b0 = bits_for((uint32_t)sd.maxh);
        s.n_dev = ns - (int)sd.ntriv;
        s.blob_bytes = (size_t)words * 4;
        plan_slab(h, s, 1, sd.maxq);
        st.host_plan_ms += ms_since(t0);
        t0 = Clock::now();
        if (!data_pinned) rc = ensure_slab(h, s, ns, (size_t)words * 4 + 64);
        if (!rc) rc = ensure_dblob(h, s, (size_t)words * 4 + 64);
        st.host_alloc_ms += ms_since(t0);
        if (rc) return rc;
        // ---- the packed sequences: in place if page-locked, else through the slab's pinned blob
        const uint8_t *src = data + (size_t)w_lo * 4;
        if (!data_pinned && words) {
            t0 = Clock::now();
            char *dst = reinterpret_cast<char *>(s.h_blob);
            const size_t bytes = (size_t)words * 4;
#pragma omp parallel for schedule(static)
            for (int t = 0; t < T; ++t) {
                const size_t a = bytes * (size_t)t / T & ~(size_t)63, b = t + 1 == T ? bytes : (bytes * (size_t)(t + 1) / T & ~(size_t)63);
                memcpy(dst + a, src + a, b - a);
            }
            src = reinterpret_cast<const uint8_t *>(s.h_blob);
            st.host_pack_ms += ms_since(t0);
        }
        if (words && (rc = cuda_rc(h, cudaMemcpyAsync(s.d_blob, src, (size_t)words * 4, cudaMemcpyHostToDevice, s.stream), "H2D blob"))) return rc;
        // (the kernels read one word past a pair's target a refill period ahead)
        if ((rc = cuda_rc(h, cudaMemsetAsync(reinterpret_cast<char *>(s.d_blob) + (size_t)words * 4, 0, 16, s.stream), "memset"))) return rc;
        st.h2d_bytes += (int64_t)(sizeof(bsw_packed_rec) * (size_t)ns + (size_t)words * 4);
        if ((rc = cuda_rc(h, cudaEventRecord(s.ev_k0, s.stream), "cudaEventRecord"))) return rc;
        if (s.n_dev) {
            if ((rc = bin_slab(h, s, s.stream))) return rc;
            if ((rc = launch_slab(h, dev, s))) return rc;
        } else if (ns) {   // only pairs with an empty sequence: the key kernel answers them
            bsw_key_kernel<<<(ns + 255) / 256, 256, 0, s.stream>>>(s.d_meta, ns, s.d_keys, s.d_ord, s.key_b1, s.key_b0, s.long_bin0, s.d_out, nullptr);
        }
        if ((rc = cuda_rc(h, cudaEventRecord(s.ev_k1, s.stream), "cudaEventRecord"))) return rc;
        PackedSlabOut &po = pending[(size_t)d * kRing + r];
        po.n = ns;
        po.dst = out_pinned ? nullptr : out + lo;
        if ((rc = cuda_rc(h, cudaMemcpyAsync(out_pinned ? reinterpret_cast<void *>(out + lo) : reinterpret_cast<void *>(s.h_out), s.d_out,
                                             sizeof(PairOut) * (size_t)ns, cudaMemcpyDeviceToHost, s.stream), "D2H out"))) return rc;
        st.d2h_bytes += (int64_t)(sizeof(PairOut) * (size_t)ns);
        if ((rc = cuda_rc(h, cudaEventRecord(s.ev_done, s.stream), "cudaEventRecord"))) return rc;
        w_lo += words;
        return BSW_OK;
    };

    {
        PSlab cur, nxt;
        rc = prep(cur);
        while (rc == BSW_OK && cur.valid) {
            if ((rc = prep(nxt))) break;      // the next slab's records are on their way while this one is planned
            if ((rc = launch(cur))) break;
            cur = nxt;
        }
    }
    // drain (also on error, so that no stream still touches memory we or the caller may free)
    for (int d = 0; d < ng; ++d) {
        cudaSetDevice(h->devs[(size_t)d].id);
        for (int r = 0; r < kRing; ++r) {
            Slab &s = h->devs[(size_t)d].ring[r];
            if (!s.busy) continue;
            if (rc == BSW_OK) rc = finish(d, r);
            if (s.busy) { cudaStreamSynchronize(s.stream); s.busy = false; }
        }
    }
    st.kernel_ms = *std::max_element(kms.begin(), kms.end());
    st.wall_ms = ms_since(t_all);
    return rc;
}

// ---- staged API --------------------------------------------------------------------------------

int bsw_gpu_stage(bsw_handle *h, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                  int64_t n, int32_t w) {
    if (!h || n < 0 || (n > 0 && (!pairs || !ref || !qer)) || w < 0) return BSW_ERR_ARG;
    drop_staged(h);
    h->big.clear(); h->invalid.clear();
    bsw_gpu_stats &st = h->stats;
    st.pairs = n; st.h2d_bytes = 0; st.d2h_bytes = 0; st.kernel_launches = 0;
    st.pairs_short = st.pairs_long = st.pairs_keyed = st.pairs_duo = 0;
    h->K.w = w;
    const int ng = (int)h->devs.size();
    std::vector<int64_t> cuts;
    std::vector<size_t> guess;
    cut_slabs(pairs, n, cuts, guess, true, nullptr, nullptr, ng);
    const int nslabs = (int)cuts.size() - 1;
    for (int sidx = 0; sidx < nslabs; ++sidx) {
        Device &dev = h->devs[(size_t)(sidx % ng)];
        CU(cudaSetDevice(dev.id));
        Slab *s = new (std::nothrow) Slab();
        if (!s) return BSW_ERR_NOMEM;
        dev.staged.push_back(s);
        int rc = prepare_slab_fit(h, *s, pairs, ref, qer, cuts[(size_t)sidx], (int)(cuts[(size_t)sidx + 1] - cuts[(size_t)sidx]),
                                  guess[(size_t)sidx]);
        if (rc) { drop_staged(h); return rc; }
        if ((rc = upload_slab(h, *s))) { drop_staged(h); return rc; }
        CU(cudaStreamSynchronize(s->stream));
    }
    if (!h->big.empty() || !h->invalid.empty()) {   // the resident (measurement) path is the int16 kernels only
        drop_staged(h);
        return BSW_ERR_RANGE;
    }
    h->staged_n = n; h->staged_w = w;
    return BSW_OK;
}

int bsw_gpu_run_staged(bsw_handle *h, float *kernel_ms) {
    if (!h) return BSW_ERR_ARG;
    if (h->staged_n < 0) return BSW_ERR_STATE;
    NvtxRange nvtx_call("bsw_gpu_run_staged");
    h->K.w = h->staged_w;
    h->stats.kernel_launches = 0;
    h->stats.pairs_short = h->stats.pairs_long = h->stats.pairs_keyed = h->stats.pairs_duo = 0;
    // all slabs of one GPU run back to back on that GPU's first stream; GPUs run concurrently
    for (Device &dev : h->devs) {
        if (dev.staged.empty()) continue;
        CU(cudaSetDevice(dev.id));
        cudaStream_t st0 = dev.staged[0]->stream;
        CU(cudaEventRecord(dev.staged[0]->ev_k0, st0));
        std::vector<Slab *> live;
        for (Slab *s : dev.staged) if (s->n_dev) live.push_back(s);
        if (!live.empty()) {
            int rc = BSW_OK;
            for (Slab *s : live) if ((rc = bin_slab(h, *s, st0))) return rc;   // binning is part of the path
            rc = launch_slabs(h, dev, st0, live.data(), (int)live.size());
            if (rc) return rc;
        }
        CU(cudaEventRecord(dev.staged[0]->ev_k1, st0));
    }
    float worst = 0.f;
    for (Device &dev : h->devs) {
        if (dev.staged.empty()) continue;
        CU(cudaSetDevice(dev.id));
        CU(cudaEventSynchronize(dev.staged[0]->ev_k1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, dev.staged[0]->ev_k0, dev.staged[0]->ev_k1));
        worst = std::max(worst, ms);
    }
    h->stats.kernel_ms = worst;
    if (kernel_ms) *kernel_ms = worst;
    return BSW_OK;
}

int bsw_gpu_count_staged(bsw_handle *h, int64_t *cells_visited) {
    if (!h || !cells_visited) return BSW_ERR_ARG;
    if (h->staged_n < 0) return BSW_ERR_STATE;
    h->K.w = h->staged_w;
    int64_t total = 0;
    for (Device &dev : h->devs) {
        CU(cudaSetDevice(dev.id));
        for (Slab *s : dev.staged) {
            if (!s->n_dev) continue;
            int rc = bin_slab(h, *s, s->stream);
            if (rc) return rc;
            if ((rc = launch_slab(h, dev, *s, true))) return rc;
            if ((rc = download_slab(h, *s))) return rc;
            CU(cudaStreamSynchronize(s->stream));
            int64_t sum = 0;
            const int n = s->n;
            const PairOut *o = s->h_out;
            // entries of pairs answered on the host were never written by the device
            std::vector<char> skip((size_t)n, 0);
            for (uint32_t k : s->trivial) skip[k] = 1;
#pragma omp parallel for reduction(+ : sum) schedule(static)
            for (int k = 0; k < n; ++k) if (!skip[(size_t)k]) sum += o[k].cells;
            total += sum;
        }
    }
    *cells_visited = total;
    return BSW_OK;
}

int bsw_gpu_fetch_staged(bsw_handle *h, bsw_seqpair *pairs, int64_t n) {
    if (!h || !pairs) return BSW_ERR_ARG;
    if (h->staged_n < 0 || n != h->staged_n) return BSW_ERR_STATE;
    h->stats.d2h_bytes = 0;
    for (Device &dev : h->devs) {
        CU(cudaSetDevice(dev.id));
        for (Slab *s : dev.staged) {
            if (s->n_dev) {
                int rc = download_slab(h, *s);
                if (rc) return rc;
                CU(cudaStreamSynchronize(s->stream));
            }
            scatter_slab(h, *s, s->h_out, pairs);
        }
    }
    return BSW_OK;
}

// ---- integer-pipe microbenchmark -----------------------------------------------------------------

int bsw_gpu_dpx_peak(int device, int which, double *ginstr_per_s, double *sm_mhz_est) {
    if (!ginstr_per_s) return BSW_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BSW_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BSW_ERR_NO_DEVICE;
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
    uint32_t *sink = nullptr;
    if (cudaMalloc((void **)&sink, 4096) != cudaSuccess) return BSW_ERR_NOMEM;
    float ms = 0.f;
    int rc = 0;
    switch (which) {
        case 0: rc = run_peak<0>(iters, blocks, threads, sink, &ms); break;
        case 1: rc = run_peak<1>(iters, blocks, threads, sink, &ms); break;
        case 2: rc = run_peak<2>(iters, blocks, threads, sink, &ms); break;
        case 3: rc = run_peak<3>(iters, blocks, threads, sink, &ms); break;
        case 4: rc = run_peak<4>(iters, blocks, threads, sink, &ms); break;
        case 5: rc = run_peak<5>(iters, blocks, threads, sink, &ms); break;
        case 6: rc = run_peak<6>(iters, blocks, threads, sink, &ms); break;
        case 7: rc = run_peak<7>(iters, blocks, threads, sink, &ms); break;
        case 8: rc = run_peak<8>(iters, blocks, threads, sink, &ms); break;
        case 9: {   // one inner-loop trip on registers: result in giga CELLS per second
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            bsw_trip_peak_kernel<<<blocks, threads>>>(sink, 16, 3u, 65536u, 2u, 1u, 128u);
            cudaEventRecord(a);
            bsw_trip_peak_kernel<<<blocks, threads>>>(sink, iters * 4, 3u, 65536u, 2u, 1u, 128u);
            cudaEventRecord(b);
            rc = cudaEventSynchronize(b) == cudaSuccess ? 0 : 1;
            cudaEventElapsedTime(&ms, a, b);
            cudaEventDestroy(a); cudaEventDestroy(b);
            cudaFree(sink);
            if (rc) return BSW_ERR_CUDA;
            *ginstr_per_s = (double)iters * 4.0 * 8.0 * (double)threads * (double)blocks / (ms * 1e-3) / 1e9;
            if (sm_mhz_est) *sm_mhz_est = (double)prop.clockRate / 1000.0;
            return BSW_OK;
        }
        case 10: case 11: {   // one inner-loop trip of the two-pairs-per-thread kernel on registers: giga CELLS per second
            const int b2 = which == 10 ? blocks : prop.multiProcessorCount * 5, t2 = which == 10 ? threads : 32;
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            bsw_duo_trip_peak_kernel<<<b2, t2>>>(sink, 16, 3u, 65536u, 2u, 1u, 256u);
            cudaEventRecord(a);
            bsw_duo_trip_peak_kernel<<<b2, t2>>>(sink, iters * 4, 3u, 65536u, 2u, 1u, 256u);
            cudaEventRecord(b);
            rc = cudaEventSynchronize(b) == cudaSuccess ? 0 : 1;
            cudaEventElapsedTime(&ms, a, b);
            cudaEventDestroy(a); cudaEventDestroy(b);
            cudaFree(sink);
            if (rc) return BSW_ERR_CUDA;
            *ginstr_per_s = (double)iters * 4.0 * 8.0 * (double)t2 * (double)b2 / (ms * 1e-3) / 1e9;
            if (sm_mhz_est) *sm_mhz_est = (double)prop.clockRate / 1000.0;
            return BSW_OK;
        }
        default: cudaFree(sink); return BSW_ERR_ARG;
    }
    cudaFree(sink);
    if (rc) return BSW_ERR_CUDA;
    // thread-instructions of the measured kind: iters * 4 (unroll) * 8 (chains) per thread
    double per_thread = (double)iters * 4.0 * 8.0 * (which == 8 ? 2.0 : 1.0);
    double total = per_thread * (double)threads * (double)blocks;
    *ginstr_per_s = total / (ms * 1e-3) / 1e9;
    if (sm_mhz_est) *sm_mhz_est = (double)prop.clockRate / 1000.0;
    return BSW_OK;
}

// Developer probe: the register-only trip loops (kind 0: extend_pair's keyed trip, 1: extend_duo2's) at a given
// number of one-warp blocks per SM -> giga cells per second. Shows how many warps each loop needs.
int bsw_gpu_trip_probe(int device, int kind, int warps_per_sm, double *gcells_per_s) {
    if (!gcells_per_s || warps_per_sm < 1 || warps_per_sm > 32) return BSW_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BSW_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BSW_ERR_NO_DEVICE;
    uint32_t *sink = nullptr;
    if (cudaMalloc((void **)&sink, 4096) != cudaSuccess) return BSW_ERR_NOMEM;
    const int blocks = prop.multiProcessorCount * warps_per_sm, iters = 1 << 16;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms = 0.f;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        if (kind == 0) bsw_trip_peak_kernel<<<blocks, 32>>>(sink, iters, 3u, 65536u, 2u, 1u, 128u);
        else bsw_duo_trip_peak_kernel<<<blocks, 32>>>(sink, iters, 3u, 65536u, 2u, 1u, 256u);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(sink);
    *gcells_per_s = (double)iters * 8.0 * 32.0 * (double)blocks / (ms * 1e-3) / 1e9;
    return BSW_OK;
}

}  // extern "C"
