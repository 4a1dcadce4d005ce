// Host-only: synthetic pair generator and the reference driver's text pair-file
// format (reader mirrors loadPairs, /root/reference/benchmarks/bsw/src/main_banded.cpp:164-206;
// the file format is documented at main_banded.cpp:152-162). Declared in include/bsw_pairio.h.
#include "bsw_pairio.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int64_t kChunk = 4096;  // pairs per independently-seeded generator chunk

struct Rng {
    std::mt19937_64 g;
    explicit Rng(uint64_t s) : g(s) {}
    uint64_t u64() { return g(); }
    // uniform integer in [lo, hi]
    int64_t range(int64_t lo, int64_t hi) { return lo + (int64_t)(u64() % (uint64_t)(hi - lo + 1)); }
    double unit() { return (double)(u64() >> 11) * (1.0 / 9007199254740992.0); }
    uint8_t base() { return (uint8_t)(u64() & 3); }
};

struct ChunkOut {
    std::vector<uint8_t> ref, qer;
};

void gen_one(Rng &r, const bsw_gen_config &c, std::vector<uint8_t> &ref, std::vector<uint8_t> &qer,
             bsw_seqpair &sp) {
    int len2 = 1, h0 = 0, len1 = 1;
    if (c.mode == BSW_GEN_READ_FLANK) {
        // a read of read_len bases with a seed of length s at position p: the extension query is the
        // flank left or right of the seed and the seed score is h0 = s * match (match = 1)
        int s = (int)r.range(c.seed_min, c.seed_max);
        int p = (int)r.range(0, c.read_len - s);
        int left = p, right = c.read_len - s - p;
        len2 = (r.u64() & 1) ? right : left;
        if (len2 == 0) len2 = left + right - len2;
        if (len2 <= 0) len2 = 1;
        h0 = s;
        len1 = len2 + std::min(std::max(len2 - 5, 1), (int)c.tail_cap);
    } else if (c.mode == BSW_GEN_UNIFORM) {
        len2 = (int)r.range(c.len2_min, c.len2_max);
        h0 = (int)r.range(c.h0_min, c.h0_max);
        len1 = len2 + std::min(std::max(len2 - 5, 1), (int)c.tail_cap);
    } else {
        double lo = std::log((double)c.len2_min), hi = std::log((double)c.len2_max);
        len2 = (int)std::lround(std::exp(lo + (hi - lo) * r.unit()));
        len2 = std::min(std::max(len2, (int)c.len2_min), (int)c.len2_max);
        h0 = (int)r.range(c.h0_min, c.h0_max);
        len1 = len2 + (int)r.range(0, c.extra_max);
    }
    if (r.unit() < c.small_h0_frac) h0 = (int)(r.u64() & 1);

    size_t q0 = qer.size(), t0 = ref.size();
    for (int j = 0; j < len2; ++j) qer.push_back(r.base());

    if (r.unit() < c.random_frac) {
        for (int i = 0; i < len1; ++i) ref.push_back(r.base());
    } else {
        int produced = 0;
        for (int j = 0; j < len2 && produced < len1; ++j) {
            double u = r.unit();
            if (u < c.indel_rate * 0.5) continue;  // base missing from the target
            if (u < c.indel_rate) {                 // extra target base
                ref.push_back(r.base());
                if (++produced >= len1) break;
            }
            uint8_t b = qer[q0 + j];
            if (r.unit() < c.sub_rate) b = (uint8_t)((b + 1 + (r.u64() % 3)) & 3);
            ref.push_back(b);
            ++produced;
        }
        for (; produced < len1; ++produced) ref.push_back(r.base());
    }
    if (r.unit() < c.n_frac) {
        if (r.u64() & 1) ref[t0 + (size_t)r.range(0, len1 - 1)] = 4;
        else qer[q0 + (size_t)r.range(0, len2 - 1)] = 4;
    }
    sp.idr = (int64_t)t0; sp.idq = (int64_t)q0;  // chunk-local for now
    sp.len1 = len1; sp.len2 = len2; sp.h0 = h0;
    sp.seqid = sp.regid = sp.score = sp.tle = sp.gtle = sp.qle = -1;
    sp.gscore = sp.max_off = -1;
}

int hw_threads(int req) {
    if (req > 0) return req;
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

}  // namespace

extern "C" {

int bsw_gen_preset(int id, bsw_gen_config *o) {
    if (!o) return 1;
    bsw_gen_config c;
    memset(&c, 0, sizeof c);
    c.read_len = 151; c.seed_min = 19; c.seed_max = 120; c.tail_cap = 200;
    c.sub_rate = 0.03; c.indel_rate = 0.006;
    c.n_frac = 0.02; c.small_h0_frac = 0.01; c.random_frac = 0.01;
    switch (id) {
        case 1: c.mode = BSW_GEN_READ_FLANK; c.seed = 1001; break;
        case 2: c.mode = BSW_GEN_UNIFORM; c.len2_min = 250; c.len2_max = 300; c.h0_min = 19;
                c.h0_max = 150; c.seed = 1002; break;
        case 3: c.mode = BSW_GEN_READ_FLANK; c.seed = 1003; break;
        case 4: c.mode = BSW_GEN_LOGUNIFORM; c.len2_min = 30; c.len2_max = 1000; c.h0_min = 19;
                c.h0_max = 150; c.extra_max = 200; c.sub_rate = 0.08; c.indel_rate = 0.02;
                c.seed = 1004; break;
        case 5: c.mode = BSW_GEN_READ_FLANK; c.seed = 1005; break;
        default: return 1;
    }
    *o = c;
    return 0;
}

int bsw_gen_pairs(const bsw_gen_config *cfg, int64_t n, bsw_seqpair *pairs, uint8_t **ref_out,
                  uint8_t **qer_out, int64_t *ref_bytes, int64_t *qer_bytes, int32_t nthreads) {
    if (!cfg || !pairs || !ref_out || !qer_out || n < 0) return 1;
    const int64_t nchunks = (n + kChunk - 1) / kChunk;
    std::vector<ChunkOut> chunks((size_t)nchunks);
    const int T = (int)std::min<int64_t>(hw_threads(nthreads), std::max<int64_t>(nchunks, 1));

    auto work = [&](int tid) {
        for (int64_t c = tid; c < nchunks; c += T) {
            Rng r(cfg->seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(c + 1));
            ChunkOut &o = chunks[(size_t)c];
            int64_t lo = c * kChunk, hi = std::min(n, lo + kChunk);
            o.ref.reserve((size_t)(hi - lo) * 160);
            o.qer.reserve((size_t)(hi - lo) * 80);
            for (int64_t k = lo; k < hi; ++k) {
                gen_one(r, *cfg, o.ref, o.qer, pairs[k]);
                pairs[k].id = k;
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    std::vector<int64_t> roff((size_t)nchunks + 1, 0), qoff((size_t)nchunks + 1, 0);
    for (int64_t c = 0; c < nchunks; ++c) {
        roff[(size_t)c + 1] = roff[(size_t)c] + (int64_t)chunks[(size_t)c].ref.size();
        qoff[(size_t)c + 1] = qoff[(size_t)c] + (int64_t)chunks[(size_t)c].qer.size();
    }
    // 64 bytes of slack so vectorised readers may overrun the last sequence
    uint8_t *ref = (uint8_t *)malloc((size_t)roff[(size_t)nchunks] + 64);
    uint8_t *qer = (uint8_t *)malloc((size_t)qoff[(size_t)nchunks] + 64);
    if (!ref || !qer) { free(ref); free(qer); return 2; }
    memset(ref + roff[(size_t)nchunks], 0, 64);
    memset(qer + qoff[(size_t)nchunks], 0, 64);
    auto place = [&](int tid) {
        for (int64_t c = tid; c < nchunks; c += T) {
            ChunkOut &o = chunks[(size_t)c];
            memcpy(ref + roff[(size_t)c], o.ref.data(), o.ref.size());
            memcpy(qer + qoff[(size_t)c], o.qer.data(), o.qer.size());
            int64_t lo = c * kChunk, hi = std::min(n, lo + kChunk);
            for (int64_t k = lo; k < hi; ++k) {
                pairs[k].idr += roff[(size_t)c];
                pairs[k].idq += qoff[(size_t)c];
            }
            std::vector<uint8_t>().swap(o.ref);
            std::vector<uint8_t>().swap(o.qer);
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(place, t);
        for (auto &t : th) t.join();
    }
    *ref_out = ref; *qer_out = qer;
    if (ref_bytes) *ref_bytes = roff[(size_t)nchunks];
    if (qer_bytes) *qer_bytes = qoff[(size_t)nchunks];
    return 0;
}

void bsw_host_free(void *p) { free(p); }

int bsw_write_pairs_text(const char *path, const bsw_seqpair *pairs, const uint8_t *ref,
                         const uint8_t *qer, int64_t n) {
    FILE *f = fopen(path, "w");
    if (!f) return 1;
    std::string line;
    for (int64_t k = 0; k < n; ++k) {
        const bsw_seqpair &p = pairs[k];
        line.clear();
        line += std::to_string(p.h0);
        line += '\n';
        for (int i = 0; i < p.len1; ++i) line += (char)('0' + ref[p.idr + i]);
        line += '\n';
        for (int j = 0; j < p.len2; ++j) line += (char)('0' + qer[p.idq + j]);
        line += '\n';
        if (fwrite(line.data(), 1, line.size(), f) != line.size()) { fclose(f); return 2; }
    }
    return fclose(f) ? 3 : 0;
}

int64_t bsw_count_pairs_text(const char *path) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    std::vector<char> buf(1 << 20);
    int64_t lines = 0;
    size_t got;
    char last = '\n';
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        lines += std::count(buf.begin(), buf.begin() + (long)got, '\n');
        last = buf[got - 1];
    }
    fclose(f);
    if (last != '\n') ++lines;   // a last line without a newline still counts (the reference would drop it)
    return lines / 3;
}

// Parallel parse: the file is read in one piece, line starts are indexed by T threads (newline count
// per block, prefix sum, fill), then pairs are parsed by T threads into dense buffers whose offsets come
// from a prefix sum over the line lengths. (The reference's loader is a serial fgets loop,
// main_banded.cpp:164-206.)
int64_t bsw_read_pairs_text(const char *path, int64_t n, bsw_seqpair *pairs, uint8_t **ref_out,
                            uint8_t **qer_out, int64_t *ref_bytes, int64_t *qer_bytes) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    const int64_t fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)fsz + 1);
    if (fsz > 0 && fread(buf.data(), 1, (size_t)fsz, f) != (size_t)fsz) { fclose(f); return -1; }
    fclose(f);
    buf[(size_t)fsz] = '\n';
    const int T = hw_threads(0);
    const int64_t blk = (fsz + T - 1) / std::max(T, 1) + 1;
    std::vector<int64_t> cnt((size_t)T + 1, 0);
    auto in_threads = [&](auto fn) {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(fn, t);
        for (auto &x : th) x.join();
    };
    in_threads([&](int t) {
        const int64_t lo = std::min<int64_t>(fsz, t * blk), hi = std::min<int64_t>(fsz, lo + blk);
        cnt[(size_t)t + 1] = std::count(buf.begin() + lo, buf.begin() + hi, '\n');
    });
    for (int t = 0; t < T; ++t) cnt[(size_t)t + 1] += cnt[(size_t)t];
    int64_t nlines = cnt[(size_t)T];
    if (fsz > 0 && buf[(size_t)fsz - 1] != '\n') ++nlines;      // last line without a newline
    std::vector<int64_t> start((size_t)nlines + 1);              // start[l] = offset of line l
    if (nlines > 0) start[0] = 0;
    in_threads([&](int t) {
        const int64_t lo = std::min<int64_t>(fsz, t * blk), hi = std::min<int64_t>(fsz, lo + blk);
        int64_t l = cnt[(size_t)t];
        for (int64_t o = lo; o < hi; ++o)
            if (buf[(size_t)o] == '\n' && l + 1 <= nlines) start[(size_t)++l] = o + 1;
    });
    start[(size_t)nlines] = std::max<int64_t>(start[(size_t)nlines], 0);
    if (fsz > 0 && buf[(size_t)fsz - 1] != '\n') start[(size_t)nlines] = fsz + 1;
    const int64_t np = std::min<int64_t>(n, nlines / 3);
    auto line_len = [&](int64_t l) {   // without the newline / carriage return
        int64_t e = start[(size_t)l + 1] - 1;
        if (e > start[(size_t)l] && buf[(size_t)e - 1] == '\r') --e;
        return e - start[(size_t)l];
    };
    // lengths, validation, offsets
    std::vector<int64_t> roff((size_t)np + 1, 0), qoff((size_t)np + 1, 0);
    std::vector<int> bad((size_t)T, 0);
    const int64_t per = (np + T - 1) / std::max(T, 1);
    in_threads([&](int t) {
        for (int64_t k = t * per; k < std::min(np, (t + 1) * per); ++k) {
            const int64_t l1 = line_len(3 * k + 1), l2 = line_len(3 * k + 2);
            if (l1 <= 0 || l2 <= 0 || l1 > BSW_MAX_SEQ_LEN || l2 > BSW_MAX_SEQ_LEN) bad[(size_t)t] = 1;
            roff[(size_t)k + 1] = l1; qoff[(size_t)k + 1] = l2;
        }
    });
    for (int t = 0; t < T; ++t) if (bad[(size_t)t]) return -1;   // the reference asserts len > 0 (main_banded.cpp:187-188)
    for (int64_t k = 0; k < np; ++k) { roff[(size_t)k + 1] += roff[(size_t)k]; qoff[(size_t)k + 1] += qoff[(size_t)k]; }
    uint8_t *r = (uint8_t *)malloc((size_t)roff[(size_t)np] + 64), *q = (uint8_t *)malloc((size_t)qoff[(size_t)np] + 64);
    if (!r || !q) { free(r); free(q); return -1; }
    memset(r + roff[(size_t)np], 0, 64);
    memset(q + qoff[(size_t)np], 0, 64);
    in_threads([&](int t) {
        for (int64_t k = t * per; k < std::min(np, (t + 1) * per); ++k) {
            bsw_seqpair &sp = pairs[k];
            sp.id = k; sp.idr = roff[(size_t)k]; sp.idq = qoff[(size_t)k];
            sp.len1 = (int32_t)(roff[(size_t)k + 1] - roff[(size_t)k]);
            sp.len2 = (int32_t)(qoff[(size_t)k + 1] - qoff[(size_t)k]);
            sp.h0 = atoi(buf.data() + start[(size_t)(3 * k)]);
            sp.seqid = sp.regid = sp.score = sp.tle = sp.gtle = sp.qle = -1;
            sp.gscore = sp.max_off = -1;
            const char *s1 = buf.data() + start[(size_t)(3 * k + 1)], *s2 = buf.data() + start[(size_t)(3 * k + 2)];
            for (int i = 0; i < sp.len1; ++i) r[sp.idr + i] = (uint8_t)(s1[i] - '0');
            for (int j = 0; j < sp.len2; ++j) q[sp.idq + j] = (uint8_t)(s2[j] - '0');
        }
    });
    *ref_out = r; *qer_out = q;
    if (ref_bytes) *ref_bytes = roff[(size_t)np];
    if (qer_bytes) *qer_bytes = qoff[(size_t)np];
    return np;
}

// ---- packed binary pair file ---------------------------------------------------
//   header  : "BSWPAIR1", u64 n, u64 data bytes
//   records : n x { u16 len1, u16 len2, i32 h0, u32 flags }   (12 bytes; flags bit 0 = 4 bits per base)
//   data    : per pair, query then target, 2 bits per base (4 if the pair holds a base code >= 4), each
//             sequence padded to 4 bytes -- the layout of the library's device blob slots
// About 3.5x smaller than the text format and parsed at memory speed.
namespace {
struct PackedRec { uint16_t len1, len2; int32_t h0; uint32_t flags; };
static_assert(sizeof(PackedRec) == 12, "record layout");
const char kMagic[8] = {'B', 'S', 'W', 'P', 'A', 'I', 'R', '1'};
inline uint64_t seq_bytes_io(uint32_t len, bool wide) {
    uint64_t b = wide ? (len + 1) >> 1 : (len + 3) >> 2;
    return (b + 3u) & ~(uint64_t)3u;
}
void pack_seq(const uint8_t *src, int len, bool wide, uint8_t *dst) {
    memset(dst, 0, (size_t)seq_bytes_io((uint32_t)len, wide));
    if (wide) for (int i = 0; i < len; ++i) dst[i >> 1] |= (uint8_t)((src[i] > 4 ? 4 : src[i]) << (4 * (i & 1)));
    else for (int i = 0; i < len; ++i) dst[i >> 2] |= (uint8_t)((src[i] & 3) << (2 * (i & 3)));
}
void unpack_seq(const uint8_t *src, int len, bool wide, uint8_t *dst) {
    if (wide) for (int i = 0; i < len; ++i) dst[i] = (src[i >> 1] >> (4 * (i & 1))) & 0xF;
    else for (int i = 0; i < len; ++i) dst[i] = (src[i >> 2] >> (2 * (i & 3))) & 3;
}
}  // namespace

int bsw_write_pairs_packed(const char *path, const bsw_seqpair *pairs, const uint8_t *ref,
                           const uint8_t *qer, int64_t n) {
    if (!path || n < 0 || (n > 0 && (!pairs || !ref || !qer))) return 1;
    std::vector<PackedRec> rec((size_t)n);
    std::vector<uint64_t> off((size_t)n + 1, 0);
    const int T = hw_threads(0);
    const int64_t per = (n + T - 1) / std::max(T, 1);
    auto in_threads = [&](auto fn) {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(fn, t);
        for (auto &x : th) x.join();
    };
    std::vector<int> bad((size_t)T, 0);
    in_threads([&](int t) {
        for (int64_t k = t * per; k < std::min(n, (t + 1) * per); ++k) {
            const bsw_seqpair &p = pairs[k];
            if (p.len1 < 0 || p.len2 < 0 || p.len1 > BSW_MAX_SEQ_LEN || p.len2 > BSW_MAX_SEQ_LEN) { bad[(size_t)t] = 1; continue; }
            bool wide = false;
            for (int i = 0; i < p.len1 && !wide; ++i) wide = ref[p.idr + i] > 3;
            for (int j = 0; j < p.len2 && !wide; ++j) wide = qer[p.idq + j] > 3;
            rec[(size_t)k] = PackedRec{(uint16_t)p.len1, (uint16_t)p.len2, p.h0, wide ? 1u : 0u};
            off[(size_t)k + 1] = seq_bytes_io((uint32_t)p.len2, wide) + seq_bytes_io((uint32_t)p.len1, wide);
        }
    });
    for (int t = 0; t < T; ++t) if (bad[(size_t)t]) return 1;
    for (int64_t k = 0; k < n; ++k) off[(size_t)k + 1] += off[(size_t)k];
    std::vector<uint8_t> data((size_t)off[(size_t)n]);
    in_threads([&](int t) {
        for (int64_t k = t * per; k < std::min(n, (t + 1) * per); ++k) {
            const bsw_seqpair &p = pairs[k];
            const bool wide = rec[(size_t)k].flags & 1u;
            uint8_t *d = data.data() + off[(size_t)k];
            pack_seq(qer + p.idq, p.len2, wide, d);
            pack_seq(ref + p.idr, p.len1, wide, d + seq_bytes_io((uint32_t)p.len2, wide));
        }
    });
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    const uint64_t hdr[2] = {(uint64_t)n, off[(size_t)n]};
    bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(hdr, 8, 2, f) == 2 &&
              (n == 0 || fwrite(rec.data(), sizeof(PackedRec), (size_t)n, f) == (size_t)n) &&
              (data.empty() || fwrite(data.data(), 1, data.size(), f) == data.size());
    return (fclose(f) == 0 && ok) ? 0 : 2;
}

int64_t bsw_count_pairs_packed(const char *path) {
    int64_t n = 0;
    return bsw_packed_file_info(path, &n, nullptr) == 0 ? n : -1;   // (header checked against the file size)
}

int64_t bsw_read_pairs_packed(const char *path, int64_t n, bsw_seqpair *pairs, uint8_t **ref_out,
                              uint8_t **qer_out, int64_t *ref_bytes, int64_t *qer_bytes) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    char m[8];
    uint64_t hdr[2];
    if (fread(m, 1, 8, f) != 8 || memcmp(m, kMagic, 8) != 0 || fread(hdr, 8, 2, f) != 2) { fclose(f); return -1; }
    // a truncated or corrupt header must not drive the allocations below
    {
        const long pos = ftell(f);
        fseek(f, 0, SEEK_END);
        const long fsize = ftell(f);
        fseek(f, pos, SEEK_SET);
        if (pos < 0 || fsize < 0 || hdr[0] > (uint64_t)fsize / sizeof(PackedRec) ||
            hdr[1] > (uint64_t)fsize || hdr[0] * sizeof(PackedRec) + hdr[1] + 24 > (uint64_t)fsize) { fclose(f); return -1; }
    }
    const int64_t total = (int64_t)hdr[0], np = std::min<int64_t>(n, total);
    std::vector<PackedRec> rec;
    std::vector<uint8_t> data;
    try { rec.resize((size_t)total); data.resize((size_t)hdr[1]); } catch (...) { fclose(f); return -1; }
    if ((total && fread(rec.data(), sizeof(PackedRec), (size_t)total, f) != (size_t)total) ||
        (!data.empty() && fread(data.data(), 1, data.size(), f) != data.size())) { fclose(f); return -1; }
    fclose(f);
    std::vector<uint64_t> doff((size_t)np + 1, 0), roff((size_t)np + 1, 0), qoff((size_t)np + 1, 0);
    for (int64_t k = 0; k < np; ++k) {
        const PackedRec &r = rec[(size_t)k];
        const bool wide = r.flags & 1u;
        if (r.len1 > BSW_MAX_SEQ_LEN || r.len2 > BSW_MAX_SEQ_LEN) return -1;
        doff[(size_t)k + 1] = doff[(size_t)k] + seq_bytes_io(r.len2, wide) + seq_bytes_io(r.len1, wide);
        roff[(size_t)k + 1] = roff[(size_t)k] + r.len1;
        qoff[(size_t)k + 1] = qoff[(size_t)k] + r.len2;
    }
    if (doff[(size_t)np] > data.size()) return -1;
    uint8_t *rr = (uint8_t *)malloc((size_t)roff[(size_t)np] + 64), *qq = (uint8_t *)malloc((size_t)qoff[(size_t)np] + 64);
    if (!rr || !qq) { free(rr); free(qq); return -1; }
    memset(rr + roff[(size_t)np], 0, 64);
    memset(qq + qoff[(size_t)np], 0, 64);
    const int T = hw_threads(0);
    const int64_t per = (np + T - 1) / std::max(T, 1);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            for (int64_t k = t * per; k < std::min(np, (t + 1) * per); ++k) {
                const PackedRec &r = rec[(size_t)k];
                const bool wide = r.flags & 1u;
                bsw_seqpair &sp = pairs[k];
                sp.id = k; sp.idr = (int64_t)roff[(size_t)k]; sp.idq = (int64_t)qoff[(size_t)k];
                sp.len1 = r.len1; sp.len2 = r.len2; sp.h0 = r.h0;
                sp.seqid = sp.regid = sp.score = sp.tle = sp.gtle = sp.qle = -1;
                sp.gscore = sp.max_off = -1;
                const uint8_t *d = data.data() + doff[(size_t)k];
                unpack_seq(d, r.len2, wide, qq + sp.idq);
                unpack_seq(d + seq_bytes_io(r.len2, wide), r.len1, wide, rr + sp.idr);
            }
        });
    for (auto &x : th) x.join();
    *ref_out = rr; *qer_out = qq;
    if (ref_bytes) *ref_bytes = (int64_t)roff[(size_t)np];
    if (qer_bytes) *qer_bytes = (int64_t)qoff[(size_t)np];
    return np;
}

// ---- the packed form in memory (bsw_gpu_batch_packed) ------------------------------------------------------
static_assert(sizeof(bsw_packed_rec) == sizeof(PackedRec), "the file's record is bsw_packed_rec");

int64_t bsw_packed_bytes(const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer, int64_t n) {
    if (n < 0 || (n > 0 && (!pairs || !ref || !qer))) return -1;
    int64_t total = 0;
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : total) reduction(| : bad)
    for (int64_t k = 0; k < n; ++k) {
        const bsw_seqpair &p = pairs[k];
        if (p.len1 < 0 || p.len2 < 0 || p.len1 > BSW_MAX_SEQ_LEN || p.len2 > BSW_MAX_SEQ_LEN) { bad = 1; continue; }
        bool wide = false;
        for (int i = 0; i < p.len1 && !wide; ++i) wide = ref[p.idr + i] > 3;
        for (int j = 0; j < p.len2 && !wide; ++j) wide = qer[p.idq + j] > 3;
        total += (int64_t)(seq_bytes_io((uint32_t)p.len2, wide) + seq_bytes_io((uint32_t)p.len1, wide));
    }
    return bad ? -1 : total;
}

int bsw_pack_pairs(const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer, int64_t n,
                   bsw_packed_rec *rec, uint8_t *data, int64_t data_cap) {
    if (n < 0 || (n > 0 && (!pairs || !ref || !qer || !rec || !data))) return 1;
    std::vector<uint64_t> off((size_t)n + 1, 0);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t k = 0; k < n; ++k) {
        const bsw_seqpair &p = pairs[k];
        if (p.len1 < 0 || p.len2 < 0 || p.len1 > BSW_MAX_SEQ_LEN || p.len2 > BSW_MAX_SEQ_LEN) { bad = 1; continue; }
        bool wide = false;
        for (int i = 0; i < p.len1 && !wide; ++i) wide = ref[p.idr + i] > 3;
        for (int j = 0; j < p.len2 && !wide; ++j) wide = qer[p.idq + j] > 3;
        rec[k] = bsw_packed_rec{(uint16_t)p.len1, (uint16_t)p.len2, p.h0, wide ? 1u : 0u};
        off[(size_t)k + 1] = seq_bytes_io((uint32_t)p.len2, wide) + seq_bytes_io((uint32_t)p.len1, wide);
    }
    if (bad) return 1;
    for (int64_t k = 0; k < n; ++k) off[(size_t)k + 1] += off[(size_t)k];
    if ((int64_t)off[(size_t)n] > data_cap) return 2;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        const bsw_seqpair &p = pairs[k];
        const bool wide = rec[k].flags & 1u;
        uint8_t *d = data + off[(size_t)k];
        pack_seq(qer + p.idq, p.len2, wide, d);
        pack_seq(ref + p.idr, p.len1, wide, d + seq_bytes_io((uint32_t)p.len2, wide));
    }
    return 0;
}

int bsw_packed_file_info(const char *path, int64_t *n, int64_t *data_bytes) {
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) return 1;
    char m[8];
    uint64_t hdr[2];
    bool ok = fread(m, 1, 8, f) == 8 && memcmp(m, kMagic, 8) == 0 && fread(hdr, 8, 2, f) == 2;
    if (ok) {
        fseek(f, 0, SEEK_END);
        const long fsize = ftell(f);
        ok = fsize >= 24 && hdr[0] <= (uint64_t)fsize / sizeof(PackedRec) && hdr[1] <= (uint64_t)fsize &&
             hdr[0] * sizeof(PackedRec) + hdr[1] + 24 <= (uint64_t)fsize;
    }
    fclose(f);
    if (!ok) return 1;
    if (n) *n = (int64_t)hdr[0];
    if (data_bytes) *data_bytes = (int64_t)hdr[1];
    return 0;
}

int64_t bsw_read_packed_raw(const char *path, int64_t n, bsw_packed_rec *rec, uint8_t *data, int64_t data_cap) {
    int64_t total = 0, dbytes = 0;
    if (!rec || (!data && data_cap > 0) || bsw_packed_file_info(path, &total, &dbytes) != 0) return -1;
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    const int64_t np = std::min(n, total);
    bool ok = fseek(f, 24, SEEK_SET) == 0 && (np == 0 || fread(rec, sizeof(PackedRec), (size_t)np, f) == (size_t)np);
    int64_t need = 0;
    for (int64_t k = 0; ok && k < np; ++k) {
        if (rec[k].len1 > BSW_MAX_SEQ_LEN || rec[k].len2 > BSW_MAX_SEQ_LEN) ok = false;
        need += (int64_t)(seq_bytes_io(rec[k].len2, rec[k].flags & 1u) + seq_bytes_io(rec[k].len1, rec[k].flags & 1u));
    }
    ok = ok && need <= dbytes && need <= data_cap && fseek(f, 24 + (long)(total * (int64_t)sizeof(PackedRec)), SEEK_SET) == 0 &&
         (need == 0 || fread(data, 1, (size_t)need, f) == (size_t)need);
    fclose(f);
    return ok ? np : -1;
}

}  // extern "C"
