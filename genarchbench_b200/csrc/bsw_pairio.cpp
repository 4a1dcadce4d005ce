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
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0)
        lines += std::count(buf.begin(), buf.begin() + (long)got, '\n');
    fclose(f);
    return lines / 3;
}

int64_t bsw_read_pairs_text(const char *path, int64_t n, bsw_seqpair *pairs, uint8_t **ref_out,
                            uint8_t **qer_out, int64_t *ref_bytes, int64_t *qer_bytes) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    std::vector<uint8_t> ref, qer;
    char *line = nullptr;
    size_t cap = 0;
    int64_t k = 0;
    auto chomp = [](char *s, ssize_t len) -> int {
        while (len > 0 && (s[len - 1] == '\n' || s[len - 1] == '\r')) --len;
        return (int)len;
    };
    while (k < n) {
        ssize_t len = getline(&line, &cap, f);
        if (len < 0) break;
        int h0 = atoi(line);
        len = getline(&line, &cap, f);
        if (len < 0) break;
        int l1 = chomp(line, len);
        size_t t0 = ref.size();
        for (int i = 0; i < l1; ++i) ref.push_back((uint8_t)(line[i] - '0'));
        len = getline(&line, &cap, f);
        if (len < 0) { ref.resize(t0); break; }
        int l2 = chomp(line, len);
        size_t q0 = qer.size();
        for (int j = 0; j < l2; ++j) qer.push_back((uint8_t)(line[j] - '0'));
        if (l1 <= 0 || l2 <= 0 || l1 > BSW_MAX_SEQ_LEN || l2 > BSW_MAX_SEQ_LEN) {
            free(line); fclose(f); return -1;  // the reference asserts len > 0 (main_banded.cpp:187-188)
        }
        bsw_seqpair &sp = pairs[k];
        sp.id = k; sp.idr = (int64_t)t0; sp.idq = (int64_t)q0;
        sp.len1 = l1; sp.len2 = l2; sp.h0 = h0;
        sp.seqid = sp.regid = sp.score = sp.tle = sp.gtle = sp.qle = -1;
        sp.gscore = sp.max_off = -1;
        ++k;
    }
    free(line);
    fclose(f);
    uint8_t *r = (uint8_t *)malloc(ref.size() + 64), *q = (uint8_t *)malloc(qer.size() + 64);
    if (!r || !q) { free(r); free(q); return -1; }
    memcpy(r, ref.data(), ref.size()); memset(r + ref.size(), 0, 64);
    memcpy(q, qer.data(), qer.size()); memset(q + qer.size(), 0, 64);
    *ref_out = r; *qer_out = q;
    if (ref_bytes) *ref_bytes = (int64_t)ref.size();
    if (qer_bytes) *qer_bytes = (int64_t)qer.size();
    return k;
}

}  // extern "C"
