// TEST INFRASTRUCTURE: runs the product's per-pair kswv device code (kswv_phase0<W> / kswv_phase1<W> from
// genarchbench_b200/csrc/kswv_kernels.cuh) on the CPU: DPX / PRMT through dpx_host_emul.h, the warp through the
// 32-fiber emulation in warp_fibers.h, 32 / W pairs per warp exactly as the kernels assign them, phase-1 tasks
// re-ordered by the key phase 0 leaves (the kernels' radix sort is a std::stable_sort here). Used only by
// tests/test_kswv_emulation.py to check the kernel's ALGORITHM against the oracle where no GPU exists; it is not a
// product path.
#define BSW_HOST_EMUL 1
#include "kswv_kernels.cuh"
#include "bsw_types.h"
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

using namespace kswvk;

namespace {
struct Scratch {
    std::vector<uint32_t> rowkey, lutw;
    std::vector<uint2> bnd;
    std::vector<uint8_t> qbuf;
    // exactly what the kernel gets per group, poisoned so that a read of something nobody stored shows up
    void reset(const Task &T) {
        rowkey.assign((size_t)T.tlen + kScratchSlack, 0xDEADBEEFu);
        lutw.assign((size_t)T.tlen + kScratchSlack, 0xDEADBEEFu);
        bnd.assign((size_t)T.tlen + 8, uint2{0xDEADBEEFu, 0xDEADBEEFu});
        qbuf.assign((size_t)T.qlen + 64, (uint8_t)0xEE);
    }
};

template <int W>
int run_batch(const KParams &K, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer, int64_t n, int32_t *aln) {
    constexpr int G = 32 / W;
    int rc = 0;
    std::vector<Task> tasks((size_t)n);
    std::vector<Result> out((size_t)n);
    std::vector<uint32_t> key((size_t)n), order((size_t)n);
    for (int64_t i = 0; i < n; ++i)
        tasks[(size_t)i] = Task{(uint32_t)pairs[i].idr, (uint32_t)pairs[i].idq, pairs[i].len1, pairs[i].len2, pairs[i].h0, (int32_t)i};
    for (int phase = 0; phase < 2; ++phase) {
        int64_t ntask = n;
        if (phase == 1) {
            std::iota(order.begin(), order.end(), 0u);
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] > key[b]; });
            ntask = 0;
            while (ntask < n && key[order[(size_t)ntask]] != 0u) ++ntask;
        }
#pragma omp parallel
        {
            wf::Warp warp;
            Scratch sc[G];
#pragma omp for schedule(dynamic, 4)
            for (int64_t j0 = 0; j0 < ntask; j0 += G) {
                int idx[G];
                for (int g = 0; g < G; ++g) {
                    idx[g] = j0 + g < ntask ? (int)(phase ? order[(size_t)(j0 + g)] : (uint32_t)(j0 + g)) : -1;
                    if (idx[g] >= 0) sc[g].reset(tasks[(size_t)idx[g]]);
                }
                Result res[32];
                uint32_t keys[32];
                wf::run_warp(warp, [&]() {
                    const int g = hw_lane() / W;
                    if (idx[g] < 0) return;
                    const Task &T = tasks[(size_t)idx[g]];
                    if (phase == 0) {
                        res[hw_lane()] = kswv_phase0<W>(K, T, ref, qer, sc[g].rowkey.data(), sc[g].bnd.data(), sc[g].lutw.data(),
                                                        &keys[hw_lane()]);
                    } else {
                        Result r = out[(size_t)T.out];
                        kswv_phase1<W>(K, T, ref, qer, sc[g].rowkey.data(), sc[g].bnd.data(), sc[g].lutw.data(), sc[g].qbuf.data(), &r);
                        res[hw_lane()] = r;
                    }
                });
                for (int g = 0; g < G; ++g) {
                    if (idx[g] < 0) continue;
                    for (int l = 1; l < W; ++l)                                                   // lanes of a group must agree
                        if (memcmp(&res[g * W + l], &res[g * W], sizeof(Result)) != 0 || (phase == 0 && keys[g * W + l] != keys[g * W]))
                            rc = -2 - l;
                    out[(size_t)idx[g]] = res[g * W];
                    if (phase == 0) key[(size_t)idx[g]] = keys[g * W];
                }
            }
        }
    }
    for (int64_t i = 0; i < n; ++i) memcpy(aln + 7 * (int64_t)pairs[i].regid, &out[(size_t)i], sizeof(Result));
    return rc;
}
}  // namespace

// params: {o_del, e_del, o_ins, e_ins, match, mismatch(+ve)}; aln[pairs[i].regid] = kswr_t of pair i.
// W = lanes per pair (8, 16, 32). Returns -1 if a pair does not fit the width (the host routes those to W = 32).
extern "C" int kswv_emul_batch(const int32_t *params, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                               int64_t n, int32_t *aln, int32_t W) {
    const KParams K = make_kparams(params[0], params[1], params[2], params[3], params[4], params[5]);
    if (W != 32)
        for (int64_t i = 0; i < n; ++i) {
            const bool byte = (pairs[i].h0 & kXByte) != 0;
            if (padded_cols(pairs[i].len2, byte) > group_cols(W) || needs_sat(K.a, K.shift, pairs[i].len1, pairs[i].len2, byte))
                return -1;
        }
    if (W == 32) return run_batch<32>(K, pairs, ref, qer, n, aln);
    if (W == 16) return run_batch<16>(K, pairs, ref, qer, n, aln);
    if (W == 8) return run_batch<8>(K, pairs, ref, qer, n, aln);
    return -1;
}
