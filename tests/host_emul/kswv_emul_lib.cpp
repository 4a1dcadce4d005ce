// TEST INFRASTRUCTURE: runs the product's per-pair kswv device code (kswv_pair<W> from
// genarchbench_b200/csrc/kswv_kernels.cuh) on the CPU: DPX / PRMT through dpx_host_emul.h, the warp through the
// 32-fiber emulation in warp_fibers.h, 32 / W pairs per warp exactly as the kernel assigns them. Used only by
// tests/test_kswv_emulation.py to check the kernel's ALGORITHM against the oracle where no GPU exists; it is not a
// product path.
#define BSW_HOST_EMUL 1
#include "kswv_kernels.cuh"
#include "bsw_types.h"
#include <cstring>
#include <vector>

using namespace kswvk;

namespace {
template <int W>
int run_batch(const KParams &K, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer, int64_t n, int32_t *aln) {
    constexpr int G = 32 / W;
    int rc = 0;
#pragma omp parallel
    {
        wf::Warp warp;
        std::vector<uint32_t> rowkey[G], lutw[G];
        std::vector<uint2> bnd[G];
        std::vector<uint8_t> qbuf[G];
#pragma omp for schedule(dynamic, 4)
        for (int64_t i0 = 0; i0 < n; i0 += G) {
            Task T[G];
            for (int g = 0; g < G; ++g) {
                if (i0 + g < n) {
                    const bsw_seqpair &sp = pairs[i0 + g];
                    T[g] = Task{(uint32_t)sp.idr, (uint32_t)sp.idq, sp.len1, sp.len2, sp.h0, (int32_t)(i0 + g)};
                } else T[g] = Task{0u, 0u, 0, 0, 0, -1};
                // exactly what the kernel gets per group, poisoned so that a read of something nobody stored shows up
                rowkey[g].assign((size_t)T[g].tlen + 8, 0xDEADBEEFu);
                lutw[g].assign((size_t)T[g].tlen + 8, 0xDEADBEEFu);
                bnd[g].assign((size_t)T[g].tlen + 8, uint2{0xDEADBEEFu, 0xDEADBEEFu});
                qbuf[g].assign((size_t)T[g].qlen + 64, (uint8_t)0xEE);
            }
            Result res[32];
            wf::run_warp(warp, [&]() {
                const int g = hw_lane() / W;
                const Result r = kswv_pair<W>(K, T[g], ref, qer, rowkey[g].data(), bnd[g].data(), lutw[g].data(), qbuf[g].data());
                res[hw_lane()] = r;
            });
            for (int g = 0; g < G; ++g) {
                for (int l = 1; l < W; ++l)
                    if (memcmp(&res[g * W + l], &res[g * W], sizeof(Result)) != 0) rc = -2 - l;   // lanes of a group must agree
                if (T[g].out >= 0) memcpy(aln + 7 * (int64_t)pairs[i0 + g].regid, &res[g * W], sizeof(Result));
            }
        }
    }
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
