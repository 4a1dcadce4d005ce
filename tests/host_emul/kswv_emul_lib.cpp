// TEST INFRASTRUCTURE: runs the product's per-warp kswv device code (kswv_pair from
// genarchbench_b200/csrc/kswv_kernels.cuh) on the CPU: DPX / PRMT through dpx_host_emul.h, the warp through the
// 32-fiber emulation in warp_fibers.h. Used only by tests/test_kswv_emulation.py to check the kernel's ALGORITHM
// against the oracle where no GPU exists; it is not a product path.
#define BSW_HOST_EMUL 1
#include "kswv_kernels.cuh"
#include "bsw_types.h"
#include <cstring>
#include <vector>

using namespace kswvk;

// params: {o_del, e_del, o_ins, e_ins, match, mismatch(+ve)}; aln[pairs[i].regid] = kswr_t of pair i
extern "C" int kswv_emul_batch(const int32_t *params, const bsw_seqpair *pairs, const uint8_t *ref, const uint8_t *qer,
                               int64_t n, int32_t *aln) {
    const KParams K = make_kparams(params[0], params[1], params[2], params[3], params[4], params[5]);
    int rc = 0;
#pragma omp parallel
    {
        wf::Warp warp;
        std::vector<uint32_t> rowmx;
        std::vector<uint2> bnd;
        std::vector<uint8_t> qbuf;
        std::vector<uint32_t> lutw;
#pragma omp for schedule(dynamic, 8)
        for (int64_t i = 0; i < n; ++i) {
            const bsw_seqpair &sp = pairs[i];
            Task T{(uint32_t)sp.idr, (uint32_t)sp.idq, sp.len1, sp.len2, sp.h0, (int32_t)i};
            // exactly what the kernel gets per warp, poisoned so that a read of a row nobody stored shows up
            rowmx.assign((size_t)sp.len1 + 1, 0xDEADBEEFu);
            bnd.assign((size_t)sp.len1 + 1, uint2{0xDEADBEEFu, 0xDEADBEEFu});
            qbuf.assign((size_t)sp.len2 + 64, (uint8_t)0xEE);
            lutw.assign((size_t)sp.len1 + 8, 0xDEADBEEFu);
            Result res[32];
            wf::run_warp(warp, [&]() {
                const Result r = kswv_pair(K, T, ref, qer, rowmx.data(), bnd.data(), lutw.data(), qbuf.data());
                res[w_lane()] = r;
            });
            for (int l = 1; l < 32; ++l)
                if (memcmp(&res[l], &res[0], sizeof(Result)) != 0) rc = -2 - l;     // lanes must agree
            memcpy(aln + 7 * (int64_t)sp.regid, &res[0], sizeof(Result));
        }
    }
    return rc;
}
