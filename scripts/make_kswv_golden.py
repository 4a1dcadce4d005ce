"""Generates tests/golden/kswv_*.npz from the compiled, UNMODIFIED reference (oracle/_ref/libkswv_ref_avx512.so,
built from /root/reference by `make -C oracle ref`). Needs an AVX512BW host and /root/reference; the GPU box has
neither, so the vectors are committed. Run: python scripts/make_kswv_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import kswv  # noqa: E402
from oracle.kswv import KSW_XBYTE, KSW_XSTART, KSW_XSTOP, KSW_XSUBO  # noqa: E402

CASES = {
    # name: (n, params, make_workload kwargs)
    "kswv_default": (400, None, dict(seed=11)),
    "kswv_16bit": (200, None, dict(seed=12, read_len=(250, 400))),
    "kswv_a2": (200, dict(match=2, mismatch=8, o_del=12, o_ins=12, e_del=2, e_ins=2),
                dict(seed=13, match=2, read_len=(80, 200))),
    "kswv_saturating8": (200, None, dict(seed=14, read_len=(240, 330), p_sub=0.005,
                                         xtra=lambda l: KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 19)),
    "kswv_tiny": (400, None, dict(seed=15, read_len=(1, 24), window=(0.5, 3.0), min_seed_len=3)),
    "kswv_stop": (200, None, dict(seed=16, xtra=lambda l: KSW_XSTOP | KSW_XSTART | (KSW_XBYTE if l < 120 else 0) | 45)),
    "kswv_ambiguous": (200, None, dict(seed=17, p_n=0.2)),
    "kswv_asym_gaps": (200, dict(o_del=4, e_del=2, o_ins=7, e_ins=1), dict(seed=18, p_indel=0.03)),
}

if __name__ == "__main__":
    assert kswv.reference_available(), "needs oracle/_ref/libkswv_ref_avx512.so and an AVX512BW host"
    out = os.path.join(ROOT, "tests", "golden")
    for name, (n, params, kw) in CASES.items():
        pairs, ref, qer = kswv.make_workload(n, **kw)
        aln = kswv.reference_batch(pairs, ref, qer, params)
        p = dict(kswv.DEFAULT_PARAMS)
        p.update(params or {})
        np.savez_compressed(os.path.join(out, name + ".npz"), len1=pairs["len1"], len2=pairs["len2"], h0=pairs["h0"],
                            idr=pairs["idr"], idq=pairs["idq"], ref=ref, qer=qer, aln=aln,
                            params=np.array([p[k] for k in kswv._ORDER], np.int32))
        print(name, n, "pairs", os.path.getsize(os.path.join(out, name + ".npz")), "bytes")
