"""TEST INFRASTRUCTURE: runs the emulated kswv device code (all three lane-group widths) of an AddressSanitizer
build against the oracle. Started by tests/test_kswv_emulation.py with libasan preloaded; compute-sanitizer is not
available on the GPU pool, so this is the out-of-bounds check of the per-pair code (the same source the kernels
compile). The emulation sizes every scratch buffer exactly as the kernels' host side does."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import kswv  # noqa: E402

L = C.CDLL(sys.argv[1])
L.kswv_emul_batch.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_int32]
total = 0
CASES = [(32, dict(n=150, read_len=(1, 320), window=(0.3, 4.0), min_seed_len=5)),
         (32, dict(n=30, read_len=(257, 600))),
         (16, dict(n=150, read_len=(1, 249), window=(0.3, 4.0), min_seed_len=5)),
         (8, dict(n=150, read_len=(1, 160), window=(0.3, 4.0), min_seed_len=5))]
for width, kw in CASES:
    pairs, ref, qer = kswv.make_workload(seed=41, **kw)
    # sequences in exactly sized buffers: a read past the last base is an error
    rb, qb = int(pairs["idr"][-1] + pairs["len1"][-1]), int(pairs["idq"][-1] + pairs["len2"][-1])
    ref, qer = np.ascontiguousarray(ref[:rb]), np.ascontiguousarray(qer[:qb])
    want, _ = kswv.oracle_batch(pairs, np.concatenate([ref, np.zeros(64, np.uint8)]), np.concatenate([qer, np.zeros(64, np.uint8)]))
    got = np.full((len(pairs), 7), -7, np.int32)
    rc = L.kswv_emul_batch(kswv._params(None), pairs.ctypes.data, ref.ctypes.data, qer.ctypes.data, len(pairs), got.ctypes.data, width)
    total += int(rc != 0) + int((got != want).any(axis=1).sum())
print("kswv_asan_check mismatches", total)
sys.exit(1 if total else 0)
