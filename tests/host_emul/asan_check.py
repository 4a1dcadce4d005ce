"""TEST INFRASTRUCTURE: runs the emulated device code (whole rows, windowed rows, two pairs per thread) of
an AddressSanitizer build against the oracle. Started by tests/test_host_emulation.py with libasan
preloaded; compute-sanitizer is not available on the GPU pool, so this is the out-of-bounds check of the
per-pair code (the same source the kernels compile)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from genarchbench_b200 import pairio  # noqa: E402

L = C.CDLL(sys.argv[1])
total = 0
for name in ("bsw_emul_batch", "bsw_emul_batch_key", "bsw_emul_batch_win", "bsw_emul_batch_duo", "bsw_emul_batch_duo_key"):
    fn = getattr(L, name)
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int32]
    for w in (1, 3, 17, 100):
        c = pairio.preset(4)
        c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 500, 0, 100, 0.3, 0.15
        b = pairio.generate(c, 600, seed=w)
        a = b.copy()
        oracle.oracle_batch(a, w=w)
        fn(oracle._params_array(None), b.pairs.ctypes.data, b.ref.ctypes.data, b.qer.ctypes.data, len(b), w)
        total += int((a.outputs() != b.outputs()).any(axis=1).sum())
# the host packers: destinations of exactly slot + one spare word, so any further overrun is caught
L.bsw_emul_pack_check.argtypes = [C.c_int64, C.c_uint32]
L.bsw_emul_pack_check.restype = C.c_int64
total += int(L.bsw_emul_pack_check(20000, 7))
print("asan_check mismatches", total)
sys.exit(1 if total else 0)
