import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genarchbench_b200 import bsw
L = bsw.lib()
L.bsw_gpu_trip_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
for kind in (0, 1):
    out = []
    for w in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32):
        v = C.c_double()
        L.bsw_gpu_trip_probe(0, kind, w, C.byref(v))
        out.append(f"{w}:{v.value:.0f}")
    print("kind", kind, " ".join(out))
