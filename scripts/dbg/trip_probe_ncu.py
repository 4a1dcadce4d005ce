import ctypes as C, sys, os
sys.path.insert(0, '/root/repo')
from genarchbench_b200 import bsw
L = bsw.lib()
v = C.c_double()
for kind in (0, 1):
    L.bsw_gpu_trip_probe(0, kind, 32, C.byref(v)); print(kind, v.value)
