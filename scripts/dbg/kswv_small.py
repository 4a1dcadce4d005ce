"""developer probe: kswv_gpu_batch latency for the batch sizes one bwa-mem2 worker produces (a few thousand pairs)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from genarchbench_b200 import kswv
from oracle import kswv as ok
pairs, ref, qer = ok.make_workload(16000, seed=3, read_len=(151, 151))
g = kswv.Kswv()
g.align(pairs, ref, qer)
for n in (250, 1000, 2000, 4000, 8000, 16000):
    p = pairs[:n].copy()
    ts = []
    for _ in range(20):
        t = time.perf_counter(); g.align(p, ref, qer); ts.append(time.perf_counter() - t)
    st = g.stats()
    print(f"n={n:6d}  call {1e3*np.median(ts):7.3f} ms  kernel {st['kernel_ms']:7.3f} ms  {st['cells']/np.median(ts)/1e9:7.1f} GCUPS  chunks {st['chunks']}")
