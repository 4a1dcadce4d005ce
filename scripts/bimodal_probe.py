"""Probe of the host pass's bimodal speed on the pool's VM. Developer tool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genarchbench_b200 import pairio, bsw

n = 4_000_000
b = pairio.generate(3, n)
a = np.ones(1 << 27, dtype=np.uint8); c = np.empty_like(a)
t0 = time.perf_counter(); c[:] = a; t1 = time.perf_counter() - t0
t0 = time.perf_counter(); c[:] = a; t2 = time.perf_counter() - t0
print(f"numpy copy 128 MiB: {0.268/t2:.1f} GB/s (r+w)  affinity {sorted(os.sched_getaffinity(0))}", flush=True)
g1 = bsw.BswGpu(devices=[0])
w1 = b.copy()
for thr in (16, 8, 4):
    # libgomp honours omp_set_num_threads only through the env at load; use the stats of repeated runs instead
    g1.batch(w1.pairs, w1.ref, w1.qer, 100)
    t0 = time.perf_counter(); g1.batch(w1.pairs, w1.ref, w1.qer, 100); dt = time.perf_counter() - t0
    st = g1.stats()
    print(f"e2e {dt*1e3:.1f} ms pack {st['host_pack_ms']:.1f} ({n/1e6:.0f} M pairs)", flush=True)
    break
