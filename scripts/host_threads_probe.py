"""Developer probe (run under gpurun): wall time of bsw_gpu_batch on 10 M config-3 pairs against the number of host
threads -- tells whether the host pass is bound by cores or by memory bandwidth.
   OMP_NUM_THREADS=k python scripts/host_threads_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OMP_PROC_BIND", "true"); os.environ.setdefault("OMP_PLACES", "cores")
from genarchbench_b200 import pairio, bsw
b = pairio.generate(3, 10_000_000, seed=1003)
g = bsw.BswGpu()
g.batch(b.pairs, b.ref, b.qer, 100)
ts = []
for _ in range(6):
    t0 = time.perf_counter(); g.batch(b.pairs, b.ref, b.qer, 100); ts.append((time.perf_counter() - t0) * 1e3)
st = g.stats()
print(json.dumps({"threads": os.environ.get("OMP_NUM_THREADS"), "ms_min": round(min(ts), 2), "ms_med": round(sorted(ts)[3], 2),
                  "pack_ms": round(st["host_pack_ms"], 2), "wait_ms": round(st["host_wait_ms"], 2)}))
