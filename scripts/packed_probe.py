"""Developer probe (run under gpurun): wall time of bsw_gpu_batch_packed on 10 M config-3 pairs from page-locked memory."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genarchbench_b200 import pairio, bsw
b = pairio.generate(3, 10_000_000, seed=1003)
rec, data = pairio.pack(b, bsw.host_alloc)
out = bsw.host_alloc(len(b) * 16).view(pairio.RESULT_DTYPE)
g = bsw.BswGpu()
g.batch_packed(rec, data, 100, out)
ts = []
for _ in range(8):
    t0 = time.perf_counter(); g.batch_packed(rec, data, 100, out); ts.append((time.perf_counter() - t0) * 1e3)
st = g.stats()
print(json.dumps({"slab": os.environ.get("BSW_SLAB_PAIRS"), "ms_min": round(min(ts), 2), "ms_med": round(sorted(ts)[4], 2),
                  "plan_ms": round(st["host_plan_ms"], 2), "cut_ms": round(st["host_cut_ms"], 2), "wait_ms": round(st["host_wait_ms"], 2),
                  "kernel_ms_sum": round(st["kernel_ms"], 2)}))
