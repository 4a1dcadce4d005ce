"""Scale check (run under gpurun): 50 M pairs, sequence offsets beyond 2^31, first and second call through
bsw_gpu_batch, a sample against the oracle. Developer tool."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genarchbench_b200 import pairio, bsw
import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
t0 = time.time(); b = pairio.generate(5, n)
print(f"generated {n} pairs in {time.time()-t0:.1f}s; ref {b.ref.nbytes/1e9:.1f} GB, max idr {int(b.pairs['idr'].max())}", flush=True)
with bsw.BswGpu(devices=[0]) as g:
    for call in ("first", "second"):
        t0 = time.time(); g.batch(b.pairs, b.ref, b.qer, 100); dt = time.time() - t0
        st = g.stats()
        print(f"{call} call: {dt*1e3:.0f} ms = {n/dt/1e6:.0f} M pairs/s; launches {st['kernel_launches']}, kernel {st['kernel_ms']:.0f} ms; "
              + " ".join(f"{k[5:-3]}={v:.0f}" for k, v in st.items() if k.startswith("host_")), flush=True)
idx = np.concatenate([np.random.default_rng(1).choice(n, 100000, replace=False), np.arange(n - 20000, n)])
s = pairio.PairBatch(b.pairs[idx].copy(), b.ref, b.qer)
want = s.copy(); oracle.oracle_batch(want)
print("sampled mismatches:", int((want.outputs() != b.outputs()[idx]).any(axis=1).sum()), "; unwritten:", int((b.pairs['score'] < 0).sum()))
