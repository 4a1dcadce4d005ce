"""Developer A/B tool (run under gpurun): resident-kernel time of one workload, with a sampled oracle check.
   python scripts/kbench.py [--workload 3] [--pairs 4000000] [--steps 5] [--check 30000] [--tag name]
Environment switches of the library (BSW_DUO2, BSW_KEY, ...) and BSW_GPU_LIB select what is measured."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genarchbench_b200 import pairio, bsw
import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--workload", type=int, default=3)
ap.add_argument("--pairs", type=int, default=4_000_000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--check", type=int, default=30000)
ap.add_argument("--tag", default="")
ap.add_argument("--w", type=int, default=100)
a = ap.parse_args()
b = pairio.generate(a.workload if a.workload != 5 else 3, a.pairs, seed=1000 + a.workload)
g = bsw.BswGpu()
g.stage(b.pairs, b.ref, b.qer, a.w)
cells = g.count_staged()
ms = [g.run_staged() for _ in range(a.steps + 2)][2:]
st = g.stats()
g.fetch_staged(b.pairs)
bad = -1
if a.check:
    idx = np.sort(np.random.default_rng(1).choice(len(b), min(len(b), a.check), replace=False))
    s = pairio.PairBatch(b.pairs[idx].copy(), b.ref, b.qer)
    got = s.outputs()
    oracle.oracle_batch(s, w=a.w)
    bad = int((got != s.outputs()).any(axis=1).sum())
best = min(ms)
print(json.dumps({"tag": a.tag, "workload": a.workload, "pairs": a.pairs, "ms_min": round(best, 3), "ms_mean": round(float(np.mean(ms)), 3),
                  "gcups": round(cells / (best * 1e-3) / 1e9, 1), "cells": cells, "runs": a.steps + 2, "mismatches": bad, "launches": st["kernel_launches"],
                  "duo": st["pairs_duo"], "keyed": st["pairs_keyed"], "short": st["pairs_short"], "long": st["pairs_long"],
                  "env": {k: v for k, v in os.environ.items() if k.startswith("BSW_")}}), flush=True)
g.close()
