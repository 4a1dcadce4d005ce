"""kswv throughput on one B200: a synthetic mate-rescue batch (151 bp reads against windows 2-5x their length)
through kswv_gpu_batch from page-locked host buffers. Prints kernel-only and end-to-end GCUPS (phase-0 cells, the
reference's padded columns included, per second) and, where the host can run it, the compiled AVX512 reference
(one thread, as mem_sam_pe_batch runs it) and the oracle port on all cores, on a bounded sample.
    python scripts/kswv_bench.py [--pairs N] [--reps R]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=200000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    a = ap.parse_args()
    from genarchbench_b200 import bsw, kswv
    from oracle import kswv as okswv
    base = min(a.pairs, 20000)
    pairs0, ref0, qer0 = okswv.make_workload(base, seed=7, read_len=(151, 151))
    # tile the base batch up to --pairs (sequence generation in Python is the slow part)
    reps = (a.pairs + base - 1) // base
    n = base * reps
    rb, qb = int(pairs0["idr"][-1] + pairs0["len1"][-1]), int(pairs0["idq"][-1] + pairs0["len2"][-1])
    L = bsw.lib()

    def pinned(nbytes, dtype):
        p = L.bsw_gpu_host_alloc(nbytes + 64)
        assert p
        return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p)).view(dtype)
    from genarchbench_b200.pairio import SEQPAIR_DTYPE
    pairs = pinned(n * SEQPAIR_DTYPE.itemsize, SEQPAIR_DTYPE)
    ref = pinned(rb * reps + 64, np.uint8)
    qer = pinned(qb * reps + 64, np.uint8)
    aln = pinned(n * 28, np.int32).reshape(n, 7)
    for r in range(reps):
        sl = slice(r * base, (r + 1) * base)
        pairs[sl] = pairs0
        pairs["idr"][sl] += r * rb
        pairs["idq"][sl] += r * qb
        pairs["regid"][sl] += r * base
        ref[r * rb:(r + 1) * rb] = ref0[:rb]
        qer[r * qb:(r + 1) * qb] = qer0[:qb]
    g = kswv.Kswv(n_gpus=a.gpus)
    g.align(pairs, ref, qer, aln)                      # warm-up: allocations
    want, cells0 = okswv.oracle_batch(pairs0, ref0, qer0)
    assert (aln[:base] == want).all() and (aln[-base:] == want).all(), "parity"
    wall, kern = [], []
    for _ in range(a.reps):
        t = time.perf_counter()
        g.align(pairs, ref, qer, aln)
        wall.append(time.perf_counter() - t)
        kern.append(g.stats()["kernel_ms"] * 1e-3)
    st = g.stats()
    cells = st["cells"]
    out = {"pairs": n, "cells": cells, "n_gpus": a.gpus, "chunks": st["chunks"], "gathered": st["gathered"],
           "kernel_gcups": cells / min(kern) / 1e9, "e2e_gcups": cells / min(wall) / 1e9,
           "kernel_ms": min(kern) * 1e3, "e2e_ms": min(wall) * 1e3, "h2d_bytes": st["h2d_bytes"],
           "d2h_bytes": st["d2h_bytes"], "parity_pairs_checked": 2 * base}
    m = min(a.cpu_sample, base)
    t = time.perf_counter()
    _, c = okswv.oracle_batch(pairs0[:m].copy(), ref0, qer0)
    out["cpu_oracle_gcups_all_cores"] = c / (time.perf_counter() - t) / 1e9
    out["cpu_cores"] = os.cpu_count()
    if okswv.reference_available():
        t = time.perf_counter()
        okswv.reference_batch(pairs0[:m].copy(), ref0, qer0)
        out["cpu_reference_avx512_gcups_1_thread"] = c / (time.perf_counter() - t) / 1e9
    print(json.dumps(out))
