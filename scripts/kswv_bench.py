"""Measurement of the kswv row in the bench line's vocabulary: one JSON line on stdout.

Workload: a synthetic mate-rescue batch shaped like mem_matesw_batch_pre's (bwamem_pair.cpp:930-1090): --read-len bp
reads (default 151) against reference windows 2-5x their length, h0 = KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 19.
metric / unit: giga cell updates per second, a cell = one (reference row, padded query column) of the forward pass --
the reference kernels' own loop nest (kswv.cpp:483-507); the reverse pass (phase 1) is work on top that the unit does
not count, on either side.
  value     kswv_gpu_batch with everything resident (kernel event time of the call's chunks, summed)
  e2e       the same call from page-locked host buffers, host clock around it, H2D and D2H inside
  roofline  ALU-pipe bound: 4.5 ALU-pipe instructions per computed cell (PRMT, VIMNMX3.RELU, 2 x VIADDMNMX.RELU,
            half a VIMNMX3 for the row key) x all cells the kernels compute (both phases, from the results)
            against the VIADDMNMX issue rate measured live on this GPU (bsw_gpu_dpx_peak)
  cpu_baseline  the compiled, unmodified AVX-512 reference (oracle/_ref/libkswv_ref_avx512.so) where the host has
            AVX512BW -- one instance per host thread, as bwa-mem2's workers run it -- else the scalar oracle port
    python scripts/kswv_bench.py [--pairs N] [--steps K] [--warmup W] [--gpus G] [--read-len L]"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ALU_INSTR_PER_CELL = 4.5

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=400000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reps", type=int, default=0, help="alias of --steps")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--read-len", type=int, default=151)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--pageable", action="store_true", help="caller buffers from malloc instead of page-locked memory")
    a = ap.parse_args()
    if a.reps:
        a.steps = a.reps
    from genarchbench_b200 import bsw, kswv
    from genarchbench_b200.pairio import SEQPAIR_DTYPE
    from oracle import kswv as okswv
    from genarchbench_b200.benchutil import ClockSampler
    base = min(a.pairs, 20000)
    pairs0, ref0, qer0 = okswv.make_workload(base, seed=7, read_len=(a.read_len, a.read_len),
                                             match=1, min_seed_len=19)
    # tile the base batch up to --pairs (sequence generation in Python is the slow part)
    reps = (a.pairs + base - 1) // base
    n = base * reps
    rb, qb = int(pairs0["idr"][-1] + pairs0["len1"][-1]), int(pairs0["idq"][-1] + pairs0["len2"][-1])
    L = bsw.lib()

    def pinned(nbytes, dtype):
        if a.pageable:
            return np.zeros(nbytes, np.uint8).view(dtype)
        p = L.bsw_gpu_host_alloc(nbytes + 64)
        assert p
        return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p)).view(dtype)
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
    for _ in range(max(a.warmup, 1)):
        g.align(pairs, ref, qer, aln)
    want, cells0 = okswv.oracle_batch(pairs0, ref0, qer0)
    mism = int((aln[:base] != want).any(axis=1).sum() + (aln[-base:] != want).any(axis=1).sum())
    sampler = ClockSampler(0)
    sampler.start()
    wall, kern = [], []
    h2d = d2h = launches = 0
    for _ in range(a.steps):
        t = time.perf_counter()
        g.align(pairs, ref, qer, aln)
        wall.append(time.perf_counter() - t)
        st = g.stats()
        kern.append(st["kernel_ms"] * 1e-3 / a.gpus)      # every GPU gets an equal share of the pairs: the summed kernel time / GPUs
        h2d += st["h2d_bytes"]; d2h += st["d2h_bytes"]; launches += st["kernel_launches"]
    clocks = sampler.stop()
    cells = st["cells"]
    # all cells the kernels computed: the forward pass plus, for pairs with a reverse pass, its rows x padded columns
    w0 = want
    byte = (pairs0["h0"] & 0x10000) != 0
    quantum = np.where(byte, 16, 8)
    did1 = w0[:, 5] >= 0
    rows1 = np.where(did1, w0[:, 1] - w0[:, 5] + 1, 0)
    cols1 = (w0[:, 2] + 1 + quantum - 1) // quantum * quantum
    computed = cells + int((rows1 * cols1).sum()) * reps
    dpx = bsw.dpx_peak(0, device=0)                                        # Ginstr/s, VIADDMNMX issue rate
    kern_s, wall_s = float(np.mean(kern)), float(np.mean(wall))
    achieved = computed / kern_s * ALU_INSTR_PER_CELL / 1e9 / a.gpus     # per GPU
    line = {
        "metric": "kswv_gcups", "value": cells / kern_s / 1e9, "unit": "GCUPS (forward-pass cell updates)", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": max(a.warmup, 1), "ms_per_step": kern_s * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int32 (8-bit class: clamped at 255 - shift)", "data": "synthetic",
        "config": {"workload": f"kswv mate rescue: {a.read_len}-bp reads x windows of 2-5 read lengths, both phases, "
                               "match 1 / mismatch 4 / gap 6+1, xtra = SUBO|START|BYTE|19",
                   "pairs": n, "l2": "inputs larger than L2 (sequences %.0f MB per step)" % ((rb + qb) * reps / 1e6)},
        "run": {"cells_forward_per_step": int(cells), "cells_computed_per_step": int(computed), "chunks": st["chunks"],
                "lanes_per_pair": st["lanes_per_pair"], "pairs_8bit_class": st["pairs8"],
                "host_ms_last_step": {k: round(st[k], 3) for k in ("host_check_ms", "host_prep_ms", "host_wait_ms", "wall_ms")}},
        "e2e": {"value": cells / wall_s / 1e9, "unit": "GCUPS", "ms_per_step": wall_s * 1e3,
                "h2d_bytes_per_step": h2d // a.steps, "d2h_bytes_per_step": d2h // a.steps,
                "api": "kswv_gpu_batch from " + ("pageable" if a.pageable else "page-locked") + " host buffers, results scattered to aln[regid]"},
        "gpu_launches": launches,
        "roofline": {"bound": "alu_int", "kernel": "kswv_phase0_kernel / kswv_phase1_kernel", "achieved": achieved, "peak": dpx,
                     "unit": "Ginstr/s (ALU-pipe thread-instructions)", "frac": achieved / dpx,
                     "instr_per_cell": ALU_INSTR_PER_CELL, "cells": "forward + reverse pass, from the results",
                     "peak_source": "measured live: VIADDMNMX issue rate, all SMs (bsw_gpu_dpx_peak)", "traffic": None},
        "parity": {"checked": 2 * base, "mismatches": mism, "against": "oracle/kswv_oracle.c (pinned to the compiled reference)"},
        "clocks": clocks,
    }
    # ---- CPU baseline on a bounded sample
    m = min(a.cpu_sample, base)
    sample = pairs0[:m].copy()
    c_sample = int((sample["len1"].astype(np.int64) * ((sample["len2"] + quantum[:m] - 1) // quantum[:m] * quantum[:m])).sum())
    nthr = os.cpu_count() or 1
    if okswv.reference_available():
        okswv.reference_batch(sample, ref0, qer0)                          # load + warm
        t = time.perf_counter()
        okswv.reference_batch(sample, ref0, qer0)
        one = c_sample / (time.perf_counter() - t) / 1e9
        ths = [threading.Thread(target=okswv.reference_batch, args=(sample.copy(), ref0, qer0)) for _ in range(nthr)]
        t = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        allc = c_sample * nthr / (time.perf_counter() - t) / 1e9
        line["cpu_baseline"] = {"value": allc, "unit": "GCUPS", "cores": nthr, "kind": "reference",
                                "one_thread": one,
                                "sample": f"first {m} pairs of the workload, one compiled AVX-512 kswv instance per host "
                                          "thread (bwa-mem2 runs one per worker, single-threaded inside), both phases"}
    else:
        t = time.perf_counter()
        okswv.oracle_batch(sample, ref0, qer0)
        line["cpu_baseline"] = {"value": c_sample / (time.perf_counter() - t) / 1e9, "unit": "GCUPS", "cores": nthr,
                                "kind": "port", "sample": f"first {m} pairs, scalar oracle on all cores (no AVX512BW here)"}
    print(json.dumps(line))
    if mism:
        sys.exit(1)
