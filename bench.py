#!/usr/bin/env python
"""bench.py -- bsw (banded Smith-Waterman seed extension) throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch: BASELINE.json config 3 ("bsw large-shape single
GPU": 10 M synthetic 151-bp read / ref-window pairs, w=100, default scoring) PER GPU (weak scaling;
N=8 is config 5, 80 M pairs). Rank r generates its own shard (seed 1003 + 7919 r); there is no
data-path collective (pairs are independent) -- torch.distributed only provides the barrier and the
max-over-ranks / sum-over-ranks reductions.

  value : GCUPS = DP cells the reference's scalar loop visits (counted on the device by
          the COUNT kernel, checked against the oracle in tests) / device time (CUDA events on the
          launching stream, inputs resident in HBM), whole job, max time over ranks.
  e2e   : the same metric through the drop-in call bsw_gpu_batch with HOST buffers: binning, 2-bit
          packing into pinned memory, H2D, kernels, D2H and the scatter into SeqPair all inside the
          timed region (wall clock around the call, max over ranks).
  roofline : integer/DPX pipe (this path is neither HBM- nor tensor-bound):
          achieved = cells/s * 2.5 packed-s16x2 instructions, peak = VIADDMNMX.S16x2.RELU issue rate
          measured live on this GPU (bsw_gpu_dpx_peak). An `hbm` sanity entry uses MEASURED_PEAKS.json.
  cpu_baseline : the UNMODIFIED reference getScores16 (oracle/_ref, OpenMP over 512-pair batches as in
          main_banded.cpp:338-350) on the box's host cores, bounded sample, rank 0 at N=1.
  --impl reference : only that CPU arm, same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RANK = int(os.environ.get("RANK", 0))
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", 0))
WORLD = int(os.environ.get("WORLD_SIZE", 1))
NCORES = os.cpu_count() or 1
# host packing threads: share the box's cores between the ranks (set before OpenMP loads).
# torchrun exports OMP_NUM_THREADS=1 to every rank by default, which is not a user's choice here:
# BSW_HOST_THREADS is the explicit override, otherwise each rank gets cores / world.
if "BSW_HOST_THREADS" in os.environ or WORLD > 1 or "OMP_NUM_THREADS" not in os.environ:
    os.environ["OMP_NUM_THREADS"] = os.environ.get("BSW_HOST_THREADS", str(max(1, NCORES // max(WORLD, 1))))
# rank 0 prints ONE JSON line on stdout: everything else that libraries write to fd 1 (NCCL's version
# banner, for one) is sent to stderr; the line itself goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
# same placement policy as the reference's regression scripts (bsw/scripts/regression_small.sh:52):
# threads stay on their cores (without it the host pass is bimodal, 29 vs 54 ms per 10 M pairs).
# Under torchrun the ranks share the box, so binding is left to the launcher there.
if WORLD == 1:
    os.environ.setdefault("OMP_PROC_BIND", "true")
    os.environ.setdefault("OMP_PLACES", "cores")

import numpy as np  # noqa: E402

SCORING = "w=100, match 1 / mismatch 4 / gap 6+1 / zdrop 100 / end bonus 5"
WORKLOADS = {
    1: ("bsw config 1 (small-shape): synthetic 151-bp read / ref-window extension pairs", 100_000),
    2: ("bsw config 2 (16-bit path): 250-300-base queries, scores beyond int8", 100_000),
    3: ("bsw config 3 (large-shape): synthetic 151-bp read / ref-window extension pairs", 10_000_000),
    4: ("bsw config 4 (skewed lengths): 30-1000-base queries, log-uniform", 5_000_000),
    5: ("bsw config 5 (scaling): config-3 shape, 80 M pairs over 8 GPUs", 10_000_000),
}
WORKLOAD = WORKLOADS[3][0] + ", " + SCORING
INSTR_PER_CELL = 2.5          # 5 packed-s16x2 DPX instructions per 2 cells
BYTES_PER_PAIR_FMT = "ceil(len1/4)+ceil(len2/4)+12+24"
# ncu figures of the dominant kernel for the CURRENT build, written by scripts/ncu_counters.py from a committed
# capture (the file names its source CSVs); absent -> `traffic` and `alu_instr_per_cell` are null
NCU_COUNTERS = os.path.join(ROOT, "profiles", "r2_ncu_counters.json")
# rough packed bytes per pair of each workload: decides, identically in both arms, whether a step's inputs
# exceed the 126 MB L2 or the L2 is flushed between timed steps
EST_BYTES_PER_PAIR = {1: 85, 2: 200, 3: 85, 4: 250, 5: 85}


def l2_policy(workload: int, pairs: int):
    flush = EST_BYTES_PER_PAIR[workload] * pairs < 2 * 126_000_000
    return flush, ("L2 flushed between timed steps (256 MB written)" if flush
                   else "inputs larger than L2 (126 MB): no flush needed")


def run_config(args, world: int) -> dict:
    """The `config` object: identical in both arms (`--impl ours` / `--impl reference`) for the same flags."""
    n_dev = max(world, args.inproc_gpus, 1)
    return {"workload": WORKLOAD, "pairs_per_gpu": args.pairs, "global_pairs": args.pairs * n_dev, "w": 100,
            "parallelism": (f"pair-sharded x{n_dev}, no collective" +
                            (" (one process, bsw_gpu_init(n_gpus))" if args.inproc_gpus > 1 else "")),
            "l2": l2_policy(args.workload, args.pairs)[1]}


def ncu_counters() -> dict:
    try:
        return json.load(open(NCU_COUNTERS))
    except (OSError, ValueError):
        return {}


def parity_sample(work, rank: int, n_check: int):
    """Compares a random sample of the e2e outputs with the oracle (checker only, outside every timed
    region). -> (pairs checked, mismatching pairs)"""
    import oracle
    n = len(work)
    rng = np.random.default_rng(4242 + rank)
    idx = np.sort(rng.choice(n, size=min(n, n_check), replace=False))
    from genarchbench_b200 import pairio
    want = pairio.PairBatch(work.pairs[idx].copy(), work.ref, work.qer)
    got = want.outputs()
    oracle.oracle_batch(want)
    return len(idx), int((got != want.outputs()).any(axis=1).sum())


from genarchbench_b200.benchutil import ClockSampler  # noqa: E402


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def algorithmic_bytes(pairs: np.ndarray) -> int:
    """per pair ceil(len1/4)+ceil(len2/4) packed bases + 12 B of lengths/h0 + 24 B out."""
    l1 = pairs["len1"].astype(np.int64)
    l2 = pairs["len2"].astype(np.int64)
    return int(((l1 + 3) // 4 + (l2 + 3) // 4 + 36).sum())


def reference_arm(args, batch, cells_fn) -> dict:
    """Times the reference CPU implementation of the path on a bounded sample, all host threads."""
    import oracle
    n = min(len(batch), args.cpu_sample)
    sample = batch.slice(0, n)
    cells = cells_fn(sample)
    kind = "reference" if oracle.reference_available() else "port"
    isa = oracle.reference_isas()[0] if kind == "reference" else None
    threads = NCORES

    def one():
        if kind == "reference":
            return oracle.reference_batch(sample, nthreads=threads, isa=isa)
        t0 = time.perf_counter()
        oracle.oracle_batch(sample, nthreads=threads)
        return time.perf_counter() - t0

    for _ in range(args.warmup_cpu):
        one()
    secs = [one() for _ in range(args.steps_cpu)]
    t = float(np.mean(secs))
    return {"value": cells / t / 1e9, "unit": "GCUPS", "cores": threads, "kind": kind,
            "sample": f"first {n} pairs of the rank-0 workload, {args.steps_cpu} passes after "
                      f"{args.warmup_cpu} warm-up, ROI = OpenMP loop over 512-pair getScores16 batches"
                      + (f" ({isa} build of the reference)" if isa else " (oracle port)"),
            "pairs_per_s": n / t, "ms_per_pass": t * 1e3, "cells": cells, "pairs": n}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the workload's size; config 3: 10 M)")
    ap.add_argument("--workload", type=int, default=3, choices=sorted(WORKLOADS),
                    help="BASELINE.json config to run (default 3, the one the metric is quoted on)")
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--steps-cpu", type=int, default=3)
    ap.add_argument("--warmup-cpu", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    ap.add_argument("--inproc-gpus", type=int, default=0,
                    help="ONE process driving N GPUs through bsw_gpu_init(n_gpus=N) (the library's own split; "
                         "not under torchrun). The batch holds N x --pairs pairs")
    ap.add_argument("--parity-sample", type=int, default=50_000,
                    help="pairs per rank of the e2e outputs compared with the oracle after the timed regions")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rule: at least 3 warm-up steps
    global WORKLOAD
    WORKLOAD = WORKLOADS[args.workload][0] + ", " + SCORING
    if args.pairs <= 0:
        args.pairs = WORKLOADS[args.workload][1]
    gen_cfg = 3 if args.workload == 5 else args.workload
    gen_seed = {1: 1001, 2: 1002, 3: 1003, 4: 1004, 5: 1005}[args.workload]

    from genarchbench_b200 import pairio, dist as bdist

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if RANK != 0:
            return
        import oracle
        batch = pairio.generate(gen_cfg, min(args.pairs, args.cpu_sample), seed=bdist.shard_seed(gen_seed, 0))
        args.steps_cpu, args.warmup_cpu = max(1, args.steps), max(0, args.warmup)
        cb = reference_arm(args, batch, lambda b: oracle.oracle_batch(b.copy()))
        line = {"impl": "reference", "metric": "bsw_gcups", "value": cb["value"], "unit": "GCUPS",
                "n_gpus": args.gpus, "steps": args.steps_cpu, "warmup": args.warmup_cpu,
                "ms_per_step": cb["ms_per_pass"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int16", "data": "synthetic",
                "config": run_config(args, max(args.gpus, 1)),
                "sample_pairs_per_step": cb["pairs"],
                "pairs_per_s": cb["pairs_per_s"],
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    from genarchbench_b200 import bsw
    rank, local_rank, world = bdist.init_process_group()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)

    def sync_all():
        torch.cuda.synchronize()
        bdist.barrier()
        torch.cuda.synchronize()

    if args.inproc_gpus > 1:
        if world > 1:
            raise SystemExit("--inproc-gpus is a single-process mode: do not run it under torchrun")
        # one process, N GPUs: the shards the N ranks of a torchrun job would own, concatenated
        parts = [pairio.generate(gen_cfg, args.pairs, seed=bdist.shard_seed(gen_seed, r)) for r in range(args.inproc_gpus)]
        batch = pairio.concat(parts)
        del parts
        g = bsw.BswGpu(n_gpus=args.inproc_gpus)
    else:
        batch = pairio.generate(gen_cfg, args.pairs, seed=bdist.shard_seed(gen_seed, rank))
        g = bsw.BswGpu(devices=[local_rank])
    n_dev = max(world, args.inproc_gpus, 1)
    g.stage(batch.pairs, batch.ref, batch.qer, 100)
    cells = g.count_staged()                      # unit of work, outside any timed region
    dpx_peak = bsw.dpx_peak(0, device=local_rank)  # Ginstr/s, measured live on this GPU
    trip_peak = bsw.dpx_peak(9, device=local_rank)  # G cells/s of the inner-loop arithmetic alone (registers only)

    # ---- value: device-resident kernel throughput
    # timing rule: inputs larger than L2, or L2 flushed between iterations (small workloads)
    alg_bytes = algorithmic_bytes(batch.pairs)
    flush_bufs = []
    if l2_policy(args.workload, args.pairs)[0]:
        flush_bufs = [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{d}")
                      for d in (range(args.inproc_gpus) if args.inproc_gpus > 1 else [local_rank])]

    def flush_l2():
        for fb in flush_bufs:
            fb.zero_()
        if flush_bufs:
            for fb in flush_bufs:
                torch.cuda.synchronize(fb.device)

    for _ in range(args.warmup):
        g.run_staged()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        flush_l2()                                # outside the event-timed region of run_staged
        dev_ms += g.run_staged()                  # CUDA events on the launching stream
        launches += g.stats()["kernel_launches"]
    sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    (sum_cells, sum_pairs, sum_launch), (max_dev_ms, max_wall_ms) = bdist.reduce_stats(
        [cells, len(batch), launches], [dev_ms, wall_ms])
    ms_per_step = max_dev_ms / args.steps
    gcups = sum_cells / (ms_per_step * 1e-3) / 1e9
    pairs_per_s = sum_pairs / (ms_per_step * 1e-3)

    # ---- e2e: through bsw_gpu_batch with host buffers
    e2e_steps = args.e2e_steps or args.steps
    work = batch.copy()
    g.batch(work.pairs, work.ref, work.qer, 100)  # warm the pinned rings
    sync_all()
    t0 = time.perf_counter()
    h2d = d2h = e2e_launch = 0
    for _ in range(e2e_steps):
        g.batch(work.pairs, work.ref, work.qer, 100)
        st = g.stats()
        h2d += st["h2d_bytes"]; d2h += st["d2h_bytes"]; e2e_launch += st["kernel_launches"]
    last = dict(st)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    (sum_h2d, sum_d2h), (max_e2e_ms,) = bdist.reduce_stats([h2d, d2h], [e2e_ms])
    e2e_step_s = max_e2e_ms / e2e_steps * 1e-3
    checksum_ok = bool((work.outputs() != -1).any())
    # ---- e2e_packed: the ingest path -- what a packed pair file holds, resident in page-locked host memory, through
    # bsw_gpu_batch_packed; results into a page-locked array of 16-byte records (H2D and D2H inside the timed region)
    rec, pdata = pairio.pack(batch, bsw.host_alloc)
    pres = bsw.host_alloc(len(batch) * pairio.RESULT_DTYPE.itemsize).view(pairio.RESULT_DTYPE)
    g.batch_packed(rec, pdata, 100, pres)            # warm
    sync_all()
    t0 = time.perf_counter()
    ph2d = pd2h = 0
    for _ in range(e2e_steps):
        g.batch_packed(rec, pdata, 100, pres)
        stp = g.stats()
        ph2d += stp["h2d_bytes"]; pd2h += stp["d2h_bytes"]
    last_p = dict(stp)
    sync_all()
    pk_ms = (time.perf_counter() - t0) * 1e3
    (sum_ph2d, sum_pd2h), (max_pk_ms,) = bdist.reduce_stats([ph2d, pd2h], [pk_ms])
    pk_step_s = max_pk_ms / e2e_steps * 1e-3
    packed_same = bool((bsw.results_to_outputs(pres[:200_000]) == work.outputs()[:200_000]).all())

    # parity of what the timed e2e calls wrote, on a random sample per rank, against the oracle
    n_chk, n_bad = parity_sample(work, rank, args.parity_sample) if args.parity_sample > 0 else (0, 0)
    (sum_chk, sum_bad), _ = bdist.reduce_stats([n_chk, n_bad], [0.0])

    if rank != 0:
        g.close()
        bdist.shutdown()
        return

    peaks = measured_peaks()
    ncu = ncu_counters()
    achieved_instr = cells / (dev_ms / args.steps * 1e-3) * INSTR_PER_CELL / 1e9   # this rank's GPU
    roofline = {
        "bound": "dpx_int", "kernel": ncu.get("kernel", "the thread-per-pair s16x2 DPX kernel (see profiles/)"),
        "achieved": achieved_instr / max(args.inproc_gpus, 1), "peak": dpx_peak, "unit": "Ginstr/s (packed s16x2 thread-instructions)",
        "frac": achieved_instr / max(args.inproc_gpus, 1) / dpx_peak, "instr_per_cell": INSTR_PER_CELL,
        "peak_source": "measured live: VIADDMNMX.S16x2.RELU issue rate, all SMs (bsw_gpu_dpx_peak)",
        # second, tighter ceiling: the kernel's own inner-loop arithmetic (8 cells per trip) run on registers
        # only at full occupancy -- no shared memory, row bookkeeping or divergence (bsw_gpu_dpx_peak(9))
        "inner_loop_ceiling_gcups": trip_peak,
        "frac_of_inner_loop_ceiling": (cells / max(args.inproc_gpus, 1) / (dev_ms / args.steps * 1e-3) / 1e9) / trip_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of the DP kernels per pair, from the committed ncu capture
        # of this build that profiles/r2_ncu_counters.json names, scaled to the pairs of this GPU's step
        "traffic": (int(ncu["dram_bytes_per_pair"] * len(batch) / max(args.inproc_gpus, 1))
                    if ncu.get("dram_bytes_per_pair") and args.workload in (3, 5) else None),
        "traffic_unit": "bytes per step (DP kernels, one GPU)",
        "traffic_source": ncu.get("source", "no ncu capture of this build committed: null"),
        # ALU-pipe thread-instructions per visited cell (smsp__inst_executed_pipe_alu x 32 / cells), same capture
        "alu_instr_per_cell": ncu.get("alu_instr_per_cell") if args.workload in (3, 5) else None,
        "hbm": {"algorithmic_bytes_per_step": alg_bytes, "bytes_per_pair": BYTES_PER_PAIR_FMT,
                "achieved_gbs": alg_bytes / (dev_ms / args.steps * 1e-3) / 1e9,
                "frac": (alg_bytes / (dev_ms / args.steps * 1e-3) / 1e9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                "peak_gbs": peaks.get("hbm_gbs"), "peak_source": "MEASURED_PEAKS.json (of measured)"
                if peaks.get("hbm_gbs") else "unavailable"},
    }
    cpu_baseline = None
    if world == 1 and args.inproc_gpus <= 1:
        g2 = bsw.BswGpu(devices=[local_rank])

        def count(b):
            g2.stage(b.pairs, b.ref, b.qer, 100)
            return g2.count_staged()
        cb = reference_arm(args, batch, count)
        g2.close()
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "pairs_per_s")}

    line = {
        "metric": "bsw_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": n_dev, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": run_config(args, world),
        "run": {"global_pairs": int(sum_pairs), "cells_visited_per_step": int(sum_cells),
                "cells_rect_rank0": batch.cells_rect(), "algorithmic_input_mb_rank0": round(alg_bytes / 1e6, 1)},
        "pairs_per_s": pairs_per_s,
        "wall_ms_per_step": max_wall_ms / args.steps,
        "clocks": clocks,
        "e2e": {"value": sum_cells / e2e_step_s / 1e9, "unit": "GCUPS",
                "pairs_per_s": sum_pairs / e2e_step_s, "ms_per_step": e2e_step_s * 1e3, "steps": e2e_steps,
                "h2d_bytes_per_step": int(sum_h2d / e2e_steps), "d2h_bytes_per_step": int(sum_d2h / e2e_steps),
                "host_ms": {k[5:-3]: round(last[k], 3) for k in last if k.startswith("host_")},
                "kernel_ms": last["kernel_ms"], "host_threads": int(os.environ["OMP_NUM_THREADS"]),
                "api": "bsw_gpu_batch(SeqPair*, ref, qer, n, w) from host buffers", "results_written": checksum_ok},
        # the same metric on the ingest path: packed pair data (what a BSWPAIR1 file holds) in page-locked host memory
        # -> bsw_gpu_batch_packed -> 16-byte result records; no byte-per-base buffers or SeqPair records on this path
        "e2e_packed": {"value": sum_cells / pk_step_s / 1e9, "unit": "GCUPS", "pairs_per_s": sum_pairs / pk_step_s,
                       "ms_per_step": pk_step_s * 1e3, "steps": e2e_steps,
                       "h2d_bytes_per_step": int(sum_ph2d / e2e_steps), "d2h_bytes_per_step": int(sum_pd2h / e2e_steps),
                       "host_ms": {k[5:-3]: round(last_p[k], 3) for k in last_p if k.startswith("host_")},
                       "kernel_ms": last_p["kernel_ms"],
                       "api": "bsw_gpu_batch_packed(rec, data, n, w, out) from page-locked host memory",
                       "same_results_as_e2e": packed_same},
        # e2e outputs of every rank against the oracle on a random sample (checked after the timed regions)
        "parity": {"pairs_checked": int(sum_chk), "mismatches": int(sum_bad), "against": "oracle/bsw_oracle.c",
                   "fields": "score, qle, tle, gtle, gscore, max_off"},
        "gpu_launches": int(sum_launch),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    emit(line)
    g.close()
    bdist.shutdown()
    if sum_bad:
        raise SystemExit(f"PARITY FAILURE: {int(sum_bad)} of {int(sum_chk)} sampled pairs differ from the oracle")


if __name__ == "__main__":
    main()
