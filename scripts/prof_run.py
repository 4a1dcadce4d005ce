"""Profiling / A-B driver (run under gpurun, optionally under ncu): stages N pairs of a config and runs
the resident kernels a few times; BSW_CHECK=1 also checks 100k pairs against the oracle. BSW_GPU_LIB
selects a variant build (scripts/build_variant.sh). Not a test and not the bench -- numbers printed
under ncu are never bench values."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("BSW_TORCH"):
    import torch
    torch.cuda.init()
    if os.environ.get("BSW_TORCH") == "2":
        torch.zeros(1, device="cuda")
from genarchbench_b200 import pairio, bsw

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
tag = os.path.basename(os.environ.get("BSW_GPU_LIB", "default"))
g = bsw.BswGpu(devices=[0])
if os.environ.get("BSW_CHECK"):
    import oracle
    for c in (1, 2, 4):
        b = pairio.generate(c, 100000 if c != 4 else 20000, seed=77 + c)
        a = b.copy(); oracle.oracle_batch(a)
        g.batch(b.pairs, b.ref, b.qer, 100)
        print(f"[{tag}] check cfg {c}: mismatches {int((a.outputs() != b.outputs()).any(axis=1).sum())}", flush=True)
b = pairio.generate(cfg, n)
g.stage(b.pairs, b.ref, b.qer, 100)
cells = g.count_staged() if not os.environ.get("BSW_NOCOUNT") else 0
best = 1e9
rng = os.environ.get("BSW_PROFILE_RANGE")
if rng:
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    g.run_staged()
    rt.cudaDeviceSynchronize()
    rt.cudaProfilerStart()
for _ in range(reps):
    ms = g.run_staged()
    best = min(best, ms)
if rng:
    rt.cudaDeviceSynchronize()
    rt.cudaProfilerStop()
print(f"[{tag}] cfg {cfg} n {n}: best kernel {best:.3f} ms, {cells / best / 1e6:.1f} GCUPS, launches {g.stats()['kernel_launches']}, keyed pairs {g.stats().get('pairs_keyed')}", flush=True)
if os.environ.get("BSW_E2E"):
    try:
        roll = {l.split(":")[0]: l.split()[1] for l in open("/proc/self/smaps_rollup") if ":" in l and len(l.split()) > 1}
        print(f"[{tag}] AnonHugePages {roll.get('AnonHugePages')} kB of Rss {roll.get('Rss')} kB; THP:",
              open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), flush=True)
    except OSError as e:
        print("smaps:", e)
    w = b.copy()
    g.batch(w.pairs, w.ref, w.qer, 100)
    for _ in range(3):
        t0 = time.perf_counter(); g.batch(w.pairs, w.ref, w.qer, 100); dt = time.perf_counter() - t0
        st = g.stats()
        print(f"[{tag}] e2e {dt * 1e3:.1f} ms  " + " ".join(f"{k[5:-3]}={v:.1f}" for k, v in st.items() if k.startswith("host_")), flush=True)
g.close()
