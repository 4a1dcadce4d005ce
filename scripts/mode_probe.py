"""Developer probe for the bimodal host pass: prints the CPU the main thread started on, the NUMA view of
the guest, and the pack time of one 4 M-pair batch. Run several times back to back under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
_libc = ctypes.CDLL("libc.so.6")
if os.environ.get("NO_THP"):
    print("prctl THP disable:", _libc.prctl(41, 1, 0, 0, 0), flush=True)
cpu0 = _libc.sched_getcpu()
import numpy as np
from genarchbench_b200 import pairio, bsw
n = 4_000_000
b = pairio.generate(3, n)
cpu1 = _libc.sched_getcpu()
g = bsw.BswGpu(devices=[0])
g.batch(b.pairs, b.ref, b.qer, 100)
best = 1e9
cpu_par = 0.0
for _ in range(3):
    c0 = time.process_time(); w0 = time.perf_counter()
    g.batch(b.pairs, b.ref, b.qer, 100)
    c1 = time.process_time(); w1 = time.perf_counter()
    if g.stats()["host_pack_ms"] < best:
        best = g.stats()["host_pack_ms"]; cpu_par = (c1 - c0) / (w1 - w0)
vol = [l.split()[1] for l in open("/proc/self/status") if l.startswith(("voluntary_ctxt", "nonvoluntary_ctxt"))]
def rd(p):
    try:
        return open(p).read().strip()
    except OSError:
        return "?"
nm = ""
try:
    lines = open("/proc/self/numa_maps").read().splitlines()
    tot = {}
    for l in lines:
        for tok in l.split():
            if tok.startswith("N") and "=" in tok and tok[1].isdigit():
                k, v = tok.split("="); tot[k] = tot.get(k, 0) + int(v)
    nm = str(tot)
except OSError as e:
    nm = f"numa_maps: {e}"
roll = {l.split(":")[0]: l.split()[1] for l in open("/proc/self/smaps_rollup") if ":" in l and len(l.split()) > 1}
print(f"thp {roll.get('AnonHugePages')} kB of rss {roll.get('Rss')} kB; pid {os.getpid()} cpu-parallelism {cpu_par:.1f} ctxt {vol} threads {len(os.listdir("/proc/self/task"))} start cpu {cpu0} after-gen cpu {cpu1} pack {best:.1f} ms per {n/1e6:.0f}M  nodes online {rd('/sys/devices/system/node/online')}  pages {nm}", flush=True)
g.close()
