"""Large randomized parity campaign (run under gpurun): the CUDA path through the C ABI against the
oracle on all six outputs, over configurations, band widths and scoring parameters, through bsw_gpu_batch and through
bsw_gpu_batch_packed. Prints one line per case and a total; exits non-zero on any mismatch."""
import itertools, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genarchbench_b200 import pairio, bsw
import oracle

cases = []
for cfg, n in ((1, 1_000_000), (2, 300_000), (3, 2_000_000), (4, 300_000)):
    for w in (100,):
        cases.append((f"config {cfg}", cfg, n, w, None, {}))
for w in (0, 1, 2, 5, 13, 50, 150, 400):
    cases.append((f"mixed lengths w={w}", 4, 150_000, w, None, dict(len2_min=1, len2_max=1200, h0_min=0, h0_max=250, n_frac=0.25, random_frac=0.15)))
scorings = [dict(o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5, match=1, mismatch=4, ambig=-1),
            dict(o_del=5, e_del=2, o_ins=7, e_ins=1, zdrop=40, end_bonus=9, match=2, mismatch=3, ambig=-1),
            dict(o_del=4, e_del=1, o_ins=4, e_ins=2, zdrop=0, end_bonus=0, match=1, mismatch=1, ambig=-1),
            dict(o_del=0, e_del=1, o_ins=0, e_ins=1, zdrop=30, end_bonus=3, match=3, mismatch=6, ambig=-1),
            dict(o_del=12, e_del=3, o_ins=10, e_ins=4, zdrop=200, end_bonus=20, match=5, mismatch=9, ambig=-1)]
for i, sc in enumerate(scorings[1:]):
    for w in (7, 100, 300):
        cases.append((f"scoring {i + 1} w={w}", 4, 100_000, w, sc, dict(len2_min=1, len2_max=900, h0_min=0, h0_max=200, n_frac=0.2, random_frac=0.1)))
# keyed row argmax at the limits of its 16-bit key: queries <= 60 bases (5 index bits, scores < 2048)
cases.append(("keyed limit below", 4, 200_000, 100, None, dict(len2_min=1, len2_max=60, h0_min=0, h0_max=1985, n_frac=0.2, random_frac=0.1)))
cases.append(("keyed limit above", 4, 200_000, 100, None, dict(len2_min=1, len2_max=60, h0_min=0, h0_max=1995, n_frac=0.2, random_frac=0.1)))
cases.append(("keyed 7-bit index", 4, 200_000, 30, None, dict(len2_min=100, len2_max=250, h0_min=0, h0_max=260, n_frac=0.2, random_frac=0.1)))
cases.append(("large h0", 4, 100_000, 100, None, dict(len2_min=10, len2_max=800, h0_min=15000, h0_max=31000, n_frac=0.1)))

total = bad_total = 0
t00 = time.time()
for name, cfg, n, w, sc, over in cases:
    c = pairio.preset(cfg)
    for k, v in over.items():
        setattr(c, k, v)
    b = pairio.generate(c, n, seed=zlib.crc32(name.encode()) % 100000)
    a = b.copy()
    oracle.oracle_batch(a, w=w, params=sc)
    with bsw.BswGpu(**(sc or {})) as g:
        g.batch(b.pairs, b.ref, b.qer, w)
        st = g.stats()
        rec, data = pairio.pack(b)                  # the same pairs through the packed entry point
        res = g.batch_packed(rec, data, w)
    bad = int((a.outputs() != b.outputs()).any(axis=1).sum())
    bad += int((a.outputs() != bsw.results_to_outputs(res)).any(axis=1).sum())
    total += n; bad_total += bad
    print(f"{name:28s} n={n:8d} short={st['pairs_short']:8d} long={st['pairs_long']:7d} keyed={st['pairs_keyed']:8d} mismatches={bad}", flush=True)
print(f"TOTAL {total} pairs, {bad_total} mismatches, {time.time() - t00:.0f} s")
sys.exit(1 if bad_total else 0)
