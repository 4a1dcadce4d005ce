"""Parity campaign for the kswv row: seeded workload families through kswv_gpu_batch on the GPU against the oracle
(oracle/kswv_oracle.c, pinned to the compiled reference). Prints one line per family and a total; exits 1 on any
mismatch. KSWV_MIN_LANES=16|32 re-runs it at the other lane-group widths.
    python scripts/kswv_campaign.py [--scale S]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    from genarchbench_b200 import kswv
    from oracle import kswv as ok
    X = ok
    FAMILIES = [
        ("151-bp mate rescue", 60000, None, dict(read_len=(151, 151))),
        ("100-151 bp", 60000, None, dict(read_len=(100, 151))),
        ("250-bp reads (16-bit class)", 20000, None, dict(read_len=(250, 250))),
        ("both classes, 200-300 bp", 20000, None, dict(read_len=(200, 300))),
        ("every strip width", 80000, None, dict(read_len=(1, 300), window=(0.3, 4.0), min_seed_len=5)),
        ("several passes", 3000, None, dict(read_len=(257, 1200))),
        ("several passes, 8-bit class", 3000, None, dict(read_len=(257, 700), p_sub=0.3,
                                                          xtra=lambda l: X.KSW_XSUBO | X.KSW_XSTART | X.KSW_XBYTE | 19)),
        ("saturating 8-bit", 20000, None, dict(read_len=(240, 330), p_sub=0.005,
                                               xtra=lambda l: X.KSW_XSUBO | X.KSW_XSTART | X.KSW_XBYTE | 19)),
        ("stop + start", 30000, None, dict(xtra=lambda l: X.KSW_XSTOP | X.KSW_XSTART | (X.KSW_XBYTE if l < 120 else 0) | 45)),
        ("no flags", 10000, None, dict(xtra=0)),
        ("start only", 10000, None, dict(xtra=X.KSW_XSTART)),
        ("ambiguous bases", 30000, None, dict(p_n=0.2)),
        ("indel-rich", 30000, None, dict(p_indel=0.05, p_sub=0.1)),
        ("a = 2, asymmetric gaps", 30000, dict(match=2, mismatch=5, o_del=4, e_del=2, o_ins=7, e_ins=1),
         dict(match=2, read_len=(60, 180), p_indel=0.03)),
        ("long windows", 4000, None, dict(read_len=(100, 151), window=(20.0, 60.0))),
        ("tiny", 60000, None, dict(read_len=(1, 24), window=(0.5, 3.0), min_seed_len=3)),
    ]
    total = bad_total = 0
    t0 = time.time()
    for k, (name, n, params, kw) in enumerate(FAMILIES):
        n = max(100, int(n * a.scale))
        # generation in Python is the slow part: a base batch, tiled
        base = min(n, 4000)
        pairs0, ref0, qer0 = ok.make_workload(base, seed=1000 + k, **kw)
        reps = (n + base - 1) // base
        rb, qb = int(pairs0["idr"][-1] + pairs0["len1"][-1]), int(pairs0["idq"][-1] + pairs0["len2"][-1])
        pairs = np.tile(pairs0, reps)
        ref = np.concatenate([ref0[:rb]] * reps + [np.zeros(64, np.uint8)])
        qer = np.concatenate([qer0[:qb]] * reps + [np.zeros(64, np.uint8)])
        for r in range(reps):
            sl = slice(r * base, (r + 1) * base)
            pairs["idr"][sl] += r * rb
            pairs["idq"][sl] += r * qb
            pairs["regid"][sl] += r * base
        rng = np.random.default_rng(k)
        if k % 2:                                   # every other family in shuffled order (gathered chunks)
            perm = rng.permutation(len(pairs))
            pairs = pairs[perm].copy()
        p = dict(ok.DEFAULT_PARAMS)
        p.update(params or {})
        g = kswv.Kswv(p["o_del"], p["e_del"], p["o_ins"], p["e_ins"], p["match"], p["mismatch"])
        got = g.align(pairs, ref, qer)
        st = g.stats()
        g.close()
        want0, _ = ok.oracle_batch(pairs0, ref0, qer0, params)
        want = np.tile(want0, (reps, 1))
        bad = int((got != want).any(axis=1).sum())
        total += len(pairs); bad_total += bad
        print(f"{name:32s} pairs {len(pairs):7d}  lanes/pair {st['lanes_per_pair']:2d}  gathered chunks {st['gathered']:2d}/{st['chunks']:2d}  "
              f"8-bit {st['pairs8']:7d}  score2>0 {(got[:, 3] > 0).mean():.2f}  tb set {(got[:, 5] >= 0).mean():.2f}  mismatches {bad}", flush=True)
    print(f"TOTAL pairs {total} mismatches {bad_total}  ({time.time() - t0:.0f} s, KSWV_MIN_LANES={os.environ.get('KSWV_MIN_LANES', 'unset')})")
    sys.exit(1 if bad_total else 0)
