"""Quick GPU bring-up check (run under gpurun): parity through the C ABI against the oracle,
integer-pipe microbenchmarks, and a first resident-kernel timing. Not a test and not the bench."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genarchbench_b200 import pairio, bsw
import oracle

def cfg(**kw):
    c = pairio.preset(1)
    for k, v in kw.items():
        setattr(c, k, v)
    return c

def main():
    out = {}
    g = bsw.BswGpu()
    cases = [("C1", 1, 100000, 100), ("C2", 2, 20000, 100), ("C4", 4, 20000, 100),
             ("short_w3", cfg(mode=1, len2_min=1, len2_max=40, h0_min=0, h0_max=30, seed=24,
                              random_frac=0.3, n_frac=0.3, small_h0_frac=0.2), 20000, 3),
             ("div_w10", cfg(mode=2, len2_min=5, len2_max=400, h0_min=1, h0_max=60, extra_max=300,
                             sub_rate=0.15, indel_rate=0.1, seed=41, random_frac=0.1, n_frac=0.2), 10000, 10)]
    for name, c, n, w in cases:
        b = pairio.generate(c, n)
        a = b.copy(); e = b.copy()
        cells = oracle.oracle_batch(a, w=w)
        t = time.time(); g.batch(e.pairs, e.ref, e.qer, w); dt = time.time() - t
        d = (a.outputs() != e.outputs()).any(axis=1)
        st = g.stats()
        print(f"{name:10s} n={n} mismatches={int(d.sum())} e2e={dt*1e3:.1f} ms kernel={st['kernel_ms']:.2f} ms "
              f"launches={st['kernel_launches']} short={st['pairs_short']} long={st['pairs_long']} "
              f"GCUPS(kernel)={cells/ (st['kernel_ms']*1e-3)/1e9:.1f}", flush=True)
        for k in np.nonzero(d)[0][:5]:
            print("    ", k, b.pairs['len1'][k], b.pairs['len2'][k], b.pairs['h0'][k], a.outputs()[k], e.outputs()[k])
        out[name] = int(d.sum())
    names = ["VIADDMNMX.S16x2.RELU", "VIMNMX3.S16x2", "VIADD.16x2", "LOP3", "PRMT", "IMAD", "SHF", "IMAD.HI", "VIADDMNMX+IMAD"]
    for w, nm in enumerate(names):
        v = bsw.dpx_peak(w)
        print(f"peak[{w}] {nm:22s} {v:9.1f} Ginstr/s", flush=True)
        out["peak_" + nm] = v
    # resident timing, 1M C1 pairs
    b = pairio.generate(1, 1000000)
    a = b.copy(); cells = oracle.oracle_batch(a)
    g.stage(b.pairs, b.ref, b.qer, 100)
    for it in range(4):
        ms = g.run_staged()
        print(f"staged C1 1M: kernel {ms:.2f} ms  GCUPS {cells/(ms*1e-3)/1e9:.1f}  pairs/s {1e6/(ms*1e-3)/1e6:.1f} M", flush=True)
    g.fetch_staged(b.pairs)
    print("staged parity mismatches", int((a.outputs() != b.outputs()).any(axis=1).sum()))
    t = time.time(); g.batch(b.pairs, b.ref, b.qer, 100); dt = time.time() - t
    print("e2e 1M C1:", dt * 1e3, "ms", g.stats())
    out["cells_1M"] = cells
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)

if __name__ == "__main__":
    main()
