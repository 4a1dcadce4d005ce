"""Generates tests/golden/*.npz from the UNMODIFIED reference kernel (oracle/_ref, compiled from
/root/reference by oracle/Makefile). Run it in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Each fixture holds the inputs (len1, len2, h0, concatenated base codes), the band width / scoring
parameters and the six outputs of the reference's getScores16 (bandedSWA.cpp:2679) for every pair.
The GPU box has no /root/reference; tests there read these files only.

The reference does not ship golden vectors for this path (its only pinned artefact is a score-only
file inside a 90 GB dataset, benchmarks/bsw/scripts/regression_small.sh:92), so these are outputs of
the reference itself, produced here.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from genarchbench_b200 import pairio  # noqa: E402
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cfg(**kw):
    c = pairio.preset(1)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def edge_cases() -> pairio.PairBatch:
    rng = np.random.default_rng(5)
    A = lambda n: rng.integers(0, 4, n).astype(np.uint8)  # noqa: E731
    items = []
    q = A(40)
    items.append((q.copy(), q.copy(), 30))                       # perfect match, square
    items.append((np.concatenate([q, A(35)]), q.copy(), 30))     # perfect match + tail
    items.append((A(60), A(40), 25))                             # unrelated
    items.append((q.copy(), q.copy(), 0))                        # h0 = 0: first row dies
    items.append((q.copy(), q.copy(), 1))
    items.append((A(1), A(1), 10))                               # 1 x 1
    items.append((A(1), A(50), 10))                              # single target row
    items.append((A(50), A(1), 10))                              # single query column
    t = q.copy(); t[5] = 4
    items.append((t, q.copy(), 40))                              # N in target
    q2 = q.copy(); q2[7] = 4
    items.append((q.copy(), q2, 40))                             # N in query
    t3 = q.copy(); t3[9] = 4; q3 = q.copy(); q3[9] = 4
    items.append((t3, q3, 40))                                   # N against N scores ambig, not match
    items.append((np.full(30, 4, np.uint8), np.full(20, 4, np.uint8), 50))  # all N
    qq = A(120)
    items.append((np.concatenate([qq[:60], A(3), qq[60:], A(50)]), qq, 60))      # 3-base insertion
    items.append((np.concatenate([qq[:50], qq[58:], A(60)]), qq, 60))            # 8-base deletion
    tt = qq.copy(); tt[20:60] = (tt[20:60] + 1) % 4
    items.append((np.concatenate([tt, A(100)]), qq, 100))        # long mismatch run: z-drop
    items.append((np.concatenate([qq, A(100)]), qq, 127))
    ql = A(300)
    items.append((np.concatenate([ql, A(200)]), ql, 150))        # 16-bit scores (> 127 and > 255)
    items.append((np.concatenate([ql[:150], A(20), ql[150:], A(150)]), ql, 150))
    items.append((A(520), ql, 19))
    qxl = A(700)
    items.append((np.concatenate([qxl, A(100)]), qxl, 80))       # beyond the shared-memory bins
    items.append((np.zeros(70, np.uint8), np.zeros(33, np.uint8), 20))  # homopolymer
    items.append((np.tile(np.array([0, 1], np.uint8), 40), np.tile(np.array([0, 1], np.uint8), 25), 33))
    return pairio.from_sequences(items)


FIXTURES = {
    # name: (batch factory, w, params)
    "c1_small": (lambda: pairio.generate(1, 3000, seed=7101), 100, {}),
    "c2_16bit": (lambda: pairio.generate(2, 600, seed=7102), 100, {}),
    "c4_skewed": (lambda: pairio.generate(4, 400, seed=7104), 100, {}),
    "edge": (edge_cases, 100, {}),
    "ambig_heavy": (lambda: pairio.generate(cfg(seed=7105, n_frac=1.0), 1500), 100, {}),
    "unrelated": (lambda: pairio.generate(cfg(seed=7106, random_frac=1.0, small_h0_frac=0.2), 1500), 100, {}),
    "w30": (lambda: pairio.generate(cfg(seed=7107, mode=2, len2_min=5, len2_max=400, h0_min=1, h0_max=60,
                                        extra_max=300, sub_rate=0.15, indel_rate=0.1), 1000), 30, {}),
    "gape2_zdrop30": (lambda: pairio.generate(cfg(seed=7108, sub_rate=0.1, indel_rate=0.05), 1500), 100,
                      dict(e_del=2, e_ins=2, zdrop=30)),
    "asym_gaps": (lambda: pairio.generate(cfg(seed=7109, sub_rate=0.1, indel_rate=0.05), 1500), 100,
                  dict(o_del=5, e_del=2, o_ins=7, e_ins=1)),
    "match2_mismatch3": (lambda: pairio.generate(cfg(seed=7110, sub_rate=0.1, indel_rate=0.05), 1500), 100,
                         dict(match=2, mismatch=3)),
}


def main():
    assert oracle.reference_available(), "build oracle/_ref first: make -C oracle ref"
    for name, (make, w, params) in FIXTURES.items():
        b = make()
        outs = {}
        for isa in oracle.reference_isas():
            r = b.copy()
            oracle.reference_batch(r, w=w, params=params, isa=isa)
            outs[isa] = r.outputs()
        base = outs[oracle.reference_isas()[0]]
        for isa, o in outs.items():
            assert (o == base).all(), f"{name}: reference ISA builds disagree ({isa})"
        p = dict(oracle.DEFAULT_PARAMS); p.update(params)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            len1=b.pairs["len1"], len2=b.pairs["len2"], h0=b.pairs["h0"],
            idr=b.pairs["idr"], idq=b.pairs["idq"],
            ref=b.ref[: int((b.pairs["idr"] + b.pairs["len1"]).max())],
            qer=b.qer[: int((b.pairs["idq"] + b.pairs["len2"]).max())],
            w=np.int32(w), params=np.array([p[k] for k in oracle._PARAM_ORDER], np.int32),
            outputs=base.astype(np.int32))
        print(name, len(b), "pairs, isas", list(outs))


if __name__ == "__main__":
    main()
