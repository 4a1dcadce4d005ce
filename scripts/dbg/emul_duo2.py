"""Developer check: extend_duo2 (host emulation) against the oracle, unsorted and sorted neighbours."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle
from genarchbench_b200 import pairio
ROOT = os.path.join(os.path.dirname(__file__), "..", "..")
L = C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libbsw_emul.so"))
L.bsw_emul_batch_duo2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]

def run(b, w, params=None, key=0):
    L.bsw_emul_batch_duo2(oracle._params_array(params), b.pairs.ctypes.data, b.ref.ctypes.data, b.qer.ctypes.data, len(b), w, key)
    return b.outputs()

def check(name, b, w, params=None, key=0, sort=False):
    if sort:
        o = np.lexsort((b.pairs["h0"], b.pairs["len1"], b.pairs["len2"]))
        b = pairio.PairBatch(b.pairs[o].copy(), b.ref, b.qer)
    a = b.copy()
    oracle.oracle_batch(a, w=w, params=params)
    got = run(b, w, params, key)
    bad = np.nonzero((got != a.outputs()).any(axis=1))[0]
    keyed = int((b.pairs["seqid"] == -1).sum())
    print(f"{name:40s} w={w:4d} key={key} sort={int(sort)} n={len(b)} keyed={keyed} mismatches={len(bad)}")
    for k in bad[:3]:
        p = b.pairs[k]
        print("   ", k, "len1", p["len1"], "len2", p["len2"], "h0", p["h0"], "got", got[k].tolist(), "want", a.outputs()[k].tolist(),
              "| partner", b.pairs[k ^ 1]["len1"], b.pairs[k ^ 1]["len2"], b.pairs[k ^ 1]["h0"])
    return len(bad)

tot = 0
for key in (0, 1):
    for sort in (False, True):
        for w in (1, 2, 5, 17, 100):
            c = pairio.preset(4)
            c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 200, 0, 60, 0.3, 0.2
            tot += check("mixed 1..200", pairio.generate(c, 6001, seed=300 + w), w, key=key, sort=sort)
        tot += check("config 1", pairio.generate(1, 20000, seed=7), 100, key=key, sort=sort)
        tot += check("config 2", pairio.generate(2, 3000, seed=8), 100, key=key, sort=sort)
        for params in (dict(o_del=5, e_del=2, o_ins=7, e_ins=1, zdrop=40, end_bonus=9, match=2, mismatch=3, ambig=-1),
                       dict(o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=0, end_bonus=5, match=1, mismatch=4, ambig=-1)):
            c = pairio.preset(4)
            c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 150, 0, 100, 0.2, 0.2
            tot += check("nondefault " + str(params["match"]), pairio.generate(c, 5000, seed=11), 30, params, key=key, sort=sort)
print("TOTAL MISMATCHES", tot)
sys.exit(1 if tot else 0)
