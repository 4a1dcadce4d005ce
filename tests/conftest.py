"""Shared fixtures. `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 and goes through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("kswv_"))
KSWV_GOLDEN_NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("kswv_") and f.endswith(".npz"))
PARAM_ORDER = ("o_del", "e_del", "o_ins", "e_ins", "zdrop", "end_bonus", "match", "mismatch", "ambig")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _ensure_built():
    from genarchbench_b200 import pairio
    if not os.path.exists(pairio._LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "genarchbench_b200", "csrc"), "host"], check=True,
                       capture_output=True)
    if not (os.path.exists(os.path.join(ROOT, "oracle", "libbsw_oracle.so"))
            and os.path.exists(os.path.join(ROOT, "oracle", "libkswv_oracle.so"))):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)


_ensure_built()


def load_golden(name):
    """-> (PairBatch with outputs reset to -1, w, params dict, reference outputs [n,6])"""
    from genarchbench_b200 import pairio
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n = len(z["len1"])
    pairs = np.zeros(n, dtype=pairio.SEQPAIR_DTYPE)
    pairs["idr"], pairs["idq"], pairs["id"] = z["idr"], z["idq"], np.arange(n)
    pairs["len1"], pairs["len2"], pairs["h0"] = z["len1"], z["len2"], z["h0"]
    for f in ("seqid", "regid") + pairio.OUTPUT_FIELDS:
        pairs[f] = -1
    pad = np.zeros(64, np.uint8)
    b = pairio.PairBatch(pairs, np.concatenate([z["ref"], pad]), np.concatenate([z["qer"], pad]))
    params = {k: int(v) for k, v in zip(PARAM_ORDER, z["params"])}
    return b, int(z["w"]), params, z["outputs"]


def load_kswv_golden(name):
    """-> (pairs, ref, qer, params dict, reference kswr_t rows [n,7]) of a tests/golden/kswv_*.npz fixture
    (scripts/make_kswv_golden.py wrote them from the compiled, unmodified reference)."""
    from genarchbench_b200 import pairio
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n = len(z["len1"])
    pairs = np.zeros(n, dtype=pairio.SEQPAIR_DTYPE)
    pairs["idr"], pairs["idq"], pairs["id"], pairs["regid"] = z["idr"], z["idq"], np.arange(n), np.arange(n)
    pairs["len1"], pairs["len2"], pairs["h0"] = z["len1"], z["len2"], z["h0"]
    for f in ("seqid",) + pairio.OUTPUT_FIELDS:
        pairs[f] = -1
    params = dict(zip(("o_del", "e_del", "o_ins", "e_ins", "match", "mismatch"), (int(v) for v in z["params"])))
    return pairs, z["ref"], z["qer"], params, z["aln"]


def assert_same_aln(got, want, pairs=None, what=""):
    bad = np.nonzero((got != want).any(axis=1))[0]
    if len(bad):
        k = int(bad[0])
        extra = "" if pairs is None else f" (len1={pairs[k]['len1']} len2={pairs[k]['len2']} h0={hex(int(pairs[k]['h0']))})"
        raise AssertionError(f"{what}: {len(bad)} of {len(got)} pairs differ; first at {k}{extra}: got {got[k].tolist()} "
                             f"want {want[k].tolist()} [score, te, qe, score2, te2, tb, qb]")


@pytest.fixture(scope="session")
def gpu():
    """A BswGpu handle with the driver's default scoring; skips nothing: -m gpu requires the device."""
    from genarchbench_b200 import bsw
    g = bsw.BswGpu()
    yield g
    g.close()


def assert_same_outputs(got, want, batch=None, what=""):
    bad = np.nonzero((got != want).any(axis=1))[0]
    if len(bad):
        k = int(bad[0])
        extra = ""
        if batch is not None:
            p = batch.pairs[k]
            extra = f" (len1={p['len1']} len2={p['len2']} h0={p['h0']})"
        raise AssertionError(f"{what}: {len(bad)} of {len(got)} pairs differ; first at {k}{extra}: "
                             f"got {got[k].tolist()} want {want[k].tolist()} "
                             f"[score, qle, tle, gtle, gscore, max_off]")
