"""TEST INFRASTRUCTURE ONLY -- checkers for the kswv path (bwa-mem2's batched mate-rescue Smith-Waterman).

* ``oracle_batch``    -- oracle/kswv_oracle.c, the plain-C restatement (file:line in its header).
* ``reference_batch`` -- oracle/_ref/libkswv_ref_avx512.so: the UNMODIFIED reference class compiled from
                         /root/reference by oracle/Makefile, driven like mem_sam_pe_batch
                         (bwamem_pair.cpp:634-704). The class only has an AVX512BW body.
* ``make_workload``   -- seeded synthetic mate-rescue batches shaped like mem_matesw_batch_pre's
                         (bwamem_pair.cpp:930-1090): a read against a reference window a few times its length.

Both checkers return an int32 array [n, 7] = kswr_t {score, te, qe, score2, te2, tb, qb} indexed by regid.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _HERE, _cpu_flags, build

KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000
DEFAULT_PARAMS = dict(o_del=6, e_del=1, o_ins=6, e_ins=1, match=1, mismatch=4)   # bwa-mem2 mem_opt_init
_ORDER = ("o_del", "e_del", "o_ins", "e_ins", "match", "mismatch")
FIELDS = ("score", "te", "qe", "score2", "te2", "tb", "qb")


def _params(params):
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    return (C.c_int32 * 6)(*[int(p[k]) for k in _ORDER])


_lib = None


def _oracle():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libkswv_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.kswv_oracle_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_int, C.POINTER(C.c_int64)]
        L.kswv_oracle_batch.restype = C.c_int
        assert L.kswv_oracle_sizeof_seqpair() == 72
        _lib = L
    return _lib


def oracle_batch(pairs, ref, qer, params=None, nthreads=0):
    L = _oracle()
    n = len(pairs)
    aln = np.full((n, 7), -1, np.int32)
    cells = C.c_int64(0)
    ref, qer = np.ascontiguousarray(ref), np.ascontiguousarray(qer)
    rc = L.kswv_oracle_batch(_params(params), pairs.ctypes.data, ref.ctypes.data, qer.ctypes.data, n,
                             aln.ctypes.data, nthreads or (os.cpu_count() or 1), C.byref(cells))
    if rc != 0:
        raise RuntimeError("kswv_oracle_batch: out of memory or a pair outside the domain")
    return aln, int(cells.value)


_REF_PATH = os.path.join(_HERE, "_ref", "libkswv_ref_avx512.so")
_ref = None


def reference_available() -> bool:
    return os.path.exists(_REF_PATH) and "avx512bw" in _cpu_flags()


def reference_batch(pairs, ref, qer, params=None):
    global _ref
    if _ref is None:
        L = C.CDLL(_REF_PATH)
        L.ref_kswv_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_int32, C.c_void_p]
        L.ref_kswv_batch.restype = C.c_int
        assert L.ref_kswv_sizeof_seqpair() == 72 and L.ref_kswv_sizeof_kswr() == 28
        _ref = L
    n = len(pairs)
    aln = np.full((n, 7), -1, np.int32)
    ref, qer = np.ascontiguousarray(ref), np.ascontiguousarray(qer)
    rc = _ref.ref_kswv_batch(_params(params), pairs.ctypes.data, ref.ctypes.data, len(ref), qer.ctypes.data,
                             len(qer), n, aln.ctypes.data)
    if rc != 0:
        raise RuntimeError("ref_kswv_batch")
    return aln


def make_workload(n, seed=0, read_len=(100, 151), window=(2.0, 5.0), match=1, min_seed_len=19,
                  p_sub=0.04, p_indel=0.004, p_n=0.002, p_unrelated=0.15, p_repeat=0.3, xtra=None):
    """-> (pairs [SEQPAIR_DTYPE], ref bytes, qer bytes). Each pair: a read of l_ms bases and a reference window;
    the read is a mutated copy of a stretch of the window (or unrelated with p_unrelated); with p_repeat a second,
    more diverged copy of the same stretch is planted so that a second-best row maximum exists. h0 carries
    xtra as mem_matesw_batch_pre builds it (bwamem_pair.cpp:1003): KSW_XSUBO | KSW_XSTART |
    (l_ms * a < 250 ? KSW_XBYTE : 0) | min_seed_len * a, unless `xtra` (int or callable(l_ms)) overrides it."""
    from genarchbench_b200 import pairio
    rng = np.random.default_rng(seed)
    pairs = np.zeros(n, dtype=pairio.SEQPAIR_DTYPE)
    refs, qers = [], []
    ro = qo = 0
    for k in range(n):
        l_ms = int(rng.integers(read_len[0], read_len[1] + 1))
        l_ref = max(l_ms, int(l_ms * rng.uniform(*window)))
        t = rng.integers(0, 4, l_ref, dtype=np.uint8)

        def mutate(s, ps, pi):
            out = []
            for b in s:
                u = rng.random()
                if u < pi / 2:
                    continue
                if u < pi:
                    out.append(rng.integers(0, 4))
                out.append(rng.integers(0, 4) if rng.random() < ps else b)
            return np.array(out[:len(s)] + list(rng.integers(0, 4, max(0, len(s) - len(out)))), dtype=np.uint8)

        if rng.random() < p_unrelated:
            q = rng.integers(0, 4, l_ms, dtype=np.uint8)
        else:
            at = int(rng.integers(0, l_ref - l_ms + 1))
            q = mutate(t[at:at + l_ms], p_sub, p_indel)
            if rng.random() < p_repeat and l_ref >= 2 * l_ms:
                at2 = int(rng.integers(0, l_ref - l_ms + 1))
                span = int(rng.integers(l_ms // 3, l_ms + 1))
                t[at2:at2 + span] = mutate(t[at:at + l_ms], 3 * p_sub, 2 * p_indel)[:span]
        t[rng.random(l_ref) < p_n] = 4
        q = q.copy()
        q[rng.random(l_ms) < p_n] = 4
        if xtra is None:
            x = KSW_XSUBO | KSW_XSTART | (KSW_XBYTE if l_ms * match < 250 else 0) | (min_seed_len * match)
        else:
            x = xtra(l_ms) if callable(xtra) else int(xtra)
        p = pairs[k]
        p["idr"], p["idq"], p["id"], p["len1"], p["len2"], p["h0"] = ro, qo, k, l_ref, l_ms, x
        p["regid"] = k
        refs.append(t); qers.append(q)
        ro += l_ref; qo += l_ms
    for f in ("seqid",) + pairio.OUTPUT_FIELDS:
        pairs[f] = -1
    pad = np.zeros(64, np.uint8)
    return pairs, np.concatenate(refs + [pad]), np.concatenate(qers + [pad])


def make_low_complexity(n, seed=0, kind="tandem"):
    """-> (pairs, ref, qer): sequences full of ties -- the cases where "first column of the row maximum", "first row
    of gmax" and the rising-row filter decide the outputs. kind: homopolymer | two-letter | tandem | identical-prefix.
    Lengths 1..259 x 1..699, a mix of flag words in both classes."""
    from genarchbench_b200 import pairio
    rng = np.random.default_rng(seed)
    pairs = np.zeros(n, dtype=pairio.SEQPAIR_DTYPE)
    refs, qers = [], []
    ro = qo = 0
    flags = [KSW_XSUBO | KSW_XSTART | 19, KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 19, KSW_XSTART, KSW_XSTART | KSW_XBYTE,
             KSW_XSTOP | KSW_XSTART | 30, KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 1, KSW_XSUBO | KSW_XSTART]
    for k in range(n):
        l2, l1 = int(rng.integers(1, 260)), int(rng.integers(1, 700))
        if kind == "homopolymer":
            t = np.full(l1, rng.integers(0, 4), np.uint8)
            q = np.full(l2, t[0] if rng.random() < 0.7 else rng.integers(0, 4), np.uint8)
            t[rng.random(l1) < 0.05] = rng.integers(0, 4)
            q[rng.random(l2) < 0.05] = rng.integers(0, 4)
        elif kind == "two-letter":
            t, q = rng.integers(0, 2, l1).astype(np.uint8), rng.integers(0, 2, l2).astype(np.uint8)
        elif kind == "tandem":
            u = rng.integers(0, 4, int(rng.integers(1, 6))).astype(np.uint8)
            t, q = np.resize(u, l1).copy(), np.resize(u, l2).copy()
            t[rng.random(l1) < 0.03] = rng.integers(0, 4)
            q[rng.random(l2) < 0.03] = rng.integers(0, 4)
        else:
            t = rng.integers(0, 4, l1).astype(np.uint8)
            q = t[:l2].copy() if l2 <= l1 else rng.integers(0, 4, l2).astype(np.uint8)
        p = pairs[k]
        p["idr"], p["idq"], p["id"], p["len1"], p["len2"], p["regid"] = ro, qo, k, l1, l2, k
        p["h0"] = int(flags[int(rng.integers(0, len(flags)))])
        refs.append(t); qers.append(q)
        ro += l1; qo += l2
    for f in ("seqid",) + pairio.OUTPUT_FIELDS:
        pairs[f] = -1
    pad = np.zeros(64, np.uint8)
    return pairs, np.concatenate(refs + [pad]), np.concatenate(qers + [pad])
