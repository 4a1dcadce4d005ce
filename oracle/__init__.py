"""TEST INFRASTRUCTURE ONLY -- the oracle for the bsw hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package. The product (``genarchbench_b200``) never does and fails loudly if its
CUDA library is missing.

Two checkers, same call shape (fill the six outputs of a ``PairBatch`` in place):

* ``oracle_batch``    -- oracle/bsw_oracle.c, our plain-C restatement of
                         /root/reference/benchmarks/bsw/src/bandedSWA.cpp (see its header for file:line).
* ``reference_batch`` -- oracle/_ref/libbsw_ref_<isa>.so: the UNMODIFIED reference ``getScores16``
                         compiled from /root/reference by oracle/Makefile, driven like the reference
                         driver's ROI loop (main_banded.cpp:338-350). Pins the restatement and is the
                         ``"kind": "reference"`` CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_PARAMS = dict(o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5,
                      match=1, mismatch=4, ambig=-1)   # main_banded.cpp:70-74,268
DEFAULT_W = 100
_PARAM_ORDER = ("o_del", "e_del", "o_ins", "e_ins", "zdrop", "end_bonus", "match", "mismatch", "ambig")


def build(verbose: bool = False) -> None:
    """Compiles the C restatement and, when /root/reference is present, the reference .so files."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _params_array(params: dict | None) -> C.Array:
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    return (C.c_int32 * 9)(*[int(p[k]) for k in _PARAM_ORDER])


_oracle_lib = None


def _oracle() -> C.CDLL:
    global _oracle_lib
    if _oracle_lib is None:
        path = os.path.join(_HERE, "libbsw_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.bsw_oracle_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_int32]
        L.bsw_oracle_batch.restype = C.c_int
        assert L.bsw_oracle_sizeof_seqpair() == 72
        _oracle_lib = L
    return _oracle_lib


def oracle_batch(batch, w: int = DEFAULT_W, params: dict | None = None, nthreads: int = 0,
                 scalar_zdrop: bool = False) -> int:
    """Fills batch.pairs outputs with the C restatement. Returns the number of DP cells visited."""
    L = _oracle()
    pa = _params_array(params)
    cells = C.c_int64(0)
    nthreads = nthreads or (os.cpu_count() or 1)
    rc = L.bsw_oracle_batch(pa, batch.pairs.ctypes.data, batch.ref.ctypes.data, batch.qer.ctypes.data,
                            len(batch), w, nthreads, C.byref(cells), 1 if scalar_zdrop else 0)
    if rc != 0:
        raise MemoryError("bsw_oracle_batch")
    return int(cells.value)


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def reference_isas() -> list:
    """ISA builds of the reference that exist in oracle/_ref AND this host CPU can run, best first."""
    flags = _cpu_flags()
    out = []
    for isa, need in (("avx512", "avx512bw"), ("avx2", "avx2"), ("sse41", "sse4_1")):
        if need in flags and os.path.exists(os.path.join(_HERE, "_ref", f"libbsw_ref_{isa}.so")):
            out.append(isa)
    return out


_ref_libs: dict = {}


def _reference(isa: str | None = None) -> C.CDLL:
    if isa is None:
        avail = reference_isas()
        if not avail:
            raise RuntimeError("oracle/_ref is empty: run `make -C oracle ref` where /root/reference exists")
        isa = avail[0]
    if isa not in _ref_libs:
        L = C.CDLL(os.path.join(_HERE, "_ref", f"libbsw_ref_{isa}.so"))
        L.ref_bsw_new.argtypes = [C.c_void_p, C.c_int]
        L.ref_bsw_new.restype = C.c_void_p
        L.ref_bsw_free.argtypes = [C.c_void_p]
        L.ref_bsw_getscores16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_int32, C.c_int32, C.POINTER(C.c_double)]
        L.ref_bsw_getscores16.restype = C.c_int
        L.ref_bsw_scalar.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        L.ref_bsw_scalar.restype = C.c_int
        assert L.ref_bsw_sizeof_seqpair() == 72
        _ref_libs[isa] = L
    return _ref_libs[isa]


def reference_available() -> bool:
    return bool(reference_isas())


def reference_batch(batch, w: int = DEFAULT_W, params: dict | None = None, nthreads: int = 0,
                    isa: str | None = None, batch_size: int = 512, scalar: bool = False) -> float:
    """Fills batch.pairs outputs with the reference's own getScores16 (or its scalar kernel).
    Returns the ROI seconds (parallel loop only), as the reference driver times it."""
    L = _reference(isa)
    nthreads = nthreads or (os.cpu_count() or 1)
    h = L.ref_bsw_new(_params_array(params), nthreads)
    try:
        secs = C.c_double(0.0)
        ref = np.ascontiguousarray(batch.ref)
        qer = np.ascontiguousarray(batch.qer)
        if scalar:
            rc = L.ref_bsw_scalar(h, batch.pairs.ctypes.data, ref.ctypes.data, qer.ctypes.data,
                                  len(batch), w)
        else:
            rc = L.ref_bsw_getscores16(h, batch.pairs.ctypes.data, ref.ctypes.data, qer.ctypes.data,
                                       len(batch), w, batch_size, C.byref(secs))
        if rc != 0:
            raise RuntimeError("reference run failed")
        return float(secs.value)
    finally:
        L.ref_bsw_free(h)
