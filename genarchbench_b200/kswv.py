"""Host-side mirror of bwa-mem2's ``kswv`` class over libbsw_gpu.so (include/kswv_gpu.h).

``Kswv`` keeps the reference constructor's scoring arguments
(/root/reference/benchmarks/fmi/bwa-mem2/x86_64/src/kswv.cpp:117-124); ``align`` is the vector branch of
``mem_sam_pe_batch`` (bwamem_pair.cpp:634-704) in one call: both score classes, phase 0 and phase 1.

There is no CPU fallback: constructing ``Kswv`` without the CUDA library or without a B200 raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import bsw
from .pairio import SEQPAIR_DTYPE

KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000   # ksw.h:31-34
RESULT_FIELDS = ("score", "te", "qe", "score2", "te2", "tb", "qb")                 # kswr_t, ksw.h:45-50

# every symbol include/kswv_gpu.h declares
EXPORTS = ("kswv_gpu_init", "kswv_gpu_free", "kswv_gpu_batch", "kswv_gpu_get_stats", "kswv_gpu_last_error")


class Params(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("o_del", "e_del", "o_ins", "e_ins", "match", "mismatch")]


class Stats(C.Structure):
    _fields_ = [("n_gpus", C.c_int32), ("chunks", C.c_int32), ("lanes_per_pair", C.c_int32), ("reserved", C.c_int32),
                ("pairs", C.c_int64), ("pairs8", C.c_int64),
                ("cells", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("kernel_launches", C.c_int64), ("gathered", C.c_int64), ("kernel_ms", C.c_double),
                ("wall_ms", C.c_double), ("host_check_ms", C.c_double), ("host_prep_ms", C.c_double),
                ("host_wait_ms", C.c_double), ("staged", C.c_int64)]

    def asdict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_bound = False


def lib() -> C.CDLL:
    global _bound
    L = bsw.lib()
    if not _bound:
        vp = C.c_void_p
        L.kswv_gpu_init.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
        L.kswv_gpu_free.argtypes = [vp]
        L.kswv_gpu_free.restype = None
        L.kswv_gpu_batch.argtypes = [vp, vp, vp, vp, C.c_int64, vp]
        L.kswv_gpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.kswv_gpu_last_error.argtypes = [vp]
        L.kswv_gpu_last_error.restype = C.c_char_p
        _bound = True
    return L


class Kswv:
    """ctypes view of a kswv_handle."""

    def __init__(self, o_del=6, e_del=1, o_ins=6, e_ins=1, w_match=1, w_mismatch=4, n_gpus: int = 1):
        self._h = C.c_void_p()
        self._L = lib()
        p = Params(o_del, e_del, o_ins, e_ins, w_match, w_mismatch)
        rc = self._L.kswv_gpu_init(C.byref(p), n_gpus, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise bsw.BswError(rc, self._L.bsw_gpu_strerror(rc).decode())

    def align(self, pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray, aln: np.ndarray | None = None) -> np.ndarray:
        """kswv_gpu_batch -> int32 [n, 7] rows of kswr_t, row index = pairs['regid']."""
        assert pairs.dtype == SEQPAIR_DTYPE and pairs.flags["C_CONTIGUOUS"]
        assert ref.dtype == np.uint8 and qer.dtype == np.uint8
        n = len(pairs)
        if aln is None:
            aln = np.full((n, 7), -9, np.int32)
        assert aln.dtype == np.int32 and aln.shape == (n, 7) and aln.flags["C_CONTIGUOUS"]
        rc = self._L.kswv_gpu_batch(self._h, pairs.ctypes.data, ref.ctypes.data, qer.ctypes.data, n, aln.ctypes.data)
        if rc != 0:
            raise bsw.BswError(rc, self._L.kswv_gpu_last_error(self._h).decode()
                               or self._L.bsw_gpu_strerror(rc).decode())
        return aln

    def stats(self) -> dict:
        s = Stats()
        self._L.kswv_gpu_get_stats(self._h, C.byref(s))
        return s.asdict()

    def close(self) -> None:
        if self._h:
            self._L.kswv_gpu_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
