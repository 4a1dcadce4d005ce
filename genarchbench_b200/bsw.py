"""Host-side mirror of the reference's kernel interface over libbsw_gpu.so (include/bsw_gpu.h).

``BandedPairWiseSW`` keeps the reference class's constructor and ``getScores16`` signatures
(/root/reference/benchmarks/bsw/src/bandedSWA.h:132-135, 300-305) so that parity tests read like the
reference driver (main_banded.cpp:266-276, 345); ``BswGpu`` is the thin ctypes view of the C ABI.

There is no CPU fallback: constructing either class without the CUDA library or without a B200
raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from .pairio import PairBatch, SEQPAIR_DTYPE, PACKED_REC_DTYPE, RESULT_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSW_GPU_LIB: developer override used by scripts/ to A/B kernel builds; the product loads lib/libbsw_gpu.so
LIB_PATH = os.environ.get("BSW_GPU_LIB") or os.path.join(_HERE, "lib", "libbsw_gpu.so")

DEFAULT_AMBIG = -1          # bandedSWA.h:61
DEFAULT_W = 100             # main_banded.cpp:268

ERRORS = {1: "BSW_ERR_ARG", 2: "BSW_ERR_NO_DEVICE", 3: "BSW_ERR_CUDA", 4: "BSW_ERR_NOMEM",
          5: "BSW_ERR_RANGE", 6: "BSW_ERR_STATE"}

# every symbol include/bsw_gpu.h declares
EXPORTS = ("bsw_gpu_init", "bsw_gpu_init_devices", "bsw_gpu_free", "bsw_gpu_reserve", "bsw_gpu_batch", "bsw_gpu_batch_retry",
           "bsw_gpu_batch_packed", "bsw_gpu_host_alloc", "bsw_gpu_host_free", "bsw_gpu_classify", "bsw_gpu_trip_probe",
           "bsw_gpu_stage",
           "bsw_gpu_run_staged", "bsw_gpu_fetch_staged", "bsw_gpu_count_staged", "bsw_gpu_get_stats", "bsw_gpu_dpx_peak",
           "bsw_gpu_strerror", "bsw_gpu_last_error", "bsw_gpu_version")


class BswError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"{ERRORS.get(code, code)}{': ' + detail if detail else ''}")


class Params(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("o_del", "e_del", "o_ins", "e_ins", "zdrop", "end_bonus",
                                          "match", "mismatch", "ambig")]


class Stats(C.Structure):
    _fields_ = [("pairs", C.c_int64), ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("pairs_short", C.c_int64), ("pairs_long", C.c_int64),
                ("host_bin_ms", C.c_double), ("host_pack_ms", C.c_double),
                ("host_scatter_ms", C.c_double), ("kernel_ms", C.c_double), ("wall_ms", C.c_double),
                ("n_gpus", C.c_int32), ("reserved", C.c_int32),
                ("host_sort_ms", C.c_double), ("host_plan_ms", C.c_double), ("host_alloc_ms", C.c_double),
                ("host_cut_ms", C.c_double), ("host_wait_ms", C.c_double), ("pairs_keyed", C.c_int64),
                ("pairs_duo", C.c_int64), ("pairs_scalar", C.c_int64), ("pairs_invalid", C.c_int64),
                ("first_invalid", C.c_int64)]

    def asdict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Loads libbsw_gpu.so; fails loudly if it was not built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                               "(make -C genarchbench_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.bsw_gpu_init.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
        L.bsw_gpu_init_devices.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
        L.bsw_gpu_free.argtypes = [vp]
        L.bsw_gpu_free.restype = None
        L.bsw_gpu_reserve.argtypes = [vp, i64, i64]
        L.bsw_gpu_batch.argtypes = [vp, vp, vp, vp, i64, i32]
        L.bsw_gpu_batch_retry.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp]
        L.bsw_gpu_batch_packed.argtypes = [vp, vp, vp, i64, i64, i32, vp]
        L.bsw_gpu_host_alloc.argtypes = [C.c_size_t]
        L.bsw_gpu_host_alloc.restype = vp
        L.bsw_gpu_host_free.argtypes = [vp]
        L.bsw_gpu_host_free.restype = None
        L.bsw_gpu_classify.argtypes = [vp, i64, i32, C.POINTER(i64), vp]
        L.bsw_gpu_trip_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.bsw_gpu_stage.argtypes = [vp, vp, vp, vp, i64, i32]
        L.bsw_gpu_run_staged.argtypes = [vp, C.POINTER(C.c_float)]
        L.bsw_gpu_fetch_staged.argtypes = [vp, vp, i64]
        L.bsw_gpu_count_staged.argtypes = [vp, C.POINTER(C.c_int64)]
        L.bsw_gpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.bsw_gpu_dpx_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.bsw_gpu_strerror.argtypes = [C.c_int]
        L.bsw_gpu_strerror.restype = C.c_char_p
        L.bsw_gpu_last_error.argtypes = [vp]
        L.bsw_gpu_last_error.restype = C.c_char_p
        L.bsw_gpu_version.restype = C.c_int
        _lib = L
    return _lib


class BswGpu:
    """ctypes view of a bsw_handle."""

    def __init__(self, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5, match=1,
                 mismatch=4, ambig=DEFAULT_AMBIG, n_gpus: int = 1,
                 devices: Optional[Sequence[int]] = None):
        self._h = C.c_void_p()
        self._L = lib()
        p = Params(o_del, e_del, o_ins, e_ins, zdrop, end_bonus, match, mismatch, ambig)
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            rc = self._L.bsw_gpu_init_devices(C.byref(p), len(devices), arr, C.byref(self._h))
        else:
            rc = self._L.bsw_gpu_init(C.byref(p), n_gpus, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise BswError(rc, self._L.bsw_gpu_strerror(rc).decode())

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise BswError(rc, self._L.bsw_gpu_last_error(self._h).decode()
                           or self._L.bsw_gpu_strerror(rc).decode())

    @staticmethod
    def _ptrs(pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray):
        assert pairs.dtype == SEQPAIR_DTYPE and pairs.flags["C_CONTIGUOUS"]
        assert ref.dtype == np.uint8 and qer.dtype == np.uint8
        return pairs.ctypes.data, ref.ctypes.data, qer.ctypes.data

    def reserve(self, n_pairs: int, total_bases: int) -> None:
        """bsw_gpu_reserve: pre-sizes pinned rings / device arenas so the first batch() does not allocate."""
        self._check(self._L.bsw_gpu_reserve(self._h, n_pairs, total_bases))

    def batch(self, pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray, w: int = DEFAULT_W,
              n: Optional[int] = None) -> None:
        """bsw_gpu_batch: end to end from host buffers; fills the six outputs in place."""
        a, b, c = self._ptrs(pairs, ref, qer)
        self._check(self._L.bsw_gpu_batch(self._h, a, b, c, len(pairs) if n is None else n, w))

    def batch_retry(self, pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray, w: int = DEFAULT_W,
                    max_tries: int = 2) -> np.ndarray:
        """bsw_gpu_batch_retry: bwa-mem2's band-doubling loop (MAX_BAND_TRY = 2); returns tries per pair."""
        a, b, c = self._ptrs(pairs, ref, qer)
        tries = np.zeros(len(pairs), dtype=np.int32)
        self._check(self._L.bsw_gpu_batch_retry(self._h, a, b, c, len(pairs), w, max_tries, tries.ctypes.data))
        return tries

    def batch_packed(self, rec: np.ndarray, data: np.ndarray, w: int = DEFAULT_W,
                     out: Optional[np.ndarray] = None) -> np.ndarray:
        """bsw_gpu_batch_packed: packed records + 2-bit / 4-bit sequences in, 16-byte result records out."""
        assert rec.dtype == PACKED_REC_DTYPE and rec.flags["C_CONTIGUOUS"] and data.dtype == np.uint8
        if out is None:
            out = np.zeros(len(rec), dtype=RESULT_DTYPE)
        assert out.dtype == RESULT_DTYPE and len(out) >= len(rec)
        self._check(self._L.bsw_gpu_batch_packed(self._h, rec.ctypes.data, data.ctypes.data, data.nbytes, len(rec), w,
                                                 out.ctypes.data))
        return out

    def stage(self, pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray, w: int = DEFAULT_W) -> None:
        a, b, c = self._ptrs(pairs, ref, qer)
        self._check(self._L.bsw_gpu_stage(self._h, a, b, c, len(pairs), w))

    def run_staged(self) -> float:
        ms = C.c_float(0.0)
        self._check(self._L.bsw_gpu_run_staged(self._h, C.byref(ms)))
        return float(ms.value)

    def fetch_staged(self, pairs: np.ndarray) -> None:
        self._check(self._L.bsw_gpu_fetch_staged(self._h, pairs.ctypes.data, len(pairs)))

    def count_staged(self) -> int:
        """DP cells the reference's scalar loop visits for the staged batch (the GCUPS unit of work)."""
        c = C.c_int64(0)
        self._check(self._L.bsw_gpu_count_staged(self._h, C.byref(c)))
        return int(c.value)

    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.bsw_gpu_get_stats(self._h, C.byref(s)))
        return s.asdict()

    def close(self) -> None:
        if self._h:
            self._L.bsw_gpu_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _Pinned:
    """Keeps a bsw_gpu_host_alloc block alive for the numpy array that views it."""
    def __init__(self, nbytes: int):
        self.ptr = lib().bsw_gpu_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError(f"bsw_gpu_host_alloc({nbytes})")

    def __del__(self):
        try:
            lib().bsw_gpu_host_free(self.ptr)
        except Exception:
            pass


def host_alloc(nbytes: int) -> np.ndarray:
    """Page-locked uint8 array (bsw_gpu_host_alloc): packed inputs / result arrays placed here are DMA'd in place."""
    blk = _Pinned(max(nbytes, 1))
    arr = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(blk.ptr))
    arr = arr[:nbytes]
    _keep[arr.ctypes.data] = blk
    return arr


_keep: dict = {}


def results_to_outputs(res: np.ndarray) -> np.ndarray:
    """bsw_result records -> [n, 6] int32 in the order of pairio.OUTPUT_FIELDS."""
    return np.stack([res[f].astype(np.int32) for f in ("score", "qle", "tle", "gtle", "gscore", "max_off")], axis=1)


def classify(pairs: np.ndarray, match: int = 1):
    """bsw_gpu_classify: bwa-mem2's a-priori 8-bit / 16-bit / scalar classes -> (counts[3], class per pair)."""
    cls = np.zeros(len(pairs), dtype=np.uint8)
    counts = (C.c_int64 * 3)()
    rc = lib().bsw_gpu_classify(pairs.ctypes.data, len(pairs), match, counts, cls.ctypes.data)
    if rc != 0:
        raise BswError(rc)
    return list(counts), cls


def dpx_peak(which: int = 0, device: int = 0) -> float:
    """Measured packed-s16x2 instruction throughput, giga thread-instructions / s (all SMs)."""
    v, mhz = C.c_double(0.0), C.c_double(0.0)
    rc = lib().bsw_gpu_dpx_peak(device, which, C.byref(v), C.byref(mhz))
    if rc != 0:
        raise BswError(rc, lib().bsw_gpu_strerror(rc).decode())
    return float(v.value)


class BandedPairWiseSW:
    """Same constructor / call shape as the reference class (bandedSWA.h:127-412).

    ``mat`` is accepted for signature parity; like the reference's vector path, the kernel scores
    from w_match / w_mismatch and the hard-coded DEFAULT_AMBIG (bandedSWA.cpp:63-65), not from it.
    """

    def __init__(self, o_del: int, e_del: int, o_ins: int, e_ins: int, zdrop: int, end_bonus: int,
                 mat=None, w_match: int = 1, w_mismatch: int = 4, numThreads: int = 1,
                 n_gpus: int = 1, devices: Optional[Sequence[int]] = None):
        del mat, numThreads
        self.gpu = BswGpu(o_del, e_del, o_ins, e_ins, zdrop, end_bonus, w_match, w_mismatch,
                          DEFAULT_AMBIG, n_gpus=n_gpus, devices=devices)

    def getScores16(self, pairArray: np.ndarray, seqBufRef: np.ndarray, seqBufQer: np.ndarray,
                    numPairs: int, numThreads: int = 1, w: int = DEFAULT_W) -> None:
        del numThreads
        self.gpu.batch(pairArray, seqBufRef, seqBufQer, w, n=numPairs)

    def run(self, batch: PairBatch, w: int = DEFAULT_W) -> PairBatch:
        self.getScores16(batch.pairs, batch.ref, batch.qer, len(batch), 1, w)
        return batch

    def close(self) -> None:
        self.gpu.close()
