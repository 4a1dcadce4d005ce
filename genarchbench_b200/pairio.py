"""ctypes binding of libbsw_pairio.so (include/bsw_pairio.h): synthetic pair batches in the layout the
reference driver hands to its kernel -- a ``SeqPair`` array plus two byte buffers of base codes
(/root/reference/benchmarks/bsw/src/main_banded.cpp:164-206, bandedSWA.h:104-113).

Host-only; no CUDA and no oracle here.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libbsw_pairio.so")

# numpy mirror of bsw_seqpair / SeqPair (72 bytes, bandedSWA.h:104-113)
SEQPAIR_DTYPE = np.dtype(
    {
        "names": ["idr", "idq", "id", "len1", "len2", "h0", "seqid", "regid",
                  "score", "tle", "gtle", "qle", "gscore", "max_off"],
        "formats": ["<i8", "<i8", "<i8"] + ["<i4"] * 11,
        "offsets": [0, 8, 16, 24, 28, 32, 36, 40, 44, 48, 52, 56, 60, 64],
        "itemsize": 72,
    }
)
OUTPUT_FIELDS = ("score", "qle", "tle", "gtle", "gscore", "max_off")
# numpy mirrors of bsw_packed_rec (12 bytes) and bsw_result (16 bytes), include/bsw_types.h
PACKED_REC_DTYPE = np.dtype([("len1", "<u2"), ("len2", "<u2"), ("h0", "<i4"), ("flags", "<u4")])
RESULT_DTYPE = np.dtype([("score", "<i2"), ("qle", "<i2"), ("tle", "<i2"), ("gtle", "<i2"), ("gscore", "<i2"),
                         ("max_off", "<i2"), ("reserved", "<u4")])


class GenConfig(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("read_len", C.c_int32),
        ("seed_min", C.c_int32), ("seed_max", C.c_int32),
        ("len2_min", C.c_int32), ("len2_max", C.c_int32),
        ("h0_min", C.c_int32), ("h0_max", C.c_int32),
        ("tail_cap", C.c_int32), ("extra_max", C.c_int32),
        ("sub_rate", C.c_double), ("indel_rate", C.c_double),
        ("n_frac", C.c_double), ("small_h0_frac", C.c_double), ("random_frac", C.c_double),
        ("seed", C.c_uint64),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} missing: run __graft_entry__.build() "
                               "(make -C genarchbench_b200/csrc)")
        L = C.CDLL(_LIB_PATH)
        L.bsw_gen_preset.argtypes = [C.c_int, C.POINTER(GenConfig)]
        L.bsw_gen_preset.restype = C.c_int
        L.bsw_gen_pairs.argtypes = [C.POINTER(GenConfig), C.c_int64, C.c_void_p,
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32]
        L.bsw_gen_pairs.restype = C.c_int
        L.bsw_host_free.argtypes = [C.c_void_p]
        L.bsw_host_free.restype = None
        L.bsw_write_pairs_text.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.bsw_write_pairs_text.restype = C.c_int
        L.bsw_count_pairs_text.argtypes = [C.c_char_p]
        L.bsw_count_pairs_text.restype = C.c_int64
        L.bsw_read_pairs_text.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int64)]
        L.bsw_read_pairs_text.restype = C.c_int64
        L.bsw_write_pairs_packed.argtypes = L.bsw_write_pairs_text.argtypes
        L.bsw_write_pairs_packed.restype = C.c_int
        L.bsw_count_pairs_packed.argtypes = [C.c_char_p]
        L.bsw_count_pairs_packed.restype = C.c_int64
        L.bsw_read_pairs_packed.argtypes = L.bsw_read_pairs_text.argtypes
        L.bsw_read_pairs_packed.restype = C.c_int64
        L.bsw_packed_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.bsw_packed_bytes.restype = C.c_int64
        L.bsw_pack_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
        L.bsw_pack_pairs.restype = C.c_int
        L.bsw_packed_file_info.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.bsw_packed_file_info.restype = C.c_int
        L.bsw_read_packed_raw.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
        L.bsw_read_packed_raw.restype = C.c_int64
        _lib = L
    return _lib


@dataclass
class PairBatch:
    """A pair batch as the reference driver holds it: SeqPair[n] + seqBufRef + seqBufQer."""
    pairs: np.ndarray   # SEQPAIR_DTYPE[n]
    ref: np.ndarray     # uint8, target bases (codes 0..4), indexed by pairs["idr"]
    qer: np.ndarray     # uint8, query bases, indexed by pairs["idq"]

    def __len__(self) -> int:
        return int(self.pairs.shape[0])

    def copy(self) -> "PairBatch":
        return PairBatch(self.pairs.copy(), self.ref, self.qer)

    def slice(self, lo: int, hi: int) -> "PairBatch":
        return PairBatch(self.pairs[lo:hi].copy(), self.ref, self.qer)

    def outputs(self) -> np.ndarray:
        return np.stack([self.pairs[f] for f in OUTPUT_FIELDS], axis=1)

    def cells_rect(self) -> int:
        return int((self.pairs["len1"].astype(np.int64) * self.pairs["len2"].astype(np.int64)).sum())


def concat(batches) -> PairBatch:
    """Concatenates batches into one (offsets rebased, ids renumbered)."""
    batches = list(batches)
    n = sum(len(b) for b in batches)
    pairs = np.zeros(n, dtype=SEQPAIR_DTYPE)
    lo = ro = qo = 0
    for b in batches:
        part = pairs[lo:lo + len(b)]
        part[...] = b.pairs
        part["idr"] += ro
        part["idq"] += qo
        lo += len(b); ro += len(b.ref); qo += len(b.qer)
    pairs["id"] = np.arange(n)
    return PairBatch(pairs, np.concatenate([b.ref for b in batches]), np.concatenate([b.qer for b in batches]))


def _take(ptr: C.c_void_p, nbytes: int) -> np.ndarray:
    """Copies a malloc'd buffer (+64 bytes slack) into numpy and frees it."""
    buf = (C.c_uint8 * (nbytes + 64)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=np.uint8).copy()
    lib().bsw_host_free(ptr)
    return arr


def preset(config_id: int) -> GenConfig:
    cfg = GenConfig()
    if lib().bsw_gen_preset(config_id, C.byref(cfg)) != 0:
        raise ValueError(f"unknown config {config_id}")
    return cfg


def generate(cfg: GenConfig | int, n: int, seed: int | None = None, nthreads: int = 0) -> PairBatch:
    if isinstance(cfg, int):
        cfg = preset(cfg)
    if seed is not None:
        cfg.seed = seed
    pairs = np.zeros(n, dtype=SEQPAIR_DTYPE)
    rp, qp = C.c_void_p(), C.c_void_p()
    rb, qb = C.c_int64(), C.c_int64()
    rc = lib().bsw_gen_pairs(C.byref(cfg), n, pairs.ctypes.data, C.byref(rp), C.byref(qp),
                             C.byref(rb), C.byref(qb), nthreads)
    if rc != 0:
        raise RuntimeError(f"bsw_gen_pairs failed ({rc})")
    return PairBatch(pairs, _take(rp, rb.value), _take(qp, qb.value))


def from_sequences(items, h0_default: int = 0) -> PairBatch:
    """items: iterable of (target_codes, query_codes, h0). For hand-written edge cases."""
    items = list(items)
    pairs = np.zeros(len(items), dtype=SEQPAIR_DTYPE)
    ref, qer = [], []
    ro = qo = 0
    for k, it in enumerate(items):
        t, q = np.asarray(it[0], dtype=np.uint8), np.asarray(it[1], dtype=np.uint8)
        h0 = it[2] if len(it) > 2 else h0_default
        pairs[k] = (ro, qo, k, len(t), len(q), h0, -1, -1, -1, -1, -1, -1, -1, -1)
        ref.append(t); qer.append(q)
        ro += len(t); qo += len(q)
    pad = np.zeros(64, dtype=np.uint8)
    return PairBatch(pairs, np.concatenate(ref + [pad]), np.concatenate(qer + [pad]))


def write_text(path: str, b: PairBatch) -> None:
    rc = lib().bsw_write_pairs_text(path.encode(), b.pairs.ctypes.data, b.ref.ctypes.data,
                                    b.qer.ctypes.data, len(b))
    if rc != 0:
        raise OSError(f"bsw_write_pairs_text({path}) failed ({rc})")


def read_text(path: str) -> PairBatch:
    n = lib().bsw_count_pairs_text(path.encode())
    if n < 0:
        raise OSError(f"cannot read {path}")
    pairs = np.zeros(n, dtype=SEQPAIR_DTYPE)
    rp, qp = C.c_void_p(), C.c_void_p()
    rb, qb = C.c_int64(), C.c_int64()
    got = lib().bsw_read_pairs_text(path.encode(), n, pairs.ctypes.data, C.byref(rp), C.byref(qp),
                                    C.byref(rb), C.byref(qb))
    if got < 0:
        raise OSError(f"malformed pair file {path}")
    return PairBatch(pairs[:got], _take(rp, rb.value), _take(qp, qb.value))


def write_packed(path: str, b: PairBatch) -> None:
    """Packed binary pair file (include/bsw_pairio.h): 2 bits per base, ~3.5x smaller than the text format."""
    rc = lib().bsw_write_pairs_packed(path.encode(), b.pairs.ctypes.data, b.ref.ctypes.data,
                                      b.qer.ctypes.data, len(b))
    if rc != 0:
        raise OSError(f"bsw_write_pairs_packed({path}) failed ({rc})")


def pack(b: PairBatch, alloc=None):
    """The packed in-memory form of a batch (what a BSWPAIR1 file holds): (records, data). `alloc(nbytes)` may
    return page-locked uint8 arrays (bsw.host_alloc) so that bsw_gpu_batch_packed DMAs them in place."""
    n = len(b)
    nbytes = lib().bsw_packed_bytes(b.pairs.ctypes.data, b.ref.ctypes.data, b.qer.ctypes.data, n)
    if nbytes < 0:
        raise ValueError("a sequence length is out of range")
    alloc = alloc or (lambda k: np.zeros(k, dtype=np.uint8))
    rec = alloc(max(n, 1) * PACKED_REC_DTYPE.itemsize)[:n * PACKED_REC_DTYPE.itemsize].view(PACKED_REC_DTYPE)
    data = alloc(nbytes + 64)
    data[nbytes:] = 0
    rc = lib().bsw_pack_pairs(b.pairs.ctypes.data, b.ref.ctypes.data, b.qer.ctypes.data, n, rec.ctypes.data,
                              data.ctypes.data, nbytes)
    if rc != 0:
        raise RuntimeError(f"bsw_pack_pairs failed ({rc})")
    return rec, data[:nbytes]


def read_packed_raw(path: str, alloc=None):
    """Records and packed data of a BSWPAIR1 file exactly as stored (no unpacking to one byte per base)."""
    n, nbytes = C.c_int64(), C.c_int64()
    if lib().bsw_packed_file_info(path.encode(), C.byref(n), C.byref(nbytes)) != 0:
        raise OSError(f"{path} is not a packed pair file")
    alloc = alloc or (lambda k: np.zeros(k, dtype=np.uint8))
    rec = alloc(max(n.value, 1) * PACKED_REC_DTYPE.itemsize)[:n.value * PACKED_REC_DTYPE.itemsize].view(PACKED_REC_DTYPE)
    data = alloc(nbytes.value + 64)
    got = lib().bsw_read_packed_raw(path.encode(), n.value, rec.ctypes.data, data.ctypes.data, nbytes.value)
    if got < 0:
        raise OSError(f"malformed packed pair file {path}")
    return rec[:got], data[:nbytes.value]


def read_packed(path: str) -> PairBatch:
    n = lib().bsw_count_pairs_packed(path.encode())
    if n < 0:
        raise OSError(f"{path} is not a packed pair file")
    pairs = np.zeros(n, dtype=SEQPAIR_DTYPE)
    rp, qp = C.c_void_p(), C.c_void_p()
    rb, qb = C.c_int64(), C.c_int64()
    got = lib().bsw_read_pairs_packed(path.encode(), n, pairs.ctypes.data, C.byref(rp), C.byref(qp),
                                      C.byref(rb), C.byref(qb))
    if got < 0:
        raise OSError(f"malformed packed pair file {path}")
    return PairBatch(pairs[:got], _take(rp, rb.value), _take(qp, qb.value))
