"""The product's per-pair device code (genarchbench_b200/csrc/bsw_kernels.cuh: unpack_pair +
extend_pair) and host packers, compiled for the CPU against an emulation of the CUDA intrinsics
(tests/host_emul), checked against the golden vectors and the oracle. Catches algorithmic errors
without a GPU; the GPU parity tests (-m gpu) remain the real gate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import GOLDEN_NAMES, ROOT, assert_same_outputs, load_golden
from genarchbench_b200 import pairio


@pytest.fixture(scope="module")
def emul():
    d = os.path.join(ROOT, "tests", "host_emul")
    so = os.path.join(d, "libbsw_emul.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-w", f"-I{d}",
                    f"-I{ROOT}/genarchbench_b200/csrc", f"-I{ROOT}/include", "-o", so,
                    os.path.join(d, "emul_lib.cpp")], check=True)
    L = C.CDLL(so)
    L.bsw_emul_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
    L.bsw_emul_batch_duo.argtypes = L.bsw_emul_batch.argtypes
    L.bsw_emul_batch_duo_key.argtypes = L.bsw_emul_batch.argtypes
    L.bsw_emul_batch_win.argtypes = L.bsw_emul_batch.argtypes
    L.bsw_emul_batch_key.argtypes = L.bsw_emul_batch.argtypes
    L.bsw_emul_pack_check.argtypes = [C.c_int64, C.c_uint32]
    L.bsw_emul_pack_check.restype = C.c_int64

    def run(b, w=100, params=None, duo=False, win=False, key=False):
        fn = L.bsw_emul_batch_duo if duo else (L.bsw_emul_batch_win if win else L.bsw_emul_batch)
        if key:
            fn = L.bsw_emul_batch_duo_key if duo else L.bsw_emul_batch_key
        fn(oracle._params_array(params), b.pairs.ctypes.data, b.ref.ctypes.data, b.qer.ctypes.data, len(b), w)
        return b.outputs()
    run.lib = L
    return run


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_device_code_matches_golden(emul, name):
    b, w, params, want = load_golden(name)
    assert_same_outputs(emul(b, w, params), want, b, f"emulated kernel vs golden[{name}]")


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_windowed_rows_match_golden(emul, name):
    """extend_pair<.., WIN>: rows kept in a sliding window of 2w + 16 columns."""
    b, w, params, want = load_golden(name)
    assert_same_outputs(emul(b, w, params, win=True), want, b, f"emulated windowed kernel vs golden[{name}]")


@pytest.mark.parametrize("w", [1, 2, 5, 17, 40])
def test_windowed_rows_match_oracle_on_long_queries(emul, w):
    """Queries several windows long (up to 600 bases against windows of 32..128 columns), ambiguous bases,
    unrelated pairs; the COUNT variant's cell count as well."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 600, 0, 120, 0.3, 0.15
    b = pairio.generate(c, 3000, seed=700 + w)
    a = b.copy()
    cells = oracle.oracle_batch(a, w=w)
    assert_same_outputs(emul(b, w, win=True), a.outputs(), b, f"emulated windowed kernel vs oracle, w={w}")
    assert int(b.pairs["seqid"].astype(np.int64).sum()) == cells


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_keyed_argmax_matches_golden(emul, name):
    """extend_pair<.., KEY>: the row's last argmax as the lane maximum of score << kbits | group."""
    b, w, params, want = load_golden(name)
    assert_same_outputs(emul(b, w, params, key=True), want, b, f"emulated keyed kernel vs golden[{name}]")


@pytest.mark.parametrize("w", [1, 5, 17, 100])
def test_keyed_argmax_matches_oracle(emul, w):
    """Scores and group indices right up to the limits of the 16-bit key (each pair runs with the
    narrowest index field its query allows), ties between lanes and groups, unrelated pairs."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 260, 0, 400, 0.3, 0.2
    b = pairio.generate(c, 8000, seed=900 + w)
    a = b.copy()
    oracle.oracle_batch(a, w=w)
    assert_same_outputs(emul(b, w, key=True), a.outputs(), b, f"emulated keyed kernel vs oracle, w={w}")
    keyed = int((b.pairs["seqid"] == -1).sum())
    assert 0.3 * len(b) < keyed < len(b), keyed      # both the keyed and the general path were exercised


@pytest.mark.parametrize("key", [False, True])
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_duo_code_matches_golden(emul, name, key):
    """Two pairs per thread (extend_duo2): neighbours in the fixture's order share the DPX lanes."""
    b, w, params, want = load_golden(name)
    assert_same_outputs(emul(b, w, params, duo=True, key=key), want, b, f"emulated duo kernel vs golden[{name}]")


@pytest.mark.parametrize("key", [False, True])
@pytest.mark.parametrize("sort", [False, True])
@pytest.mark.parametrize("w", [1, 2, 5, 17, 100])
def test_duo_code_matches_oracle_on_small_bands(emul, w, sort, key):
    """Unsorted neighbours (lengths, seeds and ends all differ: masked blocks, ghost lanes) and neighbours
    in the device's (len2, len1, h0) launch order (equal ends, the fast trips)."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 200, 0, 60, 0.3, 0.2
    b = pairio.generate(c, 6001, seed=300 + w)                # odd count: the last thread has one pair
    if sort:
        o = np.lexsort((b.pairs["h0"], b.pairs["len1"], b.pairs["len2"]))
        b = pairio.PairBatch(b.pairs[o].copy(), b.ref, b.qer)
    a = b.copy()
    oracle.oracle_batch(a, w=w)
    assert_same_outputs(emul(b, w, duo=True, key=key), a.outputs(), b, f"emulated duo kernel vs oracle, w={w}")
    if key:
        assert int((b.pairs["seqid"] == -1).sum()) > 0.9 * len(b)


def test_duo_code_nondefault_scoring(emul):
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 150, 0, 100, 0.2, 0.2
    b0 = pairio.generate(c, 5000, seed=11)
    for params in (dict(o_del=5, e_del=2, o_ins=7, e_ins=1, zdrop=40, end_bonus=9, match=2, mismatch=3, ambig=-1),
                   dict(o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=0, end_bonus=5, match=1, mismatch=4, ambig=-1)):
        for key in (False, True):
            a = b0.copy(); b = b0.copy()
            oracle.oracle_batch(a, w=30, params=params)
            assert_same_outputs(emul(b, 30, params, duo=True, key=key), a.outputs(), b, f"emulated duo kernel, {params}")


@pytest.mark.parametrize("w", [1, 2, 5, 17, 100])
def test_device_code_matches_oracle_on_small_bands(emul, w):
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 200, 0, 60, 0.3, 0.2
    b = pairio.generate(c, 6000, seed=300 + w)
    a = b.copy()
    oracle.oracle_batch(a, w=w)
    assert_same_outputs(emul(b, w), a.outputs(), b, f"emulated kernel vs oracle, w={w}")


def test_simd_packers_match_the_scalar_packer(emul):
    """pack_pair_avx2 / pack2bit_avx2 / pack2bit_pext (the host pass of bsw_gpu_batch) against the scalar
    packer: every length 0..299 x 0..699 at random, ambiguous bases, tails that end at a protected page."""
    assert emul.lib.bsw_emul_pack_check(200000, 12345) == 0


def test_emulated_device_code_is_clean_under_asan(tmp_path):
    """Out-of-bounds check of the per-pair device code (the GPU pool has no compute-sanitizer): an
    AddressSanitizer build of the emulation library, rows allocated at exactly the size the kernels use."""
    import sys
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    asan = subprocess.run([cxx, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.exists(asan):
        pytest.skip("libasan not available")
    d = os.path.join(ROOT, "tests", "host_emul")
    so = str(tmp_path / "libbsw_emul_asan.so")
    subprocess.run([cxx, "-O1", "-g", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-w", "-fsanitize=address",
                    "-fno-omit-frame-pointer", f"-I{d}", f"-I{ROOT}/genarchbench_b200/csrc", f"-I{ROOT}/include",
                    "-o", so, os.path.join(d, "emul_lib.cpp")], check=True)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(d, "asan_check.py"), so], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr, (r.stdout[-500:], r.stderr[-2000:])
