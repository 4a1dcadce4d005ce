"""The kswv device code (genarchbench_b200/csrc/kswv_kernels.cuh: strips of columns per lane, the row pipeline
through shuffles, the second-best pass, the reverse phase) compiled for the CPU -- DPX / PRMT through
tests/host_emul/dpx_host_emul.h, the warp as 32 fibers (warp_fibers.h) -- against the golden vectors and the oracle.
Catches algorithmic errors where no GPU exists; the real kernel is checked by tests/test_kswv_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import KSWV_GOLDEN_NAMES, ROOT, assert_same_aln, load_kswv_golden
from oracle import kswv
from oracle.kswv import KSW_XBYTE, KSW_XSTART, KSW_XSTOP, KSW_XSUBO


@pytest.fixture(scope="module")
def emul():
    d = os.path.join(ROOT, "tests", "host_emul")
    so = os.path.join(d, "libkswv_emul.so")
    srcs = [os.path.join(d, "kswv_emul_lib.cpp"), os.path.join(d, "warp_fibers.h"),
            os.path.join(ROOT, "genarchbench_b200", "csrc", "kswv_kernels.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-Wall", "-Wno-unknown-pragmas",
                        "-I", d, "-I", os.path.join(ROOT, "genarchbench_b200", "csrc"),
                        "-I", os.path.join(ROOT, "include"), "-o", so, srcs[0]], check=True)
    L = C.CDLL(so)
    L.kswv_emul_batch.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_int32]

    def run(pairs, ref, qer, params=None, width=32):
        aln = np.full((len(pairs), 7), -7, np.int32)
        rc = L.kswv_emul_batch(kswv._params(params), pairs.ctypes.data, np.ascontiguousarray(ref).ctypes.data,
                               np.ascontiguousarray(qer).ctypes.data, len(pairs), aln.ctypes.data, width)
        assert rc == 0, f"lanes of a group disagree, or a pair does not fit the width ({rc})"
        return aln
    return run


@pytest.mark.parametrize("name", KSWV_GOLDEN_NAMES)
def test_device_code_matches_golden(emul, name):
    pairs, ref, qer, params, want = load_kswv_golden(name)
    assert_same_aln(emul(pairs, ref, qer, params), want, pairs, f"emulated kernel vs golden[{name}]")


CASES = [
    ("every strip width", None, dict(n=800, read_len=(1, 300), window=(0.3, 4.0), min_seed_len=5)),
    ("several passes (above 256 columns), 16-bit", None, dict(n=60, read_len=(257, 700))),
    ("several passes, 8-bit class", None, dict(n=60, read_len=(257, 600), p_sub=0.3,
                                              xtra=lambda l: KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 19)),
    ("stop + start", None, dict(n=200, xtra=lambda l: KSW_XSTOP | KSW_XSTART | (KSW_XBYTE if l < 120 else 0) | 45)),
    ("stop at zero: the first row ends the pair", None, dict(n=150, read_len=(1, 300), window=(0.3, 3.0), xtra=lambda l: KSW_XSTOP | KSW_XSTART)),
    ("asymmetric gaps, a=2", dict(match=2, mismatch=5, o_del=4, e_del=2, o_ins=7, e_ins=1),
     dict(n=200, match=2, read_len=(60, 180), p_indel=0.03)),
]


@pytest.mark.parametrize("what,params,kw", CASES, ids=[c[0] for c in CASES])
def test_device_code_matches_oracle(emul, what, params, kw):
    pairs, ref, qer = kswv.make_workload(seed=21, **kw)
    want, _ = kswv.oracle_batch(pairs, ref, qer, params)
    assert_same_aln(emul(pairs, ref, qer, params), want, pairs, what)


@pytest.mark.parametrize("kind", ("homopolymer", "two-letter", "tandem", "identical-prefix"))
def test_device_code_on_ties(emul, kind):
    pairs, ref, qer = kswv.make_low_complexity(300, seed=4, kind=kind)
    want, _ = kswv.oracle_batch(pairs, ref, qer)
    assert_same_aln(emul(pairs, ref, qer), want, pairs, kind)
    keep = pairs["len2"] <= 160                      # the narrow groups take what fits them and needs no clamp
    sub = pairs[keep].copy()
    sub["h0"] &= ~KSW_XBYTE | 0                      # 16-bit class: never clamped
    sub["regid"] = np.arange(len(sub))
    want, _ = kswv.oracle_batch(sub, ref, qer)
    assert_same_aln(emul(sub, ref, qer, width=8), want, sub, kind + ", 8 lanes")


def test_empty_sequences(emul):
    pairs, ref, qer = kswv.make_workload(40, seed=3, read_len=(5, 40), min_seed_len=3)
    pairs["len1"][::4] = 0
    pairs["len2"][1::4] = 0
    want, _ = kswv.oracle_batch(pairs, ref, qer)
    assert_same_aln(emul(pairs, ref, qer), want, pairs, "empty reference / query")


# lanes per pair below 32: several pairs share a warp, each on its own group of lanes
NARROW = [
    (8, "151 bp reads", None, dict(n=203, read_len=(100, 151))),
    (8, "every strip width up to 160 columns", None, dict(n=403, read_len=(1, 160), window=(0.3, 4.0), min_seed_len=5)),
    (8, "stop + start, a = 2", dict(match=2, mismatch=5, o_del=4, e_del=2, o_ins=7, e_ins=1),
     dict(n=201, match=2, read_len=(40, 110), xtra=lambda l: KSW_XSTOP | KSW_XSTART | KSW_XBYTE | 45)),
    (16, "both classes up to 256 columns", None, dict(n=203, read_len=(120, 256), window=(0.5, 3.0))),
    (16, "every strip width up to 256 columns", None, dict(n=301, read_len=(1, 249), window=(0.3, 4.0), min_seed_len=5)),
]


@pytest.mark.parametrize("width,what,params,kw", NARROW, ids=[f"W{c[0]}: {c[1]}" for c in NARROW])
def test_several_pairs_per_warp(emul, width, what, params, kw):
    pairs, ref, qer = kswv.make_workload(seed=23, **kw)
    want, _ = kswv.oracle_batch(pairs, ref, qer, params)
    assert_same_aln(emul(pairs, ref, qer, params, width=width), want, pairs, what)


def test_groups_with_empty_and_unequal_pairs(emul):
    """Neighbouring pairs of very different sizes, empty sequences and a ragged tail share warps."""
    pairs, ref, qer = kswv.make_workload(101, seed=24, read_len=(5, 150), window=(0.2, 6.0), min_seed_len=3)
    pairs["len1"][::5] = 0
    pairs["len2"][2::7] = 0
    want, _ = kswv.oracle_batch(pairs, ref, qer)
    for width in (8, 16):
        assert_same_aln(emul(pairs, ref, qer, width=width), want, pairs, f"W={width}")


def test_emulated_device_code_is_clean_under_asan(tmp_path):
    """Out-of-bounds check of the kswv device code at all three lane-group widths (the GPU pool has no
    compute-sanitizer): an AddressSanitizer build of the emulation library, scratch and sequence buffers at exactly
    the sizes the kernels get."""
    import sys
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    asan = subprocess.run([cxx, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.exists(asan):
        pytest.skip("libasan not available")
    d = os.path.join(ROOT, "tests", "host_emul")
    so = str(tmp_path / "libkswv_emul_asan.so")
    subprocess.run([cxx, "-O1", "-g", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-w", "-fsanitize=address",
                    "-fno-omit-frame-pointer", f"-I{d}", f"-I{ROOT}/genarchbench_b200/csrc", f"-I{ROOT}/include",
                    "-o", so, os.path.join(d, "kswv_emul_lib.cpp")], check=True)
    # the fibers switch stacks behind ASan's back: its stack-use-after-return machinery is off, heap checks stay on
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0", OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(d, "kswv_asan_check.py"), so], capture_output=True, text=True,
                       env=env, timeout=900)
    assert r.returncode == 0 and "ERROR: AddressSanitizer" not in r.stderr, (r.stdout[-500:], r.stderr[-2000:])
