"""INTEGRATION.md's patch, proven on the reference driver itself: integration/Makefile applies
integration/main_banded.patch to a scratch copy of /root/reference/benchmarks/bsw/src/main_banded.cpp, links it with
libbsw_gpu.so (integration/_build/main_bsw_gpu) and builds the unmodified driver next to it (main_bsw_stock). Both
run on the same pair file; their "[i] score=" lines -- what scripts/regression_small.sh:89-96 diffs against its
golden file -- must be identical."""
import os
import re
import subprocess

import pytest

import oracle
from conftest import ROOT
from genarchbench_b200 import pairio

BUILD = os.path.join(ROOT, "integration", "_build")
STOCK = os.path.join(BUILD, "main_bsw_stock")
PATCHED = os.path.join(BUILD, "main_bsw_gpu")
HAVE_REF = os.path.exists("/root/reference/benchmarks/bsw/src/main_banded.cpp")


def scores(stderr: str, n: int):
    got = [int(m.group(2)) for m in re.finditer(r"^\[(\d+)\] score=(-?\d+)\r?$", stderr, re.M)]
    return got[:n]       # the reference also prints its uninitialised padding entries (main_banded.cpp:407-409)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (the GPU box uses the prebuilt binaries)")
def test_patch_applies_and_the_patched_driver_links(tmp_path):
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "integration"), "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.exists(STOCK) and os.path.exists(PATCHED)
    # the stock driver is the reference: its scores are the oracle's (CPU only)
    b = pairio.generate(1, 3000, seed=77)
    a = b.copy()
    oracle.oracle_batch(a)
    path = str(tmp_path / "pairs.txt")
    pairio.write_text(path, b)
    r = subprocess.run([STOCK, "-pairs", path, "-t", "2", "-b", "512"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert scores(r.stderr, len(b)) == a.pairs["score"].tolist()
    # the patched driver needs libbsw_gpu.so at run time and nothing of the CPU kernel in its call path
    ldd = subprocess.run(["ldd", PATCHED], capture_output=True, text=True).stdout
    assert "libbsw_gpu.so" in ldd


@pytest.mark.gpu
def test_patched_reference_driver_prints_the_stock_scores(tmp_path):
    if not (os.path.exists(STOCK) and os.path.exists(PATCHED)):
        pytest.skip("integration/_build not built (make -C integration, where /root/reference exists)")
    b = pairio.generate(1, 100_000)                            # BASELINE config 1
    path = str(tmp_path / "pairs.txt")
    pairio.write_text(path, b)
    ref = subprocess.run([STOCK, "-pairs", path, "-t", "8", "-b", "512"], capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0, ref.stderr[-500:]
    want = scores(ref.stderr, len(b))
    assert len(want) == len(b)
    for bflag in ("0", "512"):                                 # one call over everything / the driver's batches
        got = subprocess.run([PATCHED, "-pairs", path, "-t", "8", "-b", bflag], capture_output=True, text=True, timeout=600)
        assert got.returncode == 0, got.stderr[-500:]
        assert scores(got.stderr, len(b)) == want, f"-b {bflag}"
        assert "libbsw_gpu.so" in got.stdout and "Total Pairs processed: 100000" in got.stdout


KSWV_BINDING = os.path.join(BUILD, "kswv_binding")
KSWV_REF = os.path.join(ROOT, "oracle", "_ref", "libkswv_ref_avx512.so")


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (the GPU box uses the prebuilt binaries)")
def test_kswv_binding_compiles_against_the_reference_headers():
    """a translation unit that includes the reference's kswv.h (SeqPair, kswr_t, KSW_X*) and ours,
    with the layouts asserted equal at compile time, builds and links with libbsw_gpu.so."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "integration"), "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.exists(KSWV_BINDING)
    assert "libbsw_gpu.so" in subprocess.run(["ldd", KSWV_BINDING], capture_output=True, text=True).stdout


@pytest.mark.gpu
def test_kswv_binding_matches_the_reference_class():
    """The reference-typed arrays through kswv_gpu_batch (both classes, both phases) against the unmodified kswv class
    on the same batch, in one C++ program."""
    from oracle import kswv as okswv
    if not os.path.exists(KSWV_BINDING):
        pytest.skip("integration/_build not built (make -C integration, where /root/reference exists)")
    if not okswv.reference_available():
        pytest.skip("the compiled reference needs AVX512BW on this host")
    r = subprocess.run([KSWV_BINDING, "20000", KSWV_REF], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    m = re.search(r"kswv binding: (\d+) pairs, score sum (\d+), (\d+) with start positions, mismatches vs the reference class: (-?\d+)", r.stdout)
    assert m and int(m.group(1)) == 20000 and int(m.group(4)) == 0 and int(m.group(3)) > 10000, r.stdout
