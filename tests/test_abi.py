"""The drop-in boundary: libbsw_gpu.so loads without a GPU, exports every symbol include/bsw_gpu.h
declares, and refuses to run (no CPU fallback) when there is no device. No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from genarchbench_b200 import bsw


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bsw_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bsw_gpu_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    L = bsw.lib()
    names = declared_symbols()
    assert set(names) == set(bsw.EXPORTS), (names, bsw.EXPORTS)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/bsw_gpu.h but not exported"
    assert L.bsw_gpu_version() >= 1


def test_error_strings_and_argument_checks():
    L = bsw.lib()
    assert L.bsw_gpu_strerror(0) == b"ok"
    assert b"no CPU fallback" in L.bsw_gpu_strerror(2)
    h = C.c_void_p()
    assert L.bsw_gpu_init(None, 1, C.byref(h)) == 1                      # BSW_ERR_ARG
    bad = bsw.Params(6, 0, 6, 1, 100, 5, 1, 4, -1)                       # e_del = 0 divides by zero
    assert L.bsw_gpu_init(C.byref(bad), 1, C.byref(h)) == 1
    assert L.bsw_gpu_batch(None, None, None, None, 0, 100) == 1
    assert L.bsw_gpu_get_stats(None, None) == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bsw.BswError) as e:
        bsw.BswGpu()
    assert e.value.code == 2                                            # BSW_ERR_NO_DEVICE


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under genarchbench_b200/ may reference it."""
    pkg = os.path.join(ROOT, "genarchbench_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "libbsw_oracle" not in src and "oracle/_ref" not in src and "libbsw_ref" not in src, f


def test_a_priori_classes_follow_bwa_mem2():
    """bsw_gpu_classify == bwamem.cpp:2218-2228 (host-only, no device needed)."""
    import numpy as np
    from genarchbench_b200 import pairio
    p = np.zeros(6, dtype=pairio.SEQPAIR_DTYPE)
    #            len1  len2   h0      -> class
    cases = [(100, 100, 27, 0), (100, 100, 28, 1), (128, 10, 0, 1), (300, 200, 32567, 1), (300, 200, 32568, 2),
             (10, 32767, 32757, 1)]
    for k, (l1, l2, h0, _) in enumerate(cases):
        p[k]["len1"], p[k]["len2"], p[k]["h0"] = l1, l2, h0
    counts, cls = bsw.classify(p, 1)
    assert cls.tolist() == [c[3] for c in cases] and counts == [1, 4, 1]


def test_kswv_header_symbols_are_exported():
    from genarchbench_b200 import kswv
    text = open(os.path.join(ROOT, "include", "kswv_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(kswv_gpu_\w+)\s*\(", text)))
    assert set(names) == set(kswv.EXPORTS), (names, kswv.EXPORTS)
    L = kswv.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/kswv_gpu.h but not exported"
    h = C.c_void_p()
    assert L.kswv_gpu_init(None, 1, C.byref(h)) == 1                                   # BSW_ERR_ARG
    bad = kswv.Params(6, 0, 6, 1, 1, 4)
    assert L.kswv_gpu_init(C.byref(bad), 1, C.byref(h)) == 1
    assert L.kswv_gpu_batch(None, None, None, None, 0, None) == 1


def test_kswv_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from genarchbench_b200 import kswv
    with pytest.raises(bsw.BswError) as e:
        kswv.Kswv()
    assert e.value.code == 2                                                            # BSW_ERR_NO_DEVICE
