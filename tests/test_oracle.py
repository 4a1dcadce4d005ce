"""The oracle (oracle/bsw_oracle.c) against the reference: committed golden vectors (outputs of the
unmodified reference kernel) everywhere, and the compiled reference itself where oracle/_ref exists."""
import numpy as np
import pytest

import oracle
from conftest import GOLDEN_NAMES, assert_same_outputs, load_golden
from genarchbench_b200 import pairio


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_golden(name):
    b, w, params, want = load_golden(name)
    oracle.oracle_batch(b, w=w, params=params)
    assert_same_outputs(b.outputs(), want, b, f"oracle vs golden[{name}]")


def test_golden_covers_all_outputs_and_paths():
    b, w, params, want = load_golden("c2_16bit")
    assert (want[:, 0] > 127).mean() > 0.9, "the 16-bit fixture must exceed int8 scores"
    b, w, params, want = load_golden("edge")
    assert want.shape[1] == 6 and (want[:, 4] >= -1).all()


needs_ref = pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built / ISA unsupported")


@needs_ref
@pytest.mark.parametrize("config_id,n", [(1, 60000), (2, 8000), (4, 5000)])
def test_oracle_matches_compiled_reference(config_id, n):
    b = pairio.generate(config_id, n, seed=4242 + config_id)
    a, r = b.copy(), b.copy()
    oracle.oracle_batch(a)
    oracle.reference_batch(r)
    assert_same_outputs(a.outputs(), r.outputs(), b, f"oracle vs reference getScores16, config {config_id}")


@needs_ref
def test_reference_isa_builds_agree():
    b = pairio.generate(1, 20000, seed=99)
    outs = []
    for isa in oracle.reference_isas():
        r = b.copy()
        oracle.reference_batch(r, isa=isa)
        outs.append(r.outputs())
    for o in outs[1:]:
        assert (o == outs[0]).all()


@needs_ref
def test_vector_and_scalar_reference_agree_at_default_scoring():
    """getScores16 == scalarBandedSWA on all six fields when e_del = e_ins = 1."""
    b = pairio.generate(1, 20000, seed=100)
    v, s = b.copy(), b.copy()
    oracle.reference_batch(v)
    oracle.reference_batch(s, scalar=True)
    assert_same_outputs(v.outputs(), s.outputs(), b, "getScores16 vs scalarBandedSWA")


@needs_ref
def test_oracle_scalar_rules_match_the_reference_scalar_kernel():
    """rules = 1 of the oracle == scalarBandedSWA (bandedSWA.cpp:132-253), the kernel bwa-mem2 gives the pairs whose
    score bound leaves int16 (bwamem.cpp:2218-2228, 2384-2390): seed scores up to 50 000, non-default scoring with a
    gap-extend factor in the z-drop test, ambiguous bases scored from the matrix."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 400, 0, 50000, 0.2, 0.1
    b = pairio.generate(c, 4000, seed=123)
    for params, w in ((None, 100), (dict(o_del=5, e_del=2, o_ins=7, e_ins=3, zdrop=40, end_bonus=9, match=2, mismatch=3, ambig=-1), 30),
                      (dict(zdrop=0), 7)):
        a, r = b.copy(), b.copy()
        oracle.oracle_batch(a, w=w, params=params, scalar_zdrop=True)
        oracle.reference_batch(r, w=w, params=params, scalar=True)
        assert_same_outputs(a.outputs(), r.outputs(), b, f"oracle scalar rules vs scalarBandedSWA, {params}, w={w}")


def test_oracle_scalar_rule_differs_only_with_gap_extend_factor():
    """The vector z-drop rule drops the e_del/e_ins factor (bandedSWA.cpp:1889-1902): no effect at e=1."""
    b = pairio.generate(4, 3000, seed=5)
    v, s = b.copy(), b.copy()
    oracle.oracle_batch(v, params=dict(zdrop=30))
    oracle.oracle_batch(s, params=dict(zdrop=30), scalar_zdrop=True)
    assert (v.outputs() == s.outputs()).all()
    v, s = b.copy(), b.copy()
    oracle.oracle_batch(v, params=dict(zdrop=30, e_del=2, e_ins=2))
    oracle.oracle_batch(s, params=dict(zdrop=30, e_del=2, e_ins=2), scalar_zdrop=True)
    assert (v.outputs() != s.outputs()).any()


def test_cells_visited_is_below_rectangle():
    b = pairio.generate(1, 5000, seed=6)
    cells = oracle.oracle_batch(b)
    assert 0 < cells < b.cells_rect()
    assert 0.4 < cells / b.cells_rect() < 0.9   # ~0.66 on extension-shaped data


def test_empty_and_degenerate_inputs():
    b = pairio.from_sequences([([0, 1, 2], [0, 1, 2], 5), ([], [0, 1], 7), ([0, 1], [], 9)])
    oracle.oracle_batch(b)
    o = b.outputs()
    assert o[1].tolist() == [7, 0, 0, 0, -1, 0] and o[2].tolist() == [9, 0, 0, 0, -1, 0]
    assert o[0, 0] == 8   # h0 + 3 matches
