"""The kswv oracle (oracle/kswv_oracle.c) against the reference: the committed golden vectors everywhere, and the
compiled unmodified class (oracle/_ref/libkswv_ref_avx512.so) on seeded batches where the host has AVX512BW."""
import numpy as np
import pytest

from conftest import KSWV_GOLDEN_NAMES, assert_same_aln, load_kswv_golden
from oracle import kswv
from oracle.kswv import KSW_XBYTE, KSW_XSTART, KSW_XSTOP, KSW_XSUBO


def test_golden_fixtures_exist():
    assert len(KSWV_GOLDEN_NAMES) >= 8


@pytest.mark.parametrize("name", KSWV_GOLDEN_NAMES)
def test_oracle_matches_golden(name):
    pairs, ref, qer, params, want = load_kswv_golden(name)
    got, cells = kswv.oracle_batch(pairs, ref, qer, params)
    assert cells > 0
    assert_same_aln(got, want, pairs, name)


def test_golden_vectors_exercise_the_outputs():
    """The fixtures are not degenerate: second-best scores, start positions, saturated 8-bit lanes, early stops."""
    pairs, _, _, _, a = load_kswv_golden("kswv_default")
    assert (a[:, 3] > 0).mean() > 0.1 and (a[:, 5] >= 0).mean() > 0.5 and (pairs["h0"] & KSW_XBYTE).all()
    _, _, _, _, a = load_kswv_golden("kswv_saturating8")
    assert (a[:, 0] == 255).sum() > 50 and (a[a[:, 0] == 255][:, 3:5] == -1).all()
    pairs, _, _, _, a = load_kswv_golden("kswv_16bit")
    assert not (pairs["h0"] & KSW_XBYTE).any() and a[:, 0].max() > 255
    _, _, _, _, a = load_kswv_golden("kswv_stop")
    assert (a[:, 0] >= 45).mean() > 0.5 and a[:, 0].max() < 60


CASES = [
    ("default", None, {}),
    ("16bit", None, dict(read_len=(250, 400))),
    ("mixed classes", None, dict(read_len=(200, 300))),
    ("a=2", dict(match=2, mismatch=8, o_del=12, o_ins=12, e_del=2, e_ins=2), dict(match=2, read_len=(80, 200))),
    ("forced 8-bit, saturating", None, dict(read_len=(240, 400), p_sub=0.005,
                                            xtra=lambda l: KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 19)),
    ("tiny", None, dict(read_len=(1, 24), window=(0.5, 3.0), min_seed_len=3)),
    ("no flags", None, dict(xtra=0)),
    ("start only", None, dict(xtra=KSW_XSTART)),
    ("subo only, 16-bit", None, dict(xtra=KSW_XSUBO | 30)),
    ("stop, 8-bit", None, dict(xtra=KSW_XSTOP | KSW_XBYTE | 40)),
    ("stop + start, 16-bit", None, dict(xtra=KSW_XSTOP | KSW_XSTART | 60)),
    ("ambiguous bases", None, dict(p_n=0.2)),
    ("stop at zero, 16-bit", None, dict(read_len=(1, 200), window=(0.3, 3.0), xtra=KSW_XSTOP | KSW_XSTART)),
    ("stop at zero, 8-bit", None, dict(read_len=(1, 200), window=(0.3, 3.0), xtra=KSW_XSTOP | KSW_XSUBO | KSW_XSTART | KSW_XBYTE)),
    ("indel-rich", None, dict(p_indel=0.05, p_sub=0.1)),
    ("asymmetric gaps", dict(o_del=4, e_del=2, o_ins=7, e_ins=1), {}),
    ("minsc above every score", None, dict(xtra=KSW_XSUBO | KSW_XSTART | 300)),
    ("minsc beyond 8 bits", None, dict(xtra=KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 300)),
]


@pytest.mark.skipif(not kswv.reference_available(), reason="needs oracle/_ref/libkswv_ref_avx512.so and AVX512BW")
@pytest.mark.parametrize("what,params,kw", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_compiled_reference(what, params, kw):
    for seed in (1, 2):
        pairs, ref, qer = kswv.make_workload(300, seed=seed, **kw)
        got, _ = kswv.oracle_batch(pairs, ref, qer, params)
        want = kswv.reference_batch(pairs, ref, qer, params)
        assert_same_aln(got, want, pairs, what)


@pytest.mark.skipif(not kswv.reference_available(), reason="needs oracle/_ref/libkswv_ref_avx512.so and AVX512BW")
def test_a_lane_does_not_depend_on_its_batch_mates():
    """The restatement is per pair; the reference runs 64 / 32 pairs in lock step. Shuffling the batch must not
    change any pair's result (regid keeps the output slot)."""
    pairs, ref, qer = kswv.make_workload(500, seed=5, read_len=(60, 300))
    want = kswv.reference_batch(pairs, ref, qer)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(pairs))
    got = kswv.reference_batch(pairs[perm].copy(), ref, qer)
    assert_same_aln(got, want, pairs, "shuffled batch")


KINDS = ("homopolymer", "two-letter", "tandem", "identical-prefix")


@pytest.mark.skipif(not kswv.reference_available(), reason="needs oracle/_ref/libkswv_ref_avx512.so and AVX512BW")
@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_compiled_reference_on_ties(kind):
    """Low-complexity sequences: many equal row maxima and equal scores, where first-column / first-row rules and the
    rising-row filter decide the outputs."""
    for seed in (1, 2):
        pairs, ref, qer = kswv.make_low_complexity(400, seed=seed, kind=kind)
        got, _ = kswv.oracle_batch(pairs, ref, qer)
        assert_same_aln(got, kswv.reference_batch(pairs, ref, qer), pairs, kind)
