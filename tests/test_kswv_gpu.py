"""kswv on the B200 through the C ABI (kswv_gpu_batch) against the golden vectors (compiled reference) and the
oracle (pinned to it): bit-exact on all seven kswr_t fields."""
import numpy as np
import pytest

from conftest import KSWV_GOLDEN_NAMES, assert_same_aln, load_kswv_golden
from oracle import kswv as okswv
from oracle.kswv import KSW_XBYTE, KSW_XSTART, KSW_XSTOP, KSW_XSUBO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from genarchbench_b200 import kswv
    g = kswv.Kswv()
    yield g
    g.close()


def handle_for(params):
    from genarchbench_b200 import kswv
    p = dict(okswv.DEFAULT_PARAMS)
    p.update(params or {})
    return kswv.Kswv(p["o_del"], p["e_del"], p["o_ins"], p["e_ins"], p["match"], p["mismatch"])


@pytest.mark.parametrize("name", KSWV_GOLDEN_NAMES)
def test_matches_golden(name):
    pairs, ref, qer, params, want = load_kswv_golden(name)
    g = handle_for(params)
    try:
        assert_same_aln(g.align(pairs, ref, qer), want, pairs, f"GPU vs golden[{name}]")
    finally:
        g.close()


CASES = [
    ("mate rescue, 151 bp reads", dict(n=6000, read_len=(100, 151))),
    ("both classes mixed", dict(n=3000, read_len=(200, 300))),
    ("every strip width", dict(n=6000, read_len=(1, 300), window=(0.3, 4.0), min_seed_len=5)),
    ("several passes (above 256 columns)", dict(n=400, read_len=(257, 900))),
    ("several passes, 8-bit class", dict(n=400, read_len=(257, 600), p_sub=0.3,
                                         xtra=lambda l: KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 19)),
    ("saturating 8-bit", dict(n=2000, read_len=(240, 330), p_sub=0.005,
                              xtra=lambda l: KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 19)),
    ("stop + start", dict(n=2000, xtra=lambda l: KSW_XSTOP | KSW_XSTART | (KSW_XBYTE if l < 120 else 0) | 45)),
    ("no flags", dict(n=1000, xtra=0)),
    ("stop at zero: the first row ends the pair", dict(n=2000, read_len=(1, 300), window=(0.3, 3.0), xtra=lambda l: KSW_XSTOP | KSW_XSTART)),
    ("ambiguous bases", dict(n=2000, p_n=0.2)),
    ("long windows", dict(n=300, read_len=(100, 151), window=(20.0, 60.0))),
]


@pytest.mark.parametrize("what,kw", CASES, ids=[c[0] for c in CASES])
def test_matches_oracle(dev, what, kw):
    pairs, ref, qer = okswv.make_workload(seed=31, **kw)
    want, cells = okswv.oracle_batch(pairs, ref, qer)
    got = dev.align(pairs, ref, qer)
    assert_same_aln(got, want, pairs, what)
    st = dev.stats()
    assert st["cells"] == cells and st["pairs"] == len(pairs) and st["kernel_launches"] >= 1


def test_nondefault_scoring():
    params = dict(match=2, mismatch=5, o_del=4, e_del=2, o_ins=7, e_ins=1)
    pairs, ref, qer = okswv.make_workload(3000, seed=32, match=2, read_len=(60, 180), p_indel=0.03)
    want, _ = okswv.oracle_batch(pairs, ref, qer, params)
    g = handle_for(params)
    try:
        assert_same_aln(g.align(pairs, ref, qer), want, pairs, "a=2, asymmetric gaps")
    finally:
        g.close()


def test_many_chunks_shuffled_order_and_regid(dev):
    """More pairs than one chunk, in shuffled order (so the sequences are gathered, not sent as a range), with
    regid a permutation: every result lands in aln[regid]."""
    pairs, ref, qer = okswv.make_workload(80000, seed=33, read_len=(20, 60), window=(1.0, 3.0), min_seed_len=5)
    rng = np.random.default_rng(1)
    want, _ = okswv.oracle_batch(pairs, ref, qer)
    perm = rng.permutation(len(pairs))
    shuffled = pairs[perm].copy()
    got = dev.align(shuffled, ref, qer)
    assert_same_aln(got, want, pairs, "shuffled order")
    st = dev.stats()
    assert st["chunks"] >= 3 and st["gathered"] == st["chunks"]
    got = dev.align(pairs, ref, qer)            # dense, in order: sent as ranges (numpy memory is pageable: staged)
    assert_same_aln(got, want, pairs, "dense order")
    st = dev.stats()
    assert st["gathered"] == 0 and st["staged"] == st["chunks"]
    # the same from page-locked buffers: DMA'd in place
    import ctypes as C
    from genarchbench_b200 import bsw
    L = bsw.lib()

    def pinned_copy(a):
        p = L.bsw_gpu_host_alloc(a.nbytes + 64)
        v = np.ctypeslib.as_array((C.c_uint8 * a.nbytes).from_address(p)).view(a.dtype)
        v[:] = a
        return v, p
    (pref, p1), (pqer, p2) = pinned_copy(ref), pinned_copy(qer)
    got = dev.align(pairs, pref, pqer)
    assert_same_aln(got, want, pairs, "page-locked buffers")
    st = dev.stats()
    assert st["gathered"] == 0 and st["staged"] == 0
    L.bsw_gpu_host_free(p1); L.bsw_gpu_host_free(p2)


def test_empty_sequences_and_empty_batch(dev):
    pairs, ref, qer = okswv.make_workload(400, seed=34, read_len=(5, 40), min_seed_len=3)
    pairs["len1"][::4] = 0
    pairs["len2"][1::4] = 0
    want, _ = okswv.oracle_batch(pairs, ref, qer)
    assert_same_aln(dev.align(pairs, ref, qer), want, pairs, "empty reference / query")
    assert dev.align(pairs[:0].copy(), ref, qer).shape == (0, 7)


def test_domain_errors(dev):
    from genarchbench_b200 import bsw
    pairs, ref, qer = okswv.make_workload(10, seed=35)
    bad = pairs.copy()
    bad["len1"][3] = 40000
    with pytest.raises(bsw.BswError) as e:
        dev.align(bad, ref, qer)
    assert e.value.code == 5                                                   # BSW_ERR_RANGE
    bad = pairs.copy()
    bad["regid"][2] = 10
    with pytest.raises(bsw.BswError) as e:
        dev.align(bad, ref, qer)
    assert e.value.code == 1                                                   # BSW_ERR_ARG
    want, _ = okswv.oracle_batch(pairs, ref, qer)
    assert_same_aln(dev.align(pairs, ref, qer), want, pairs, "after rejected calls")


def test_chunks_over_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from genarchbench_b200 import kswv
    pairs, ref, qer = okswv.make_workload(70000, seed=36, read_len=(20, 60), window=(1.0, 3.0), min_seed_len=5)
    want, _ = okswv.oracle_batch(pairs, ref, qer)
    g = kswv.Kswv(n_gpus=2)
    try:
        assert_same_aln(g.align(pairs, ref, qer), want, pairs, "two GPUs")
        st = g.stats()
        assert st["n_gpus"] == 2 and st["chunks"] >= 2
        small = pairs[:5].copy()
        assert_same_aln(g.align(small, ref, qer), want[:5], small, "five pairs on a two-GPU handle")
    finally:
        g.close()


def test_random_call_sequences_on_one_handle(dev):
    """Twenty calls of random size (0 .. 70k pairs), order (dense / shuffled), read lengths and flags on one handle:
    slots, scratch and the phase-1 ordering buffers grow and are reused; every call against the oracle."""
    rng = np.random.default_rng(77)
    lens = [(1, 24), (20, 60), (100, 151), (200, 300), (257, 500)]
    for call in range(20):
        lo, hi = lens[int(rng.integers(0, len(lens)))]
        big = hi <= 60
        n = int(rng.choice([0, 1, 2, 33, 500, 3000] + ([40000, 70000] if big else [])))
        if n == 0:
            pairs, ref, qer = okswv.make_workload(1, seed=call)
            assert dev.align(pairs[:0].copy(), ref, qer).shape == (0, 7)
            continue
        flags = [None, 0, KSW_XSTART, KSW_XSUBO | 25, lambda l: KSW_XSTOP | KSW_XSTART | (KSW_XBYTE if l < 200 else 0) | 30]
        pairs, ref, qer = okswv.make_workload(n, seed=100 + call, read_len=(lo, hi), window=(0.5, 3.0), min_seed_len=3,
                                              xtra=flags[int(rng.integers(0, len(flags)))])
        want, _ = okswv.oracle_batch(pairs, ref, qer)
        if rng.random() < 0.5:
            pairs = pairs[rng.permutation(n)].copy()
        assert_same_aln(dev.align(pairs, ref, qer), want, None, f"call {call}: n={n} reads {lo}-{hi}")


@pytest.mark.parametrize("kind", ("homopolymer", "two-letter", "tandem", "identical-prefix"))
def test_ties(dev, kind):
    """Low-complexity sequences: equal row maxima and scores everywhere."""
    pairs, ref, qer = okswv.make_low_complexity(4000, seed=5, kind=kind)
    want, _ = okswv.oracle_batch(pairs, ref, qer)
    assert_same_aln(dev.align(pairs, ref, qer), want, pairs, kind)


def test_full_size_properties(dev):
    """The bench's full-size batch (400 000 pairs; the oracle is only sampled there): properties every result must
    have whatever the size -- ends inside the sequences, starts before the ends, a score the aligned span can pay
    for, a second best that never beats the best, and the tiled batch repeating its base batch exactly."""
    base = 20000
    pairs0, ref0, qer0 = okswv.make_workload(base, seed=7, read_len=(151, 151))
    reps = 20
    rb, qb = int(pairs0["idr"][-1] + pairs0["len1"][-1]), int(pairs0["idq"][-1] + pairs0["len2"][-1])
    pairs = np.tile(pairs0, reps)
    ref = np.concatenate([ref0[:rb]] * reps + [np.zeros(64, np.uint8)])
    qer = np.concatenate([qer0[:qb]] * reps + [np.zeros(64, np.uint8)])
    for r in range(reps):
        sl = slice(r * base, (r + 1) * base)
        pairs["idr"][sl] += r * rb
        pairs["idq"][sl] += r * qb
        pairs["regid"][sl] += r * base
    a = dev.align(pairs, ref, qer)
    score, te, qe, score2, te2, tb, qb_ = (a[:, k] for k in range(7))
    hit = score > 0
    assert (te[hit] >= 0).all() and (te < pairs["len1"]).all() and (qe[hit] < pairs["len2"][hit]).all()
    assert (te[~hit] == -1).all() and (qe[~hit] == 0).all()
    started = tb >= 0
    assert (tb[started] <= te[started]).all() and (qb_[started] >= 0).all() and (qb_[started] <= qe[started]).all()
    span = np.minimum(te - tb + 1, qe - qb_ + 1)
    assert (score[started] <= span[started]).all()                       # match = 1: a span of s columns pays at most s
    assert ((score2 == -1) | (score2 <= score)).all() and ((score2 == -1) == (te2 == -1)).all()
    assert (np.abs(te2[score2 > 0] - te[score2 > 0]) >= 1).all()
    # a checksum of checksums: every tile equals the first, and the first equals the oracle
    tiles = a.reshape(reps, base, 7)
    assert (tiles == tiles[0]).all()
    want, _ = okswv.oracle_batch(pairs0, ref0, qer0)
    assert_same_aln(tiles[0], want, pairs0, "first tile")
