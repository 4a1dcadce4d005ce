"""Parity tests proper: the CUDA path, called through the C ABI (libbsw_gpu.so), against the golden
vectors (reference outputs) and the oracle. Bit-exact on all six outputs. Run with -m gpu on a B200."""
import numpy as np
import pytest

import oracle
from conftest import GOLDEN_NAMES, assert_same_outputs, load_golden
from genarchbench_b200 import bsw, pairio

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_gpu_matches_golden(name):
    b, w, params, want = load_golden(name)
    with bsw.BswGpu(**params) as g:
        g.batch(b.pairs, b.ref, b.qer, w)
    assert_same_outputs(b.outputs(), want, b, f"GPU vs golden[{name}]")


# BASELINE.json configs 1, 2 and 4 at parity-test sizes (the oracle finishes in seconds)
@pytest.mark.parametrize("config_id,n", [(1, 100000), (2, 100000), (4, 30000)])
def test_gpu_matches_oracle_on_baseline_configs(gpu, config_id, n):
    b = pairio.generate(config_id, n)
    a = b.copy()
    oracle.oracle_batch(a)
    gpu.batch(b.pairs, b.ref, b.qer, 100)
    assert_same_outputs(b.outputs(), a.outputs(), b, f"GPU vs oracle, config {config_id}")
    st = gpu.stats()
    assert st["kernel_launches"] > 0 and st["pairs"] == n
    if config_id == 2:
        assert (a.outputs()[:, 0] > 127).mean() > 0.9      # the path int8 lanes could not hold


def test_reference_interface_mirror():
    """Reads like the reference driver: construct once, getScores16 per batch (main_banded.cpp:271-276,345)."""
    b = pairio.generate(1, 5000, seed=31)
    a = b.copy()
    oracle.oracle_batch(a)
    sw = bsw.BandedPairWiseSW(6, 1, 6, 1, 100, 5, None, 1, 4, 1)
    for lo in range(0, len(b), 512):                          # -b 512 as in regression_small.sh
        part = b.pairs[lo:lo + 512]
        sw.getScores16(part, b.ref, b.qer, len(part), 1, 100)
    sw.close()
    assert_same_outputs(b.outputs(), a.outputs(), b, "mirror class, 512-pair batches")


@pytest.mark.parametrize("w", [0, 1, 3, 10, 30, 200])
def test_band_widths(gpu, w):
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 300, 0, 80, 0.3, 0.2
    b = pairio.generate(c, 20000, seed=500 + w)
    a = b.copy()
    oracle.oracle_batch(a, w=w)
    gpu.batch(b.pairs, b.ref, b.qer, w)
    assert_same_outputs(b.outputs(), a.outputs(), b, f"w={w}")


@pytest.mark.parametrize("w", [2, 7, 40, 100, 500])
def test_long_pairs_warp_kernel(gpu, w):
    """Queries of 300..1500 bases run the warp-per-pair kernel (tiles, max-plus scan of F, warp
    reductions for the row decisions): narrow and wide bands, ambiguous bases, unrelated pairs."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 300, 1500, 0, 200, 0.3, 0.15
    b = pairio.generate(c, 3000, seed=900 + w)
    a = b.copy()
    cells = oracle.oracle_batch(a, w=w)
    gpu.batch(b.pairs, b.ref, b.qer, w)
    assert gpu.stats()["pairs_long"] > 1000
    assert_same_outputs(b.outputs(), a.outputs(), b, f"long pairs, w={w}")
    gpu.stage(b.pairs, b.ref, b.qer, w)
    assert gpu.count_staged() == cells                        # the warp kernel's beg is the reference's


def test_long_pairs_nondefault_scoring():
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac = 350, 900, 0, 100, 0.2
    b = pairio.generate(c, 1500, seed=77)
    for params in (dict(o_del=5, e_del=2, o_ins=7, e_ins=1, zdrop=40, end_bonus=9, match=2, mismatch=3, ambig=-1),
                   dict(o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=0, end_bonus=5, match=1, mismatch=4, ambig=-1)):
        a = b.copy(); e = b.copy()
        oracle.oracle_batch(a, w=60, params=params)
        with bsw.BswGpu(**params) as g:
            g.batch(e.pairs, e.ref, e.qer, 60)
        assert_same_outputs(e.outputs(), a.outputs(), e, f"long pairs, {params}")


def test_large_seed_scores_take_the_general_m_path(gpu):
    """h0 close to the int16 limit: score * (match + 1) no longer fits, so the launches use the general
    M = Hd ? Hd + s : 0 instructions (FASTM off) -- thread-per-pair, windowed and warp kernels."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac = 20, 900, 20000, 31000, 0.2
    b = pairio.generate(c, 4000, seed=31)
    for w in (100, 300):
        a = b.copy(); e = b.copy()
        oracle.oracle_batch(a, w=w)
        gpu.batch(e.pairs, e.ref, e.qer, w)
        assert_same_outputs(e.outputs(), a.outputs(), e, f"large h0, w={w}")


@pytest.mark.parametrize("h0_max,all_keyed", [(1985, True), (1995, False), (40000, False)])
def test_keyed_argmax_at_the_limits_of_its_key(gpu, h0_max, all_keyed):
    """extend_pair<.., KEY>: launches whose scores and group indices share 16 bits take the row argmax as
    one unsigned lane maximum of score << kbits | group. Queries <= 60 bases need 5 index bits, so scores up
    to 2047 are keyed: h0 <= 1985 keeps every launch below, h0 <= 1995 pushes the longest bin above (general
    argmax for that launch only), a huge h0 turns the keyed path off altogether."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 60, 0, min(h0_max, 30000), 0.2, 0.1
    b = pairio.generate(c, 40000, seed=77)
    a = b.copy()
    oracle.oracle_batch(a)
    gpu.batch(b.pairs, b.ref, b.qer, 100)
    assert_same_outputs(b.outputs(), a.outputs(), b, f"keyed argmax, h0 <= {h0_max}")
    st = gpu.stats()
    if all_keyed:
        assert st["pairs_keyed"] == st["pairs_short"] > 0
    elif h0_max < 30000:
        assert 0 < st["pairs_keyed"] < st["pairs_short"]
    else:
        assert st["pairs_keyed"] == 0


def test_huge_band_argument(gpu):
    """w far beyond any sequence length (the per-pair band clamps it, bandedSWA.cpp:2898-2919)."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max = 100, 700
    b = pairio.generate(c, 3000, seed=21)
    a = b.copy()
    oracle.oracle_batch(a, w=1 << 29)
    gpu.batch(b.pairs, b.ref, b.qer, 1 << 29)
    assert_same_outputs(b.outputs(), a.outputs(), b, "w = 2^29")


def test_very_long_pair(gpu):
    """One pair near the kernel's limits: 20k x 30k bases, a single warp with a 150 KB row."""
    rng = np.random.default_rng(5)
    q = rng.integers(0, 4, 20000).astype(np.uint8)
    t = np.concatenate([q[:15000], rng.integers(0, 4, 15000).astype(np.uint8)])
    t[::97] = (t[::97] + 1) & 3
    b = pairio.from_sequences([(t, q, 100), (t[:5000], q[:4000], 50)])
    a = b.copy()
    oracle.oracle_batch(a)
    gpu.batch(b.pairs, b.ref, b.qer, 100)
    assert_same_outputs(b.outputs(), a.outputs(), b, "very long pair")


def test_band_doubling_retry_matches_reference_loop(gpu):
    """bsw_gpu_batch_retry == the production caller's loop around getScores16 (bwamem.cpp:2448-2508),
    replayed with the oracle: final outputs and the number of tries per pair."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.indel_rate = 40, 400, 10, 80, 0.05
    b = pairio.generate(c, 20000, seed=4242)
    w, max_tries = 8, 3
    want = b.copy()
    tries_want = np.zeros(len(b), dtype=np.int32)
    active = np.arange(len(b))
    prev = np.full(len(b), -1, dtype=np.int64)
    for t in range(max_tries):
        wt = w << t
        sub = pairio.PairBatch(want.pairs[active].copy(), b.ref, b.qer)
        oracle.oracle_batch(sub, w=wt)
        for f in pairio.OUTPUT_FIELDS:
            want.pairs[f][active] = sub.pairs[f]
        tries_want[active] = t + 1
        final = (sub.pairs["score"] == prev[active]) | (sub.pairs["max_off"] < (wt >> 1) + (wt >> 2))
        prev[active] = sub.pairs["score"]
        active = active[~final]
        if len(active) == 0:
            break
    tries = gpu.batch_retry(b.pairs, b.ref, b.qer, w, max_tries)
    assert_same_outputs(b.outputs(), want.outputs(), b, "band-doubling retry")
    assert (tries == tries_want).all() and tries.max() == max_tries and (tries == 1).any()


def test_empty_single_and_ragged(gpu):
    e = pairio.from_sequences([([0], [0], 1)])
    gpu.batch(e.pairs[:0], e.ref, e.qer, 100)                  # n = 0 is a no-op
    gpu.batch(e.pairs, e.ref, e.qer, 100)
    assert e.outputs()[0].tolist() == [2, 1, 1, 1, 2, 0]
    rng = np.random.default_rng(3)
    items = [(rng.integers(0, 5, l1).astype(np.uint8), rng.integers(0, 5, l2).astype(np.uint8), h0)
             for l1, l2, h0 in [(0, 5, 3), (5, 0, 4), (1, 1, 0), (1, 900, 50), (900, 1, 50), (3, 2, 127),
                                (129, 128, 60), (128, 127, 60), (2000, 1500, 200), (31, 32, 9), (33, 31, 9)]]
    b = pairio.from_sequences(items)
    a = b.copy()
    oracle.oracle_batch(a)
    gpu.batch(b.pairs, b.ref, b.qer, 100)
    assert_same_outputs(b.outputs(), a.outputs(), b, "ragged")


def test_empty_sequences_spread_over_many_slabs(gpu):
    """Pairs with an empty target or query are answered without a DP launch (bandedSWA.cpp:181 never runs its
    loop). A streaming call of several slabs scatters a slab's results while the next one is packed: the
    answers of the empty pairs must survive that deferred scatter (they did not: ADVICE r1, claim_slab)."""
    b = pairio.generate(1, 700_000, seed=41)                   # the taper cuts this into >= 3 slabs
    rng = np.random.default_rng(5)
    k1 = rng.choice(len(b), 3000, replace=False)
    b.pairs["len1"][k1[:1500]] = 0
    b.pairs["len2"][k1[1500:]] = 0
    a = b.copy()
    oracle.oracle_batch(a)
    for _ in range(2):                                         # second call: ring slots hold older results
        for f in pairio.OUTPUT_FIELDS:
            b.pairs[f] = -1
        gpu.batch(b.pairs, b.ref, b.qer, 100)
        assert_same_outputs(b.outputs(), a.outputs(), b, "empty sequences over several slabs")
    want = np.array([[h0, 0, 0, 0, -1, 0] for h0 in b.pairs["h0"][k1]])
    assert (b.outputs()[k1] == want).all()


def test_blob_capacity_retry_with_many_ambiguous_pairs():
    """Half of the pairs hold an ambiguous base, so their 4-bit copies triple the packed bytes and the first
    capacity guess (sampled lengths + 15 %) is short: prepare_slab reports the exact need and runs again
    (the reported size used to be inflated by an arena per remaining pair)."""
    c = pairio.preset(1)
    c.n_frac = 0.5
    b = pairio.generate(c, 400_000, seed=43)
    a = b.copy()
    oracle.oracle_batch(a)
    with bsw.BswGpu() as g:                                    # fresh handle: ring slots sized by this call
        g.batch(b.pairs, b.ref, b.qer, 100)
    assert_same_outputs(b.outputs(), a.outputs(), b, "capacity retry")


def test_growing_slabs_do_not_lose_queued_results():
    """Early slabs cut by the base-count limit hold few (long) pairs, later slabs of the same call hold many
    short ones: the ring slot's pair-sized buffers grow while a scatter job of the older slab is still queued
    (use-after-free of h_out)."""
    cl = pairio.preset(4)
    cl.len2_min, cl.len2_max = 900, 1000
    long_ = pairio.generate(cl, 300_000, seed=44)              # > 512 Mi bases: the first slab is cut by bases
    short = pairio.generate(1, 2_600_000, seed=45)             # slots 1, 2, then slot 0 again with 4x the pairs
    b = pairio.concat([long_, short])
    with bsw.BswGpu() as g:
        g.batch(b.pairs, b.ref, b.qer, 100)
    idx = np.random.default_rng(6).choice(len(b), 40000, replace=False)
    idx = np.concatenate([idx, np.arange(0, 300_000, 97)])
    samp = pairio.PairBatch(b.pairs[idx].copy(), b.ref, b.qer)
    got = samp.outputs()
    oracle.oracle_batch(samp)
    assert_same_outputs(got, samp.outputs(), samp, "growing slabs, sampled")


def test_untouched_fields_and_padding(gpu):
    """Only the six outputs of entries < n are written (the reference also clobbers n..round)."""
    b = pairio.generate(1, 1000, seed=8)
    before = b.pairs.copy()
    gpu.batch(b.pairs, b.ref, b.qer, 100, n=900)
    for f in ("idr", "idq", "id", "len1", "len2", "h0", "seqid", "regid"):
        assert (b.pairs[f] == before[f]).all()
    assert (b.pairs["score"][900:] == -1).all() and (b.pairs["score"][:900] >= 0).all()


def test_scalar_class_pairs_are_computed_not_rejected(gpu):
    """A pair whose score bound h0 + len2 * match leaves int16 belongs to bwa-mem2's scalar class (bwamem.cpp:
    2218-2228) and runs, as there (:2384-2390), by the rules of scalarBandedSWA in int32 -- the call does not fail.
    Replayed with the oracle: scalar rules for that class, the vector rules (getScores16) for everything else."""
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.h0_min, c.h0_max, c.n_frac, c.random_frac = 1, 600, 0, 200, 0.2, 0.1
    b = pairio.generate(c, 300_000, seed=21)                   # several slabs (taper)
    rng = np.random.default_rng(8)
    big = np.sort(rng.choice(len(b), 400, replace=False))
    b.pairs["h0"][big] = rng.integers(32768, 60000, len(big))
    b.pairs["h0"][big[:50]] = 32767 - b.pairs["len2"][big[:50]] + 1       # just beyond the bound
    edge = big[50:80]
    b.pairs["h0"][edge] = 32767 - b.pairs["len2"][edge]                    # exactly on it: still the int16 kernels
    is_big = b.pairs["h0"].astype(np.int64) + np.minimum(b.pairs["len1"], b.pairs["len2"]) > 32767
    assert (bsw.classify(b.pairs, 1)[1] == 2).sum() == is_big.sum()
    assert is_big.sum() == len(big) - len(edge)
    want = b.copy()
    oracle.oracle_batch(want)
    sc = pairio.PairBatch(b.pairs[is_big].copy(), b.ref, b.qer)
    oracle.oracle_batch(sc, scalar_zdrop=True)
    wo = want.outputs()
    wo[is_big] = sc.outputs()
    gpu.batch(b.pairs, b.ref, b.qer, 100)
    assert_same_outputs(b.outputs(), wo, b, "batch with scalar-class pairs")
    st = gpu.stats()
    assert st["pairs_scalar"] == int(is_big.sum()) and st["pairs_invalid"] == 0
    assert b.pairs["score"][is_big].max() > 40000               # beyond what an int16 lane could hold


def test_invalid_record_does_not_poison_the_batch(gpu):
    """One record outside the domain (negative length) in the middle of a multi-slab batch: every other pair is
    computed, the record's outputs are -1, the call reports BSW_ERR_RANGE and where."""
    b = pairio.generate(1, 600_000, seed=22)
    a = b.copy()
    oracle.oracle_batch(a)
    k = 333_333
    b.pairs["len1"][k] = -5
    b.pairs["h0"][k + 1000] = -1
    with pytest.raises(bsw.BswError) as e:
        gpu.batch(b.pairs, b.ref, b.qer, 100)
    assert e.value.code == 5
    st = gpu.stats()
    assert st["pairs_invalid"] == 2 and st["first_invalid"] == k
    ok = np.ones(len(b), bool)
    ok[[k, k + 1000]] = False
    assert (b.outputs()[~ok] == -1).all()
    assert_same_outputs(b.outputs()[ok], a.outputs()[ok], None, "valid pairs around an invalid record")
    gpu.batch(a.pairs, a.ref, a.qer, 100)                       # the handle is fine afterwards


def test_reserve_presizes_the_rings():
    b = pairio.generate(1, 300000, seed=2)
    a = b.copy()
    oracle.oracle_batch(a)
    with bsw.BswGpu() as g:
        g.reserve(len(b), int(b.pairs["len1"].sum() + b.pairs["len2"].sum()))
        g.batch(b.pairs, b.ref, b.qer, 100)
        assert g.stats()["host_alloc_ms"] < 5.0               # nothing left to allocate in the call
    assert_same_outputs(b.outputs(), a.outputs(), b, "after reserve")


def test_staged_api_matches_batch(gpu):
    b = pairio.generate(1, 200000, seed=17)
    s = b.copy()
    gpu.batch(b.pairs, b.ref, b.qer, 100)
    gpu.stage(s.pairs, s.ref, s.qer, 100)
    ms = gpu.run_staged()
    ms2 = gpu.run_staged()                                     # re-runnable on resident data
    assert ms > 0 and ms2 > 0
    gpu.fetch_staged(s.pairs)
    assert (b.outputs() == s.outputs()).all()


@pytest.mark.parametrize("config_id,n,w", [(1, 50000, 100), (2, 10000, 100), (4, 10000, 100), (4, 10000, 7)])
def test_cell_count_matches_oracle(gpu, config_id, n, w):
    """The GCUPS unit of work: the COUNT kernel's visited cells == the oracle's inner-loop count."""
    b = pairio.generate(config_id, n, seed=60 + config_id)
    a = b.copy()
    cells = oracle.oracle_batch(a, w=w)
    gpu.stage(b.pairs, b.ref, b.qer, w)
    assert gpu.count_staged() == cells
    gpu.fetch_staged(b.pairs)                                   # the COUNT variant's outputs are exact too
    assert_same_outputs(b.outputs(), a.outputs(), b, "COUNT kernel outputs")


def test_order_and_batching_invariance(gpu):
    """Per-pair results do not depend on neighbours, order or slab boundaries."""
    b = pairio.generate(4, 40000, seed=23)
    gpu.batch(b.pairs, b.ref, b.qer, 100)
    perm = np.random.default_rng(1).permutation(len(b))
    p = pairio.PairBatch(b.pairs[perm].copy(), b.ref, b.qer)
    for f in pairio.OUTPUT_FIELDS:
        p.pairs[f] = -1
    gpu.batch(p.pairs, p.ref, p.qer, 100)
    assert (p.outputs() == b.outputs()[perm]).all()


def test_full_size_properties_config3_sample(gpu):
    """BASELINE config 3 shape at 4M pairs (multi-slab): permutation-invariant checksum, duplicate
    pairs agree, and a random sample is bit-exact against the oracle."""
    n = 4_000_000
    b = pairio.generate(3, n)
    # duplicate the first 1000 pairs at the end: same inputs, far-apart slabs
    b.pairs[-1000:] = b.pairs[:1000]
    gpu.batch(b.pairs, b.ref, b.qer, 100)
    out = b.outputs()
    assert (out[-1000:] == out[:1000]).all()
    idx = np.random.default_rng(2).choice(n, 50000, replace=False)
    samp = pairio.PairBatch(b.pairs[idx].copy(), b.ref, b.qer)
    oracle.oracle_batch(samp)
    assert_same_outputs(out[idx], samp.outputs(), samp, "4M-pair run, sampled")
    assert gpu.stats()["pairs"] == n


def test_bsw_main_driver_prints_reference_style_scores(tmp_path):
    """The C++ driver (reference CLI, pair-file loader, ROI timer, score writer; main_banded.cpp) on top of
    the C ABI: its "[i] score=" lines are what scripts/regression_small.sh:89-96 diffs against the golden
    file, its "Overall SW cycles" line is what the script greps as kernel time."""
    import os
    import re
    import subprocess
    from conftest import ROOT
    b = pairio.generate(1, 20000, seed=5)
    a = b.copy()
    oracle.oracle_batch(a)
    path = str(tmp_path / "pairs.txt")
    pairio.write_text(path, b)
    exe = os.path.join(ROOT, "genarchbench_b200", "bin", "bsw_main")
    r = subprocess.run([exe, "-pairs", path, "-t", "4", "-b", "512"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    scores = [int(m.group(2)) for m in re.finditer(r"^\[(\d+)\] score=(-?\d+)$", r.stderr, re.M)]
    assert scores == a.pairs["score"].tolist()
    assert re.search(r"Overall SW cycles = \d+, [0-9.]+ s", r.stdout) and "Total Pairs processed: 20000" in r.stdout


def test_in_process_multi_gpu_split():
    """bsw_gpu_init(params, n_gpus): slabs go round-robin over the GPUs of one process (host split and
    gather, no collective). Needs at least two devices."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("one GPU visible")
    b = pairio.generate(3, 3_000_000, seed=3)
    b.pairs[-500:] = b.pairs[:500]
    with bsw.BswGpu(n_gpus=min(ndev, 4)) as g:
        g.batch(b.pairs, b.ref, b.qer, 100)
        st = g.stats()
        assert st["n_gpus"] == min(ndev, 4) and st["pairs"] == len(b)
        out = b.outputs().copy()
        g.stage(b.pairs, b.ref, b.qer, 100)
        assert g.run_staged() > 0
        for f in pairio.OUTPUT_FIELDS:
            b.pairs[f] = -1
        g.fetch_staged(b.pairs)
        assert (b.outputs() == out).all()
    assert (out[-500:] == out[:500]).all()
    idx = np.random.default_rng(4).choice(len(b), 30000, replace=False)
    samp = pairio.PairBatch(b.pairs[idx].copy(), b.ref, b.qer)
    oracle.oracle_batch(samp)
    assert_same_outputs(out[idx], samp.outputs(), samp, "multi-GPU split, sampled")


@pytest.mark.parametrize("pinned", [False, True])
def test_packed_entry_point_matches_oracle(gpu, pinned):
    """bsw_gpu_batch_packed: what a BSWPAIR1 file holds goes to the GPU as it is (no byte-per-base buffers, no
    SeqPair records), from pageable memory (staged through the pinned rings) and from page-locked memory (DMA'd in
    place); configs 1 / 2 / 4, ambiguous bases (4-bit blobs read directly), empty sequences, several slabs."""
    alloc = bsw.host_alloc if pinned else None
    for cfg, n, w in ((1, 1_300_000, 100), (2, 20_000, 100), (4, 30_000, 100), (4, 20_000, 7)):
        c = pairio.preset(cfg)
        c.n_frac = 0.2
        b = pairio.generate(c, n, seed=70 + cfg)
        b.pairs["len1"][5::997] = 0
        b.pairs["len2"][7::991] = 0
        rec, data = pairio.pack(b, alloc)
        out = None
        if pinned:
            out = bsw.host_alloc(len(b) * pairio.RESULT_DTYPE.itemsize).view(pairio.RESULT_DTYPE)
        res = gpu.batch_packed(rec, data, w, out)
        if n > 100_000:                                        # the oracle on a sample of the large case
            idx = np.sort(np.random.default_rng(9).choice(n, 60_000, replace=False))
            idx = np.union1d(idx, np.arange(5, n, 997)[:50])
        else:
            idx = np.arange(n)
        a = pairio.PairBatch(b.pairs[idx].copy(), b.ref, b.qer)
        oracle.oracle_batch(a, w=w)
        assert_same_outputs(bsw.results_to_outputs(res[idx]), a.outputs(), a, f"packed input, config {cfg}, w={w}")
        st = gpu.stats()
        assert st["pairs"] == n and st["kernel_launches"] > 0


def test_packed_file_through_the_driver(tmp_path):
    """bsw_main on a packed pair file: file -> page-locked memory -> bsw_gpu_batch_packed; same score lines."""
    import os
    import re
    import subprocess
    from conftest import ROOT
    b = pairio.generate(1, 30000, seed=6)
    a = b.copy()
    oracle.oracle_batch(a)
    path = str(tmp_path / "pairs.bswp")
    pairio.write_packed(path, b)
    exe = os.path.join(ROOT, "genarchbench_b200", "bin", "bsw_main")
    r = subprocess.run([exe, "-pairs", path, "-t", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    scores = [int(m.group(2)) for m in re.finditer(r"^\[(\d+)\] score=(-?\d+)$", r.stderr, re.M)]
    assert scores == a.pairs["score"].tolist()
    assert "Total Pairs processed: 30000" in r.stdout


def test_random_call_sequences_on_one_handle(gpu):
    """Thirty calls of random kind (byte buffers, packed from pageable / page-locked memory, staged, band-doubling
    retry), size (0 .. 400 k pairs: single tiny slabs up to several slabs, so ring slots are reused with smaller and
    larger contents in every order) and band on ONE handle; every result against the oracle."""
    rng = np.random.default_rng(2024)
    for step in range(30):
        kind = ("batch", "packed", "packed_pinned", "staged", "retry")[int(rng.integers(0, 5))]
        n = int(rng.choice([0, 1, 7, 513, 4097, 70_000, 150_000, 400_000]))
        w = int(rng.choice([3, 30, 100]))
        cfg = int(rng.choice([1, 4]))
        c = pairio.preset(cfg)
        c.n_frac = float(rng.choice([0.0, 0.3]))
        if cfg == 4:
            c.len2_max = 400
        b = pairio.generate(c, max(n, 1), seed=1000 + step)
        b = pairio.PairBatch(b.pairs[:n].copy(), b.ref, b.qer)
        if n > 100:
            b.pairs["len2"][3::211] = 0
        a = b.copy()
        if kind == "retry":
            continue_ok = n > 0
            if not continue_ok:
                continue
            tries = gpu.batch_retry(b.pairs, b.ref, b.qer, w, 2)
            # the same loop with the oracle (bwamem.cpp:2448-2508)
            oracle.oracle_batch(a, w=w)
            lim = (w >> 1) + (w >> 2)
            again = np.nonzero(a.pairs["max_off"] >= lim)[0]
            if len(again):
                sub = pairio.PairBatch(a.pairs[again].copy(), a.ref, a.qer)
                oracle.oracle_batch(sub, w=2 * w)
                for f in pairio.OUTPUT_FIELDS:
                    a.pairs[f][again] = sub.pairs[f]
            assert (tries[again] == 2).all()
            got = b.outputs()
        elif kind == "batch":
            gpu.batch(b.pairs, b.ref, b.qer, w)
            oracle.oracle_batch(a, w=w)
            got = b.outputs()
        elif kind == "staged":
            if n == 0:
                continue
            gpu.stage(b.pairs, b.ref, b.qer, w)
            gpu.run_staged()
            gpu.fetch_staged(b.pairs)
            oracle.oracle_batch(a, w=w)
            got = b.outputs()
        else:
            rec, data = pairio.pack(b, bsw.host_alloc if kind == "packed_pinned" else None)
            res = gpu.batch_packed(rec, data, w)
            oracle.oracle_batch(a, w=w)
            got = bsw.results_to_outputs(res)
        assert_same_outputs(got, a.outputs(), a, f"call {step}: {kind}, n={n}, w={w}, config {cfg}")


def test_two_pairs_per_thread_kernel_matches_oracle():
    """BSW_DUO2=1 routes the short bins to extend_duo2 (bsw_duo.cuh: the two DPX lanes are the same cell of two
    neighbouring pairs; an evaluated alternative, off by default). The switch is read once per process."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import numpy as np, oracle\n"
        "from genarchbench_b200 import bsw, pairio\n"
        "bad = 0\n"
        "for cfg, n, w in ((1, 200000, 100), (4, 20000, 100), (4, 20000, 7)):\n"
        "    b = pairio.generate(cfg, n, seed=90 + cfg); a = b.copy(); oracle.oracle_batch(a, w=w)\n"
        "    g = bsw.BswGpu(); g.batch(b.pairs, b.ref, b.qer, w); st = g.stats(); g.close()\n"
        "    assert st['pairs_duo'] > 0, st\n"
        "    bad += int((a.outputs() != b.outputs()).any(axis=1).sum())\n"
        "print('MISMATCHES', bad)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, BSW_DUO2="1", BSW_DUO2_MINWARPS="1"))
    assert r.returncode == 0 and "MISMATCHES 0" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])


def test_dpx_peak_is_measurable():
    v = bsw.dpx_peak(0)
    assert 5e3 < v < 1e5                                        # giga thread-instructions / s
