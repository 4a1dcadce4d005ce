"""Pair generator and the reference's 3-line text pair format (main_banded.cpp:152-206)."""
import os

import numpy as np
import pytest

from genarchbench_b200 import pairio


def test_generator_is_deterministic_and_thread_independent():
    a = pairio.generate(1, 20000, nthreads=1)
    b = pairio.generate(1, 20000, nthreads=7)
    assert (a.pairs == b.pairs).all() and (a.ref == b.ref).all() and (a.qer == b.qer).all()
    c = pairio.generate(1, 20000, seed=1)
    assert not (a.pairs["len2"] == c.pairs["len2"]).all()


def test_preset_shapes():
    c1 = pairio.generate(1, 50000)
    assert c1.pairs["len2"].min() >= 1 and c1.pairs["len2"].max() <= 132
    assert (c1.pairs["len1"] == c1.pairs["len2"] + np.minimum(np.maximum(c1.pairs["len2"] - 5, 1), 200)).all()
    h0 = c1.pairs["h0"]
    assert ((h0 >= 19) & (h0 <= 120) | (h0 <= 1)).all()
    c2 = pairio.generate(2, 5000)
    assert c2.pairs["len2"].min() >= 250 and c2.pairs["len2"].max() <= 300
    c4 = pairio.generate(4, 20000)
    assert c4.pairs["len2"].min() >= 30 and c4.pairs["len2"].max() <= 1000
    assert ((c4.pairs["len1"] - c4.pairs["len2"]) <= 200).all()
    # ambiguous bases present but rare; only codes 0..4
    assert c1.ref.max() <= 4 and c1.qer.max() <= 4
    n_amb = sum(int((c1.ref[p["idr"]:p["idr"] + p["len1"]] == 4).any() or
                    (c1.qer[p["idq"]:p["idq"] + p["len2"]] == 4).any()) for p in c1.pairs[:5000])
    assert 30 < n_amb < 200


def test_text_roundtrip(tmp_path):
    b = pairio.generate(4, 300, seed=3)
    path = str(tmp_path / "pairs.txt")
    pairio.write_text(path, b)
    with open(path) as f:
        lines = f.read().split("\n")
    assert len(lines) == 3 * len(b) + 1 and set("".join(lines[1:3])) <= set("01234")
    r = pairio.read_text(path)
    assert len(r) == len(b)
    for f in ("len1", "len2", "h0"):
        assert (r.pairs[f] == b.pairs[f]).all()
    for k in (0, 17, len(b) - 1):
        p, q = b.pairs[k], r.pairs[k]
        assert (b.ref[p["idr"]:p["idr"] + p["len1"]] == r.ref[q["idr"]:q["idr"] + q["len1"]]).all()
        assert (b.qer[p["idq"]:p["idq"] + p["len2"]] == r.qer[q["idq"]:q["idq"] + q["len2"]]).all()
    assert (r.pairs["score"] == -1).all()   # outputs initialised like loadPairs (main_banded.cpp:200-201)


def test_seqpair_layout_matches_reference_struct():
    d = pairio.SEQPAIR_DTYPE
    assert d.itemsize == 72
    want = dict(idr=0, idq=8, id=16, len1=24, len2=28, h0=32, seqid=36, regid=40, score=44, tle=48,
                gtle=52, qle=56, gscore=60, max_off=64)     # SURVEY 8a row 1 (offsetof-verified)
    assert {k: d.fields[k][1] for k in want} == want


def test_text_and_packed_files_round_trip(tmp_path):
    """The reference's 3-line text format (parallel reader) and the packed binary format hold the same
    pairs: lengths, h0 and every base, including ambiguous ones and ragged / very short sequences."""
    import numpy as np
    from genarchbench_b200 import pairio
    c = pairio.preset(4)
    c.len2_min, c.len2_max, c.n_frac = 1, 700, 0.3
    b = pairio.generate(c, 5000, seed=99)
    t, p = str(tmp_path / "pairs.txt"), str(tmp_path / "pairs.bswp")
    pairio.write_text(t, b)
    pairio.write_packed(p, b)
    import os
    assert os.path.getsize(p) * 2 < os.path.getsize(t)      # 30 % of these pairs are 4-bit (ambiguous base)
    for got in (pairio.read_text(t), pairio.read_packed(p)):
        assert len(got) == len(b)
        for f in ("len1", "len2", "h0"):
            assert (got.pairs[f] == b.pairs[f]).all()
        assert (got.pairs["score"] == -1).all()
        for k in list(range(50)) + [len(b) - 1]:
            pa, pb = got.pairs[k], b.pairs[k]
            assert (got.ref[pa["idr"]:pa["idr"] + pa["len1"]] == b.ref[pb["idr"]:pb["idr"] + pb["len1"]]).all()
            assert (got.qer[pa["idq"]:pa["idq"] + pa["len2"]] == b.qer[pb["idq"]:pb["idq"] + pb["len2"]]).all()
        assert int(got.ref[:got.pairs["len1"].sum()].astype(np.int64).sum()) == int(b.ref[:b.pairs["len1"].sum()].astype(np.int64).sum())
    with open(str(tmp_path / "no_newline.txt"), "w") as fh:      # last line without '\n', CRLF line ends
        fh.write("7\r\n0123\r\n012\r\n5\n3210\n32")
    g = pairio.read_text(str(tmp_path / "no_newline.txt"))
    assert g.pairs["len1"].tolist() == [4, 4] and g.pairs["len2"].tolist() == [3, 2] and g.pairs["h0"].tolist() == [7, 5]
    assert g.qer[g.pairs["idq"][1]:g.pairs["idq"][1] + 2].tolist() == [3, 2]


def test_packed_in_memory_form_equals_the_file(tmp_path):
    """pairio.pack (bsw_pack_pairs) produces exactly the records and data of a BSWPAIR1 file; a truncated or
    corrupt file is refused instead of driving huge allocations."""
    c = pairio.preset(4)
    c.n_frac = 0.3
    b = pairio.generate(c, 3000, seed=12)
    rec, data = pairio.pack(b)
    path = str(tmp_path / "p.bswp")
    pairio.write_packed(path, b)
    r2, d2 = pairio.read_packed_raw(path)
    assert (r2 == rec).all() and (d2 == data).all()
    assert os.path.getsize(path) == 24 + rec.nbytes + data.nbytes
    raw = open(path, "rb").read()
    bad = str(tmp_path / "bad.bswp")
    open(bad, "wb").write(raw[: len(raw) // 2])                # truncated
    with pytest.raises(OSError):
        pairio.read_packed_raw(bad)
    with pytest.raises(OSError):
        pairio.read_packed(bad)
    open(bad, "wb").write(raw[:8] + (2 ** 60).to_bytes(8, "little") + raw[16:])   # absurd pair count
    with pytest.raises(OSError):
        pairio.read_packed(bad)
