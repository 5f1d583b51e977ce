"""Pins the banded oracle (oracle/banded_oracle.cpp) to the reference's own tests for
banded::Aligner / Band / sparse::{find_kmer_matches, sdpkpp} (banded.rs:29-90, 1470-2416; sparse.rs:505-770)."""
import json
import os

import numpy as np
import pytest

from golden_util import HERE, clip, load_cases, parse_ops, scoring_fields

with open(os.path.join(HERE, "golden", "banded_vectors.json")) as f:
    G = json.load(f)
FULL = {c["name"]: c for c in load_cases()}


def _scoring(orc, sc):
    f = scoring_fields(sc)
    table = None
    if f["matrix"]:
        from rust_bio_b200 import scores
        table = scores.matrix_table256(f["matrix"])
    return orc.make_scoring(f["gap_open"], f["gap_extend"], f["match"], f["mismatch"], table,
                            f["xclip_prefix"], f["xclip_suffix"], f["yclip_prefix"], f["yclip_suffix"],
                            1 if f["from_scores"] else 0)


def test_band_add_entry_ranges(oracle):
    for step in G["band_add_entry"]["steps"]:
        got = oracle.band_ops(step["m"], step["n"], [tuple(o) for o in step["ops"]])
        assert got == step["ranges"], step


def test_band_add_kmer_equals_k_add_entries(oracle):
    for r, c, k, w, m, n in G["band_add_kmer_equals_entries"]["cases"]:
        a = oracle.band_ops(m, n, [("kmer", r, c, k, w)])
        b = oracle.band_ops(m, n, [("entry", r + i, c + i, 0, w) for i in range(k)])
        assert a == b


@pytest.mark.parametrize("pair", G["differential_k10_w10"]["pairs"], ids=lambda p: p["name"])
def test_banded_equals_full_like_the_reference(oracle, pair):
    s, _ = oracle.make_scoring(-5, -1, 1, -1)
    x, y = pair["x"].encode(), pair["y"].encode()
    for mode in pair["modes"]:
        b_aln, b_ops = oracle.banded_align(mode, s, 10, 10, x, y)
        f_aln, f_ops = oracle.align(mode, s, x, y)
        assert b_aln == f_aln and b_ops == f_ops, (pair["name"], mode)


def _check(case, aln, ops):
    exp = case["expect"]
    for k in ("score", "xstart", "xend", "ystart", "yend"):
        if k in exp:
            assert aln[k] == exp[k], (case["name"], k, aln)
    if "ops" in exp:
        assert ops == parse_ops(exp["ops"]), (case["name"], ops)
    if "x_aln_len" in exp:
        assert aln["xend"] - aln["xstart"] == exp["x_aln_len"] and aln["yend"] - aln["ystart"] == exp["y_aln_len"]
    if exp.get("yend_is_ylen"):
        assert aln["yend"] == aln["ylen"]


@pytest.mark.parametrize("name", G["same_as_full"]["cases"])
def test_full_vectors_through_banded_k10_w10(oracle, name):
    case = FULL[name]
    s, keep = _scoring(oracle, case["scoring"])
    aln, ops = oracle.banded_align(case["mode"], s, 10, 10, case["x"].encode(), case["y"].encode())
    _check(case, aln, ops)


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: c["name"])
def test_banded_known_answers(oracle, case):
    s, keep = _scoring(oracle, case["scoring"])
    aln, ops = oracle.banded_align(case["mode"], s, case["k"], case["w"], case["x"].encode(), case["y"].encode())
    _check(case, aln, ops)


def test_sparse_known_answers(oracle):
    sp = G["sparse"]
    c = sp["find_kmer_count"]
    assert len(oracle.find_kmer_matches(c["s1"].encode(), c["s2"].encode(), c["k"])) == c["count"]
    for c in sp["expanded_find_kmer"]["cases"]:
        assert oracle.find_kmer_matches(c["x"].encode(), c["y"].encode(), 6) == [tuple(m) for m in c["matches"]]
    g = sp["sdpkpp_same"]
    for c in g["cases"]:
        m = oracle.find_kmer_matches(c["x"].encode(), c["y"].encode(), g["k"])
        path, score = oracle.sdpkpp(m, g["k"], g["match_score"], g["gap_open"], g["gap_extend"])
        assert path == c["path"] and score == c["score"]
    t = sp["sdpkpp_tandem_repeat"]
    m = oracle.find_kmer_matches(t["query"].encode(), t["target"].encode(), t["k"])
    path, score = oracle.sdpkpp(m, t["k"], 1, -1, -1)
    assert score == len(t["query"])
    assert [m[p] for p in path] == [(i, i) for i in range(len(path))]


def test_lcskpp_and_expand_known_answers(oracle):
    """sparse::lcskpp / expand_kmer_matches against the reference's unit vectors (sparse.rs:518-770)."""
    sp = G["sparse"]
    for c in sp["lcskpp"]["cases"]:
        m = oracle.find_kmer_matches(c["s1"].encode(), c["s2"].encode(), c["k"])
        path, score = oracle.lcskpp(m, c["k"])
        assert score == c["score"]
        if "match_path" in c:
            assert [list(m[p]) for p in path] == c["match_path"]
        if "path" in c:
            assert path == c["path"]
        if c.get("diagonal"):
            assert [m[p] for p in path] == [(i, i) for i in range(len(path))]
    for a, b in sp["lcskpp"]["equals_sdpkpp_1_0_0"]["pairs"]:  # strict_compare_lcskpp_sdpkpp
        m = oracle.find_kmer_matches(a.encode(), b.encode(), 8)
        assert oracle.lcskpp(m, 8) == oracle.sdpkpp(m, 8, 1, 0, 0)
    t = sp["sdpkpp_tandem_repeat"]
    m = oracle.find_kmer_matches(t["query"].encode(), t["target"].encode(), t["k"])
    assert oracle.lcskpp(m, t["k"])[1] == len(t["query"])
    for c in sp["expanded_find_kmer"]["cases"]:
        m = oracle.find_kmer_matches(c["x"].encode(), c["y"].encode(), 6)
        assert [list(v) for v in oracle.expand_kmer_matches(c["x"].encode(), c["y"].encode(), 6, m, 1)] == c["expanded_1"]


def test_hinted_entry_points_reduce_to_custom(oracle):
    """custom_with_matches(find_kmer_matches(x, y)) == custom(x, y); custom_with_match_path(sdpkpp path) too;
    expanded matches with 0 allowed mismatches only add diagonal neighbours that are matches already."""
    rng = np.random.default_rng(11)
    s, _ = oracle.make_scoring(-5, -1, 1, -1, None, -3, -4, -2, -6, has_match_scores=1)
    for trial in range(12):
        y = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 160)])
        x = bytearray(y[30:110])
        for p in rng.integers(0, len(x), 4):
            x[p] = b"ACGT"[rng.integers(0, 4)]
        x = bytes(x)
        k, w = 6, 5
        want = oracle.banded_align("custom", s, k, w, x, y)
        m = oracle.find_kmer_matches(x, y, k)
        got = oracle.banded_align_hinted(s, k, w, x, y, m)
        assert got is not None and (got[0], got[1]) == want
        path, _ = oracle.sdpkpp(m, k, 1, -5, -1)
        got = oracle.banded_align_hinted(s, k, w, x, y, m, path=path)
        assert (got[0], got[1]) == want
        e0 = oracle.expand_kmer_matches(x, y, k, m, 0)
        assert set(m) <= set(e0)
        # unsorted matches: the reference asserts (sparse.rs:213-218)
        if len(m) >= 2:
            assert oracle.banded_align_hinted(s, k, w, x, y, m[::-1]) is None


def test_degenerate_c4_case_returns_empty_alignment(oracle):
    """Independent random 500 x 10,000 has no 32-mer match -> full matrix -> 5,010,501 > MAX_CELLS ->
    the reference returns the empty MIN_SCORE alignment (banded.rs:104, 407-420; BASELINE.md note on C4)."""
    rng = np.random.default_rng(3)
    x = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 500)])
    y = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 10000)])
    s, _ = oracle.make_scoring(-5, -1, 1, -1)
    aln, ops = oracle.banded_align("semiglobal", s, 32, 32, x, y)
    assert aln["score"] == oracle.MIN_SCORE and ops == [] and aln["xlen"] == 0 and aln["ylen"] == 0
