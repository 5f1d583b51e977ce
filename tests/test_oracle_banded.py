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


def test_degenerate_c4_case_returns_empty_alignment(oracle):
    """Independent random 500 x 10,000 has no 32-mer match -> full matrix -> 5,010,501 > MAX_CELLS ->
    the reference returns the empty MIN_SCORE alignment (banded.rs:104, 407-420; BASELINE.md note on C4)."""
    rng = np.random.default_rng(3)
    x = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 500)])
    y = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 10000)])
    s, _ = oracle.make_scoring(-5, -1, 1, -1)
    aln, ops = oracle.banded_align("semiglobal", s, 32, 32, x, y)
    assert aln["score"] == oracle.MIN_SCORE and ops == [] and aln["xlen"] == 0 and aln["ylen"] == 0
