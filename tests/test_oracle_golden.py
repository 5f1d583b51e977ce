"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for
bio::alignment::pairwise (reference mod.rs:21-160 doctests, mod.rs:1203-1769 unit tests)."""
import numpy as np
import pytest

from golden_util import load_cases, parse_ops, scoring_fields

CASES = load_cases()


def _orc_scoring(orc, sc):
    f = scoring_fields(sc)
    table = None
    if f["matrix"]:
        from rust_bio_b200 import scores
        table = scores.matrix_table256(f["matrix"])
    return orc.make_scoring(f["gap_open"], f["gap_extend"], f["match"], f["mismatch"], table,
                            f["xclip_prefix"], f["xclip_suffix"], f["yclip_prefix"], f["yclip_suffix"],
                            1 if f["from_scores"] else 0)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_vector(oracle, case):
    s, keep = _orc_scoring(oracle, case["scoring"])
    aln, ops = oracle.align(case["mode"], s, case["x"].encode(), case["y"].encode())
    exp = case["expect"]
    for k in ("score", "xstart", "xend", "ystart", "yend"):
        if k in exp:
            assert aln[k] == exp[k], (case["name"], k, aln)
    if "ops" in exp:
        assert ops == parse_ops(exp["ops"]), (case["name"], ops)
    assert aln["xlen"] == len(case["x"]) and aln["ylen"] == len(case["y"])
    assert aln["mode"] == oracle.MODES[case["mode"]]


def test_blosum62_known_answers():
    """scores/blosum62.rs:64-77 and the doctest at blosum62.rs:49-53."""
    from rust_bio_b200.scores import blosum62
    o = ord
    assert blosum62(o("H"), o("H")) == 8
    assert blosum62(o("O"), o("*")) == -4
    assert blosum62(o("A"), o("*")) == -4
    assert blosum62(o("*"), o("*")) == 1
    assert blosum62(o("X"), o("X")) == -1
    assert blosum62(o("X"), o("Z")) == -1
    assert blosum62(o("H"), o("A")) == -2


def test_score_tables_match_committed_fixture():
    import json, os
    from rust_bio_b200 import _score_tables as t
    with open(os.path.join(os.path.dirname(__file__), "golden", "score_tables.json")) as f:
        g = json.load(f)
    assert "".join(g["symbols"]) == t.SYMBOLS
    for name, flat in g["tables"].items():
        got = [v for s in t.SYMBOLS for v in t.MATRICES[name][s]]
        assert got == flat, name


def test_oracle_batch_equals_single(oracle):
    """The threaded batch entry (used as the CPU baseline) returns what the single-pair entry returns."""
    from rust_bio_b200 import synth
    blob, xo, xl, yo, yl = synth.ragged_pairs(7, 64, 40, 50)
    s, _ = oracle.make_scoring(-5, -1, 1, -1)
    for mode in ("local", "global", "semiglobal"):
        out, ops, off, _ = oracle.align_batch(mode, s, blob, xo, xl, yo, yl, threads=3)
        for p in range(len(xl)):
            x = bytes(blob[int(xo[p]):int(xo[p]) + int(xl[p])])
            y = bytes(blob[int(yo[p]):int(yo[p]) + int(yl[p])])
            aln, o1 = oracle.align(mode, s, x, y)
            assert aln["score"] == out["score"][p]
            got = [(int(v) & 7, int(v) >> 3) for v in ops[int(off[p]):int(off[p]) + int(out["n_ops"][p])]]
            assert got == o1
