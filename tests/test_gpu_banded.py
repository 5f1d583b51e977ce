"""GPU parity of the banded path (K4 band construction + K3 banded fill/walk) against the banded oracle
and the reference's banded known-answer tests, through the C ABI."""
import json
import os

import numpy as np
import pytest

from golden_util import HERE, load_cases, parse_ops, scoring_fields
from parity_util import MODES
from test_sim_banded import _mutated_window_batch

pytestmark = pytest.mark.gpu
MIN = -858993459
with open(os.path.join(HERE, "golden", "banded_vectors.json")) as f:
    G = json.load(f)
FULL = {c["name"]: c for c in load_cases()}


@pytest.fixture(scope="module")
def eng():
    from rust_bio_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _mirror_scoring(sc):
    from rust_bio_b200 import scores
    from rust_bio_b200.pairwise import Scoring
    f = scoring_fields(sc)
    if f["matrix"]:
        s = Scoring.new(f["gap_open"], f["gap_extend"], getattr(scores, f["matrix"]))
    elif f["from_scores"]:
        s = Scoring.from_scores(f["gap_open"], f["gap_extend"], f["match"], f["mismatch"])
    else:
        ma, mi = f["match"], f["mismatch"]
        s = Scoring.new(f["gap_open"], f["gap_extend"], lambda a, b: ma if a == b else mi)
    s.xclip_prefix, s.xclip_suffix = f["xclip_prefix"], f["xclip_suffix"]
    s.yclip_prefix, s.yclip_suffix = f["yclip_prefix"], f["yclip_suffix"]
    return s


def _run(eng, case, k, w):
    from rust_bio_b200.banded import Aligner
    aligner = Aligner.with_scoring(_mirror_scoring(case["scoring"]), k, w, engine=eng)
    method = {"custom": aligner.custom, "global": aligner.global_, "semiglobal": aligner.semiglobal,
              "local": aligner.local}[case["mode"]]
    return method(case["x"].encode(), case["y"].encode())


def _check(case, aln):
    exp = case["expect"]
    for k in ("score", "xstart", "xend", "ystart", "yend"):
        if k in exp:
            assert getattr(aln, k) == exp[k], (case["name"], k, aln)
    if "ops" in exp:
        assert [(o.code, o.len) for o in aln.operations] == parse_ops(exp["ops"])
    if "x_aln_len" in exp:
        assert aln.x_aln_len() == exp["x_aln_len"] and aln.y_aln_len() == exp["y_aln_len"]
    if exp.get("yend_is_ylen"):
        assert aln.yend == aln.ylen


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: c["name"])
def test_banded_reference_known_answers(eng, case):
    _check(case, _run(eng, case, case["k"], case["w"]))


@pytest.mark.parametrize("name", G["same_as_full"]["cases"])
def test_full_vectors_through_banded(eng, name):
    _check(FULL[name], _run(eng, FULL[name], 10, 10))


@pytest.mark.parametrize("pair", G["differential_k10_w10"]["pairs"], ids=lambda p: p["name"])
def test_banded_equals_full_on_gpu(eng, pair):
    """banded.rs:1621-1753: assert_eq!(banded_alignment, full_alignment), both on the GPU."""
    from rust_bio_b200 import banded, pairwise
    score = lambda a, b: 1 if a == b else -1
    x, y = pair["x"].encode(), pair["y"].encode()
    ba = banded.Aligner.with_capacity(len(x), len(y), -5, -1, score, 10, 10, engine=eng)
    fa = pairwise.Aligner.with_capacity(len(x), len(y), -5, -1, score, engine=eng)
    for mode in pair["modes"]:
        name = "global_" if mode == "global" else mode
        assert getattr(ba, name)(x, y) == getattr(fa, name)(x, y), (pair["name"], mode)


def _c_scoring(go, ge, ma, mi, clips=(MIN, MIN, MIN, MIN), has_ms=1):
    from rust_bio_b200._lib import CScoring
    return CScoring(go, ge, clips[0], clips[1], clips[2], clips[3], ma, mi, has_ms, None, None, 0)


def _compare(eng, oracle, mode, cs, s, k, w, batch, what):
    ref, rops, roff, _, ref_cells = oracle.banded_align_batch(mode, s, k, w, *batch, threads=8)
    res = eng.align_batch_banded(MODES[mode], cs, k, w, batch)
    assert int(eng.stats.cells) == ref_cells, what
    for f in ("score", "xstart", "xend", "ystart", "yend"):
        assert np.array_equal(getattr(res, f).astype(np.int64), ref[f].astype(np.int64)), (what, f)
    for p in range(len(batch[2])):
        want = [(int(v) & 7, int(v) >> 3) for v in rops[int(roff[p]):int(roff[p]) + int(ref["n_ops"][p])]]
        assert res.ops_of(p) == want, (what, p)


@pytest.mark.parametrize("mode", ["semiglobal", "local", "global"])
def test_mutated_windows_parity(eng, oracle, mode):
    batch = _mutated_window_batch(5, 300, 120, 700)
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    _compare(eng, oracle, mode, _c_scoring(-5, -1, 1, -1), s, 12, 8, batch, f"mutated windows {mode}")


def test_c4_shape_sample_k32_w32(eng, oracle):
    """BASELINE config 4 shape: 500 x 10,000 semiglobal, banded k=32 (w=32), mutated-window generator."""
    batch = _mutated_window_batch(9, 48, 500, 10000, sub=0.05, indel=0.01)
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    _compare(eng, oracle, "semiglobal", _c_scoring(-5, -1, 1, -1), s, 32, 32, batch, "C4 sample")


def test_c4_degenerate_independent_random_is_refused_like_the_reference(eng, oracle):
    """Independent random 500 x 10,000: no 32-mer match -> full matrix -> 5,010,501 > MAX_CELLS -> empty alignment."""
    from rust_bio_b200 import synth
    batch = synth.uniform_pairs(synth.BASES["C4"], 0, 4, 500, 10000)
    res = eng.align_batch_banded(MODES["semiglobal"], _c_scoring(-5, -1, 1, -1), 32, 32, batch)
    assert all(int(v) == MIN for v in res.score) and int(res.ops_off[-1]) == 0
    assert int(eng.stats.cells) == 4 * 501 * 10001
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    ref, *_ = oracle.banded_align_batch("semiglobal", s, 32, 32, *batch, threads=4)
    assert all(int(v) == MIN for v in ref["score"])


@pytest.mark.parametrize("seed", [1, 3, 5])
def test_random_custom_clips_banded(eng, oracle, seed):
    rng = np.random.default_rng(50 + seed)
    pick = lambda: int(rng.choice([MIN, 0, 0, -2, -9, -40]))
    go, ge = int(rng.choice([0, -1, -5, -13])), int(rng.choice([0, -1, -2]))
    ma, mi = int(rng.choice([1, 2, 3])), int(rng.choice([-1, -3, -5]))
    clips = (pick(), pick(), pick(), pick())
    s, _ = oracle.make_scoring(go, ge, ma, mi, None, *clips, has_match_scores=1)
    batch = _mutated_window_batch(seed, 200, 60, 150, sub=0.08, indel=0.04)
    k, w = int(rng.choice([4, 6, 8])), int(rng.choice([3, 5, 9]))
    ref, *_ = oracle.banded_align_batch("custom", s, k, w, *batch, threads=8)
    if np.any(ref["n_ops"] == 0xFFFFFFFF):
        from rust_bio_b200._lib import B2AError
        with pytest.raises(B2AError):  # the reference panics / hangs on some pair: the batch is refused
            eng.align_batch_banded(MODES["custom"], _c_scoring(go, ge, ma, mi, clips), k, w, batch)
    else:
        _compare(eng, oracle, "custom", _c_scoring(go, ge, ma, mi, clips), s, k, w, batch, f"banded custom {seed}")


@pytest.mark.parametrize("seed", range(4))
def test_caller_supplied_band_inputs_batch(eng, oracle, seed):
    """custom_with_matches / custom_with_expanded_matches / custom_with_match_path over batches, through
    b2a_align_batch_banded_hinted, against the oracle's restatement of banded.rs:313-401."""
    from rust_bio_b200.banded import Aligner, find_kmer_matches
    from rust_bio_b200.pairwise import Scoring
    from test_sim_banded import _window_pair
    rng = np.random.default_rng(4000 + seed)
    clips = [int(rng.choice([MIN, 0, -3, -9])) for _ in range(4)]
    go, ge, ma, mi = int(rng.choice([-1, -5])), int(rng.choice([0, -1])), int(rng.choice([1, 2])), -2
    s_o, _ = oracle.make_scoring(go, ge, ma, mi, None, *clips, has_match_scores=1)
    sc = Scoring.from_scores(go, ge, ma, mi)
    sc.xclip_prefix, sc.xclip_suffix, sc.yclip_prefix, sc.yclip_suffix = clips
    k, w = int(rng.choice([5, 7])), int(rng.choice([3, 6]))
    aligner = Aligner.with_scoring(sc, k, w, engine=eng)
    pairs = [_window_pair(rng, 80, 200) for _ in range(60)]
    matches = []
    for i, (x, y) in enumerate(pairs):
        m = find_kmer_matches(x, y, k)
        assert m == oracle.find_kmer_matches(x, y, k)
        matches.append([mt for j, mt in enumerate(m) if j % 4 != 2] if i % 2 else m)

    def check(got, kw):
        n_ok = 0
        for (x, y), m, a in zip(pairs, matches, got):
            want = oracle.banded_align_hinted(s_o, k, w, x, y, m, **kw)
            assert want is not None
            f, ops, _ = want
            assert (a.score, a.xstart, a.xend, a.ystart, a.yend) == (f["score"], f["xstart"], f["xend"], f["ystart"], f["yend"]), kw
            assert [(o.code, o.len) for o in a.operations] == ops, kw
            n_ok += 1
        return n_ok

    keep = [i for i, ((x, y), m) in enumerate(zip(pairs, matches))
            if all(oracle.banded_align_hinted(s_o, k, w, x, y, m, **kw) is not None
                   for kw in (dict(), dict(allowed_mismatches=1), dict(use_lcskpp_union=True),
                              dict(allowed_mismatches=2, use_lcskpp_union=True)))]
    assert len(keep) > 40  # pairs on which the reference itself panics / hangs are left out of the batch
    pairs = [pairs[i] for i in keep]
    matches = [matches[i] for i in keep]
    check(aligner.custom_with_matches_batch(pairs, matches), dict())
    check(aligner.custom_with_expanded_matches_batch(pairs, matches, 1, False), dict(allowed_mismatches=1))
    check(aligner.custom_with_expanded_matches_batch(pairs, matches, None, True), dict(use_lcskpp_union=True))
    check(aligner.custom_with_expanded_matches_batch(pairs, matches, 2, True),
          dict(allowed_mismatches=2, use_lcskpp_union=True))
    # a caller-chosen path: the lcskpp chain of each pair
    sub = [(p, m, oracle.lcskpp(m, k)[0]) for p, m in zip(pairs, matches) if m]
    sub = [(p, m, pa) for p, m, pa in sub if oracle.banded_align_hinted(s_o, k, w, p[0], p[1], m, path=pa) is not None]
    got = aligner.custom_with_match_path_batch([p for p, _, _ in sub], [m for _, m, _ in sub], [pa for _, _, pa in sub])
    for (p, m, pa), a in zip(sub, got):
        f, ops, _ = oracle.banded_align_hinted(s_o, k, w, p[0], p[1], m, path=pa)
        assert a.score == f["score"] and [(o.code, o.len) for o in a.operations] == ops
    # prehash forms are custom / semiglobal
    x, y = pairs[0]
    assert aligner.custom_with_prehash(x, y, None).operations == aligner.custom(x, y).operations
    assert aligner.semiglobal_with_prehash(x, y, None).score == aligner.semiglobal(x, y).score


def test_caller_supplied_band_inputs_refusals(eng, oracle):
    from rust_bio_b200._lib import B2AError
    from rust_bio_b200.banded import Aligner, find_kmer_matches
    from rust_bio_b200.pairwise import Scoring
    aligner = Aligner.with_scoring(Scoring.from_scores(-5, -1, 1, -1), 6, 3, engine=eng)
    x, y = b"ACGTACGTTGCAACGT", b"TTACGTACGTTGCAACGTAA"
    m = find_kmer_matches(x, y, 6)
    with pytest.raises(B2AError, match="reference panics"):
        aligner.custom_with_matches(x, y, m[::-1])
    with pytest.raises(B2AError, match="reference panics"):
        aligner.custom_with_match_path(x, y, m, [0, len(m)])
    with pytest.raises(B2AError, match="reference panics"):
        aligner.custom_with_match_path(x, y, m, [])
    with pytest.raises(B2AError, match="reference panics"):
        aligner.custom_with_matches(x, y, [(2, 500)])
    a = aligner.custom_with_matches(x, y, [])  # no matches: the full matrix (banded.rs:1309-1313)
    s_o, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    f, ops, _ = oracle.banded_align_hinted(s_o, 6, 3, x, y, [])
    assert a.score == f["score"] and [(o.code, o.len) for o in a.operations] == ops


def test_band_ranges_and_visualize(eng, oracle):
    """b2a_banded_band_ranges returns Band::ranges of the last call (banded.rs:1053-1065); visualize draws it with
    the path like banded.rs:1007-1030."""
    import io
    from rust_bio_b200.banded import Aligner
    from rust_bio_b200.pairwise import Scoring
    from test_sim_banded import _window_pair
    rng = np.random.default_rng(77)
    s_o, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    aligner = Aligner.with_scoring(Scoring.from_scores(-5, -1, 1, -1), 6, 4, engine=eng)
    for _ in range(4):
        x, y = _window_pair(rng, 40, 90)
        aln = aligner.semiglobal(x, y)
        want, cells = oracle.band_create("semiglobal", s_o, 6, 4, x, y)
        got = eng.banded_band_ranges(0, len(y))
        assert [(int(a), int(b)) for a, b in got] == want
        buf = io.StringIO()
        text = aligner.visualize(aln, file=buf)
        lines = text.split("\n")
        assert len(lines) == len(x) + 1 and all(len(l) == len(y) + 1 for l in lines) and buf.getvalue() == text + "\n"
        in_band = {(i, j) for j, (a, b) in enumerate(want) for i in range(a, b)}
        on_path = {(p[0], p[1]) for p in aln.path()}
        for i, l in enumerate(lines):
            for j, ch in enumerate(l):
                assert ch == ("\\" if (i, j) in on_path else "x" if (i, j) in in_band else "."), (i, j)
    # batches: ranges of any pair of the (single) wave
    pairs = [_window_pair(rng, 40, 90) for _ in range(5)]
    aligner.custom_batch(pairs)
    x, y = pairs[3]
    want, _ = oracle.band_create("custom", s_o, 6, 4, x, y)
    assert [(int(a), int(b)) for a, b in eng.banded_band_ranges(3, len(y))] == want


# --------------------------------------------------------------------------------------------------------
# Round 2: the named C4 generator at the SURVEY 8d sample size, band ranges included, and the fuzz target's
# path re-scoring property on banded results

def test_c4_named_generator_2000_pairs_with_ranges_and_rescoring(eng, oracle):
    """BASELINE config 4 with its own generator (synth.mutated_window_pairs: y uniform, x = mutated window of
    y): 2,000 pairs of 500 x 10,000, banded semiglobal k=32 w=32 -- alignments, Band::num_cells, the band
    ranges of sampled pairs, and (independent of the oracle) every path re-scored."""
    from parity_util import rescore_path
    from rust_bio_b200 import synth
    batch = synth.mutated_window_pairs(synth.BASES["C4"], 0, 2000, 500, 10000)
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    cs = _c_scoring(-5, -1, 1, -1)
    ref, rops, roff, _, ref_cells = oracle.banded_align_batch("semiglobal", s, 32, 32, *batch, threads=8)
    res = eng.align_batch_banded(MODES["semiglobal"], cs, 32, 32, batch)
    assert int(eng.stats.cells) == ref_cells
    assert eng.banded_strip_pairs() >= 1900  # config 4's pairs run the strip-wavefront fill (a window at y's end does not)
    for f in ("score", "xstart", "xend", "ystart", "yend"):
        assert np.array_equal(getattr(res, f).astype(np.int64), ref[f].astype(np.int64)), f
    blob, xo, xl, yo, yl = batch
    for p in range(2000):
        want = [(int(v) & 7, int(v) >> 3) for v in rops[int(roff[p]):int(roff[p]) + int(ref["n_ops"][p])]]
        got = res.ops_of(p)
        assert got == want, p
        if int(res.score[p]) == MIN:  # refused pair (no 32-mer match -> full matrix > MAX_CELLS)
            assert got == []
            continue
        x = bytes(blob[int(xo[p]):int(xo[p]) + 500])
        y = bytes(blob[int(yo[p]):int(yo[p]) + 10000])
        f = {k: getattr(res, k)[p] for k in ("xstart", "xend", "ystart", "yend")}
        assert rescore_path(x, y, got, f, "semiglobal", -5, -1, lambda a, b: 1 if a == b else -1,
                            (MIN, MIN, 0, 0)) == int(res.score[p]), p
    for p in (0, 7, 1999):
        x = bytes(blob[int(xo[p]):int(xo[p]) + 500])
        y = bytes(blob[int(yo[p]):int(yo[p]) + 10000])
        want, _ = oracle.band_create("semiglobal", s, 32, 32, x, y)
        assert [(int(a), int(b)) for a, b in eng.banded_band_ranges(p, 10000)] == want


@pytest.mark.parametrize("mode", ["global", "semiglobal", "local", "custom"])
def test_banded_paths_rescore_to_their_score(eng, oracle, mode):
    """fuzz/fuzz_targets/banded_aligner.rs:10-56 on the banded engine: recomputed path score == reported score."""
    from parity_util import rescore_path
    rng = np.random.default_rng(91)
    for trial in range(3):
        go, ge = int(rng.choice([-1, -5, -7])), int(rng.choice([0, -1, -2]))
        ge = max(ge, go)  # opening costs at least as much as extending (the fuzz target's model)
        ma, mi = int(rng.choice([1, 2, 3])), int(rng.choice([-1, -3]))
        clips = (MIN,) * 4
        if mode == "custom":
            clips = tuple(int(rng.choice([MIN, 0, -2, -9])) for _ in range(4))
        batch = _mutated_window_batch(700 + trial, 300, 90, 260, sub=0.06, indel=0.03)
        s, _ = oracle.make_scoring(go, ge, ma, mi, None, *clips, has_match_scores=1)
        ref, *_ = oracle.banded_align_batch(mode, s, 8, 6, *batch, threads=8)
        if np.any(ref["n_ops"] == 0xFFFFFFFF):
            continue  # the reference panics on a pair of this batch
        res = eng.align_batch_banded(MODES[mode], _c_scoring(go, ge, ma, mi, clips), 8, 6, batch)
        blob, xo, xl, yo, yl = batch
        eff = {"global": (MIN,) * 4, "semiglobal": (MIN, MIN, 0, 0), "local": (0, 0, 0, 0)}.get(mode, clips)
        for p in range(len(xl)):
            x = bytes(blob[int(xo[p]):int(xo[p]) + int(xl[p])])
            y = bytes(blob[int(yo[p]):int(yo[p]) + int(yl[p])])
            f = {k: getattr(res, k)[p] for k in ("xstart", "xend", "ystart", "yend")}
            sc = rescore_path(x, y, res.ops_of(p), f, mode, go, ge, lambda a, b: ma if a == b else mi, eff)
            assert sc == int(res.score[p]) == int(ref["score"][p]), (mode, trial, p)


def test_per_pair_status_instead_of_failing_the_batch(eng, oracle):
    """The reference fails per CALL (an assert on unsorted caller matches, banded.rs:313-321 -> sparse.rs:212-217):
    with b2a_results.status the batch succeeds, exactly that pair is flagged, every other pair is bit-exact;
    without it the batch is refused as before."""
    from rust_bio_b200._lib import B2AError
    from rust_bio_b200.banded import Aligner, find_kmer_matches
    from rust_bio_b200.pairwise import Scoring
    from test_sim_banded import _window_pair
    rng = np.random.default_rng(99)
    s_o, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    aligner = Aligner.with_scoring(Scoring.from_scores(-5, -1, 1, -1), 6, 4, engine=eng)
    pairs = [_window_pair(rng, 70, 180) for _ in range(40)]
    matches = [find_kmer_matches(x, y, 6) for x, y in pairs]
    bad = [7, 23]
    for b in bad:
        assert len(matches[b]) >= 2
        matches[b] = matches[b][::-1]
    got = aligner.custom_with_matches_batch(pairs, matches, on_panic="none")
    for i, ((x, y), m, a) in enumerate(zip(pairs, matches, got)):
        if i in bad:
            assert a is None
            continue
        f, ops, _ = oracle.banded_align_hinted(s_o, 6, 4, x, y, m)
        assert a is not None and a.score == f["score"] and [(o.code, o.len) for o in a.operations] == ops, i
    with pytest.raises(B2AError, match="pair 7.*reference panics"):
        aligner.custom_with_matches_batch(pairs, matches)


def test_kmer_match_capacity_grows_instead_of_failing(eng, oracle, monkeypatch):
    """Low-complexity sequences give O(m n) k-mer matches (sparse::find_kmer_matches has no limit); a wave whose
    pairs overflow the per-pair slab is redone with a larger one.  B2A_BANDED_CAP starts the capacity tiny."""
    rng = np.random.default_rng(5)
    pairs = []
    for _ in range(24):
        x = bytearray(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 150)].tobytes())
        y = bytearray(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 150)].tobytes())
        x[40:100] = b"A" * 60  # a 60-nt homopolymer in both reads: 53 x 53 8-mer matches
        y[70:130] = b"A" * 60
        pairs.append((bytes(x), bytes(y)))
    from rust_bio_b200.engine import pack_pairs
    batch = pack_pairs(pairs)
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    assert len(oracle.find_kmer_matches(pairs[0][0], pairs[0][1], 8)) >= 53 * 53
    monkeypatch.setenv("B2A_BANDED_CAP", "64")
    _compare(eng, oracle, "semiglobal", _c_scoring(-5, -1, 1, -1), s, 8, 6, batch, "capacity retry (cap 64)")
    monkeypatch.delenv("B2A_BANDED_CAP")
    _compare(eng, oracle, "semiglobal", _c_scoring(-5, -1, 1, -1), s, 8, 6, batch, "default capacity")


@pytest.mark.parametrize("mode", ["semiglobal", "local", "global", "custom"])
def test_register_resident_k3_equals_literal_k3_and_oracle(oracle, mode, monkeypatch):
    """K3's register-resident column loop (the pairs K4 marks) against the literal loop (B2A_BANDED_LITERAL=1) and
    the oracle: random clips, gap_open > gap_extend cases, bands that qualify and bands that are too tall (w = 45)."""
    from rust_bio_b200.engine import Engine
    rng = np.random.default_rng({"semiglobal": 31, "local": 32, "global": 33, "custom": 34}[mode])
    fast_eng = Engine(0)
    monkeypatch.setenv("B2A_BANDED_LITERAL", "1")
    lit_eng = Engine(0)
    monkeypatch.delenv("B2A_BANDED_LITERAL")
    try:
        for trial in range(6):
            pick = lambda: int(rng.choice([MIN, 0, 0, -2, -9]))
            go, ge = int(rng.choice([0, -1, -5])), int(rng.choice([0, -1, -2]))
            ma, mi = int(rng.choice([1, 2])), int(rng.choice([-1, -3]))
            clips = (pick(), pick(), pick(), pick()) if mode == "custom" else (MIN,) * 4
            k, w = int(rng.choice([4, 6, 9])), int(rng.choice([3, 8, 20, 45]))
            batch = _mutated_window_batch(900 + trial, 150, int(rng.integers(60, 220)), int(rng.integers(260, 700)),
                                          sub=0.07, indel=0.03)
            s, _ = oracle.make_scoring(go, ge, ma, mi, None, *clips, has_match_scores=1)
            ref, rops, roff, _, ref_cells = oracle.banded_align_batch(mode, s, k, w, *batch, threads=8)
            if np.any(ref["n_ops"] == 0xFFFFFFFF):
                continue
            cs = _c_scoring(go, ge, ma, mi, clips)
            a = fast_eng.align_batch_banded(MODES[mode], cs, k, w, batch)
            b = lit_eng.align_batch_banded(MODES[mode], cs, k, w, batch)
            assert int(fast_eng.stats.cells) == ref_cells == int(lit_eng.stats.cells)
            for f in ("score", "xstart", "xend", "ystart", "yend", "ops_off", "clip_len"):
                assert np.array_equal(getattr(a, f), getattr(b, f)), (mode, trial, f)
                if f in ref.dtype.names:
                    assert np.array_equal(getattr(a, f).astype(np.int64), ref[f].astype(np.int64)), (mode, trial, f)
            tot = int(a.ops_off[-1])
            assert np.array_equal(a.ops[:tot], b.ops[:tot])
            for p in range(0, 150, 7):
                want = [(int(v) & 7, int(v) >> 3) for v in rops[int(roff[p]):int(roff[p]) + int(ref["n_ops"][p])]]
                assert a.ops_of(p) == want, (mode, trial, p)
    finally:
        fast_eng.close()
        lit_eng.close()


@pytest.mark.parametrize("mode", ["semiglobal", "custom_y", "custom_xy", "local", "custom_xs", "x_prefix_only", "global"])
def test_strip_wavefront_fill_equals_column_loops_and_oracle(oracle, mode, monkeypatch):
    """K3s (b2a_banded_strip.cuh: the packed cell of K1 on fixed 128-row strips with a band mask, four pairs to a
    warp, 4-bit traceback, finish pass) against the K3 column loops (B2A_BANDED_STRIP=0) and the oracle: ragged
    batches (1-5 strips per pair), gap_open > gap_extend cases, y clips with different penalties, both prefix
    clips live (their shared priority code), local mode and a custom x-suffix clip (the column tracker).  The path
    must have taken most of the semiglobal and local pairs."""
    from rust_bio_b200.engine import Engine
    rng = np.random.default_rng({"semiglobal": 41, "custom_y": 42, "custom_xy": 43, "local": 44, "custom_xs": 45, "x_prefix_only": 46,
                                 "global": 47}[mode])
    strip_eng = Engine(0)
    monkeypatch.setenv("B2A_BANDED_STRIP", "0")
    loop_eng = Engine(0)
    monkeypatch.delenv("B2A_BANDED_STRIP")
    taken = total = 0
    try:
        for trial in range(6):
            go, ge = int(rng.choice([0, -1, -5, -5])), int(rng.choice([0, -1, -1, -2]))
            ma, mi = int(rng.choice([1, 2])), int(rng.choice([-1, -3]))
            if mode == "semiglobal":
                omode, clips = "semiglobal", (MIN,) * 4
            elif mode == "custom_y":
                omode, clips = "custom", (MIN, MIN, int(rng.choice([0, -2, -7])), int(rng.choice([0, -1, -6])))
            elif mode == "global":     # the band holds cells of column n: the finish pass runs the literal loop on it
                omode, clips = "global", (MIN,) * 4
            elif mode == "local":      # row and column trackers, both prefix clip terms
                omode, clips = "local", (MIN,) * 4
            elif mode == "x_prefix_only":  # every y clip dead: band cells hold sentinel-derived values, pairs whose
                # final score is not real are handed back to the column loops (bit 10) inside the same call
                omode, clips = "custom", (int(rng.choice([0, -3])), MIN, MIN, MIN)
            elif mode == "custom_xs":  # the column tracker with its own penalty
                omode, clips = "custom", (int(rng.choice([MIN, 0, -3])), int(rng.choice([0, -2, -5])), int(rng.choice([0, -1])),
                                          int(rng.choice([MIN, 0, -4])))
            else:
                omode, clips = "custom", (int(rng.choice([0, -2, -8])), MIN, int(rng.choice([0, -3])), int(rng.choice([0, -4])))
            k, w = int(rng.choice([4, 6, 9])), int(rng.choice([3, 8, 20, 45]))
            batch = _mutated_window_batch(1200 + trial, 203, int(rng.integers(60, 600)), int(rng.integers(700, 1500)),
                                          sub=0.07, indel=0.03)
            s, _ = oracle.make_scoring(go, ge, ma, mi, None, *clips, has_match_scores=1)
            ref, rops, roff, _, ref_cells = oracle.banded_align_batch(omode, s, k, w, *batch, threads=8)
            if np.any(ref["n_ops"] == 0xFFFFFFFF):
                continue
            cs = _c_scoring(go, ge, ma, mi, clips)
            a = strip_eng.align_batch_banded(MODES[omode], cs, k, w, batch)
            taken += strip_eng.banded_strip_pairs()
            total += 203
            b = loop_eng.align_batch_banded(MODES[omode], cs, k, w, batch)
            assert loop_eng.banded_strip_pairs() == 0
            assert int(strip_eng.stats.cells) == ref_cells == int(loop_eng.stats.cells)
            for f in ("score", "xstart", "xend", "ystart", "yend", "ops_off", "clip_len"):
                assert np.array_equal(getattr(a, f), getattr(b, f)), (mode, trial, f)
                if f in ref.dtype.names:
                    assert np.array_equal(getattr(a, f).astype(np.int64), ref[f].astype(np.int64)), (mode, trial, f)
            tot = int(a.ops_off[-1])
            assert np.array_equal(a.ops[:tot], b.ops[:tot])
            for p in range(0, 203, 7):
                want = [(int(v) & 7, int(v) >> 3) for v in rops[int(roff[p]):int(roff[p]) + int(ref["n_ops"][p])]]
                assert a.ops_of(p) == want, (mode, trial, p)
        if mode in ("semiglobal", "local", "global"):
            assert taken * 2 >= total, (taken, total)
    finally:
        strip_eng.close()
        loop_eng.close()


def test_strip_path_in_several_sub_waves(oracle):
    """A small scratch budget cuts the banded call into several K3 sub-waves (b2a_engine.cu banded_impl): each has its
    own list of strip pairs, strip areas and side-stream pass; results must not depend on the cut."""
    from rust_bio_b200.engine import Engine
    batch = _mutated_window_batch(4321, 240, 300, 1200, sub=0.06, indel=0.02)
    s, _ = oracle.make_scoring(-5, -1, 1, -1, None, MIN, MIN, MIN, MIN, has_match_scores=1)
    ref, rops, roff, _, ref_cells = oracle.banded_align_batch("semiglobal", s, 8, 12, *batch, threads=8)
    cs = _c_scoring(-5, -1, 1, -1)
    eng = Engine(0)
    try:
        whole = eng.align_batch_banded(MODES["semiglobal"], cs, 8, 12, batch)
        taken = eng.banded_strip_pairs()
        eng.set_traceback_budget(6 << 20)  # ~3 MB of K3 slabs + strip areas per sub-wave: a few dozen pairs each
        cut = eng.align_batch_banded(MODES["semiglobal"], cs, 8, 12, batch)
        assert eng.banded_strip_pairs() == taken and taken >= 120
        assert eng.stats.kernel_launches > 12  # several sub-waves' worth of launches
        for f in ("score", "xstart", "xend", "ystart", "yend", "ops_off", "clip_len"):
            assert np.array_equal(getattr(whole, f), getattr(cut, f)), f
            if f in ref.dtype.names:
                assert np.array_equal(getattr(cut, f).astype(np.int64), ref[f].astype(np.int64)), f
        tot = int(whole.ops_off[-1])
        assert np.array_equal(whole.ops[:tot], cut.ops[:tot])
        for p in range(0, 240, 5):
            want = [(int(v) & 7, int(v) >> 3) for v in rops[int(roff[p]):int(roff[p]) + int(ref["n_ops"][p])]]
            assert cut.ops_of(p) == want, p
    finally:
        eng.close()


@pytest.mark.parametrize("mode", ["local", "semiglobal"])
def test_strip_wavefront_fill_blosum62(oracle, mode, monkeypatch):
    """A tabulated MatchFunc through the strip fill (sequence bytes mapped to LUT codes as they are loaded, scores from
    K1's scaled LUT in shared memory): protein reads against windows of longer sequences, BLOSUM62, against the
    column loops and the oracle."""
    import ctypes as C
    from rust_bio_b200 import scores, synth
    from rust_bio_b200._lib import CScoring
    from rust_bio_b200.engine import Engine, pack_pairs
    table = np.ascontiguousarray(scores.matrix_table256("blosum62"), dtype=np.int32)
    alpha = np.frombuffer(bytes(range(65, 91)) + b"*", dtype=np.uint8).copy()
    aa = np.frombuffer(synth.PROTEIN, dtype=np.uint8)
    rng = np.random.default_rng(7 if mode == "local" else 8)
    pairs = []
    for _ in range(150):
        yl, xl = int(rng.integers(400, 900)), int(rng.integers(60, 380))
        y = aa[rng.integers(0, 20, yl)].copy()
        st = int(rng.integers(0, yl - xl))
        x = y[st:st + xl].copy()
        x[rng.integers(0, xl, max(1, xl // 12))] = aa[rng.integers(0, 20, max(1, xl // 12))]
        pairs.append((bytes(x), bytes(y)))
    batch = pack_pairs(pairs)
    cs = CScoring(-10, -1, MIN, MIN, MIN, MIN, 0, 0, 0, table.ctypes.data_as(C.c_void_p), alpha.ctypes.data_as(C.c_void_p), len(alpha))
    s, keep = oracle.make_scoring(-10, -1, 0, 0, table)
    ref, rops, roff, _, ref_cells = oracle.banded_align_batch(mode, s, 5, 7, *batch, threads=8)
    strip_eng = Engine(0)
    monkeypatch.setenv("B2A_BANDED_STRIP", "0")
    loop_eng = Engine(0)
    monkeypatch.delenv("B2A_BANDED_STRIP")
    try:
        a = strip_eng.align_batch_banded(MODES[mode], cs, 5, 7, batch)
        assert strip_eng.banded_strip_pairs() >= 75
        b = loop_eng.align_batch_banded(MODES[mode], cs, 5, 7, batch)
        assert int(strip_eng.stats.cells) == ref_cells
        for f in ("score", "xstart", "xend", "ystart", "yend", "ops_off", "clip_len"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), (mode, f)
            if f in ref.dtype.names:
                assert np.array_equal(getattr(a, f).astype(np.int64), ref[f].astype(np.int64)), (mode, f)
        for p in range(150):
            want = [(int(v) & 7, int(v) >> 3) for v in rops[int(roff[p]):int(roff[p]) + int(ref["n_ops"][p])]]
            assert a.ops_of(p) == want, (mode, p)
    finally:
        strip_eng.close()
        loop_eng.close()


def test_refused_band_keeps_the_called_methods_mode(eng):
    """banded.rs:407-420 returns the empty MIN_SCORE alignment (xlen = ylen = 0) above MAX_CELLS; global /
    semiglobal / local then overwrite only `.mode` (banded.rs:889-890) -- the mirror must not report Custom."""
    from rust_bio_b200 import synth
    from rust_bio_b200.alignment import AlignmentMode
    from rust_bio_b200.banded import Aligner
    from rust_bio_b200.pairwise import Scoring
    blob, xo, xl, yo, yl = synth.uniform_pairs(synth.BASES["C4"], 0, 1, 500, 10000)
    x, y = bytes(blob[int(xo[0]):int(xo[0]) + 500]), bytes(blob[int(yo[0]):int(yo[0]) + 10000])
    al = Aligner.with_scoring(Scoring.from_scores(-5, -1, 1, -1), 32, 32, engine=eng)
    for method, mode in ((al.global_, AlignmentMode.Global), (al.semiglobal, AlignmentMode.Semiglobal),
                         (al.local, AlignmentMode.Local), (al.custom, AlignmentMode.Custom)):
        a = method(x, y)
        assert a.score == MIN and a.operations == [] and (a.xlen, a.ylen) == (0, 0) and a.mode == mode
