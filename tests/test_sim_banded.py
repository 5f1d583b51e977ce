"""Host logic of the banded path: the device functions of b2a_banded.cuh (K4 band construction, K3
banded fill + walk), compiled for the CPU by tests/sim, against the banded oracle."""
import json
import os

import numpy as np
import pytest

import sim_util
from golden_util import HERE, load_cases, parse_ops, scoring_fields
from parity_util import MODES
from rust_bio_b200 import synth

with open(os.path.join(HERE, "golden", "banded_vectors.json")) as f:
    G = json.load(f)
FULL = {c["name"]: c for c in load_cases()}
MIN = -858993459


def _one(x: bytes, y: bytes):
    blob = np.frombuffer(x + y + b"\0", dtype=np.uint8)
    return (blob, np.array([0], dtype=np.uint64), np.array([len(x)], dtype=np.uint32),
            np.array([len(x)], dtype=np.uint64), np.array([len(y)], dtype=np.uint32))


def _scoring(orc, sc):
    f = scoring_fields(sc)
    table = None
    if f["matrix"]:
        from rust_bio_b200 import scores
        table = scores.matrix_table256(f["matrix"])
    return orc.make_scoring(f["gap_open"], f["gap_extend"], f["match"], f["mismatch"], table,
                            f["xclip_prefix"], f["xclip_suffix"], f["yclip_prefix"], f["yclip_suffix"],
                            1 if f["from_scores"] else 0)


def _check(case, got, ops):
    exp = case["expect"]
    for k in ("score", "xstart", "xend", "ystart", "yend"):
        if k in exp:
            assert int(got[k][0]) == exp[k], (case["name"], k)
    if "ops" in exp:
        assert ops[0] == parse_ops(exp["ops"]), (case["name"], ops[0])
    assert got["status"][0] == 0


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: c["name"])
def test_sim_banded_known_answers(oracle, case):
    s, keep = _scoring(oracle, case["scoring"])
    got, ops, _ = sim_util.banded_batch(MODES[case["mode"]], s, case["k"], case["w"],
                                        *_one(case["x"].encode(), case["y"].encode()))
    if "x_aln_len" in case["expect"] or case["expect"].get("yend_is_ylen"):
        ref, ref_ops = oracle.banded_align(case["mode"], s, case["k"], case["w"], case["x"].encode(), case["y"].encode())
        assert int(got["score"][0]) == ref["score"] and ops[0] == ref_ops
    else:
        _check(case, got, ops)


@pytest.mark.parametrize("name", G["same_as_full"]["cases"])
def test_sim_banded_full_vectors(oracle, name):
    case = FULL[name]
    s, keep = _scoring(oracle, case["scoring"])
    got, ops, _ = sim_util.banded_batch(MODES[case["mode"]], s, 10, 10, *_one(case["x"].encode(), case["y"].encode()))
    _check(case, got, ops)


def _mutated_window_batch(seed, n_pairs, xlen, ylen, sub=0.05, indel=0.01):
    """x = window of y with substitutions/indels, trimmed/padded to xlen (the C4 generator)."""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs, xo, xl, yo, yl = [], [], [], [], []
    pos = 0
    for _ in range(n_pairs):
        y = alpha[rng.integers(0, 4, ylen)]
        st = int(rng.integers(0, max(1, ylen - xlen)))
        src = y[st:st + xlen + 20]
        out = []
        for c in src:
            r = rng.random()
            if r < indel / 2:
                continue
            if r < indel:
                out.append(alpha[rng.integers(0, 4)])
            out.append(alpha[rng.integers(0, 4)] if rng.random() < sub else c)
        x = np.array(out[:xlen], dtype=np.uint8)
        if len(x) < xlen:
            x = np.concatenate([x, alpha[rng.integers(0, 4, xlen - len(x))]])
        xo.append(pos); xl.append(len(x)); pos += len(x)
        yo.append(pos); yl.append(len(y)); pos += len(y)
        seqs += [x, y]
    blob = np.concatenate(seqs + [np.zeros(1, np.uint8)])
    return (blob, np.array(xo, np.uint64), np.array(xl, np.uint32), np.array(yo, np.uint64), np.array(yl, np.uint32))


def _compare_batch(oracle, mode, s, k, w, batch, what):
    ref, rops, roff, _, ref_cells = oracle.banded_align_batch(mode, s, k, w, *batch, threads=4)
    got, ops, rng = sim_util.banded_batch(MODES[mode], s, k, w, *batch, want_ranges=True)
    blob, xo, xl, yo, yl = batch
    pos = 0
    n_panic = [0]
    for p in range(len(xl)):
        x = bytes(blob[int(xo[p]):int(xo[p]) + int(xl[p])])
        y = bytes(blob[int(yo[p]):int(yo[p]) + int(yl[p])])
        want_rng, cells = oracle.band_create(mode, s, k, w, x, y)
        n = int(yl[p])
        got_rng = [(int(rng[pos + 2 * j]), int(rng[pos + 2 * j + 1])) for j in range(n + 1)]
        pos += 2 * (n + 1)
        assert got_rng == want_rng, (what, p, "band ranges differ")
        assert int(got["num_cells"][p]) == cells
        if int(ref["n_ops"][p]) == 0xFFFFFFFF:
            # the reference itself panics / never terminates on this input: the device must flag it
            assert got["status"][p] != 0, (what, p, "reference panics but the device path returned a result")
            n_panic[0] += 1
            continue
        assert got["status"][p] == 0
        for f in ("score", "xstart", "xend", "ystart", "yend"):
            assert int(got[f][p]) == int(ref[f][p]), (what, p, f, x, y)
        want = [(int(v) & 7, int(v) >> 3) for v in rops[int(roff[p]):int(roff[p]) + int(ref["n_ops"][p])]]
        assert ops[p] == want, (what, p, x, y)
    assert int(got["num_cells"].sum()) == ref_cells
    assert n_panic[0] < len(xl) // 2, "too many reference panics for a meaningful comparison"


@pytest.mark.parametrize("mode", ["semiglobal", "local", "global"])
def test_sim_banded_mutated_windows(oracle, mode):
    batch = _mutated_window_batch(5, 40, 120, 700)
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    _compare_batch(oracle, mode, s, 12, 8, batch, f"mutated windows {mode}")


@pytest.mark.parametrize("seed", range(6))
def test_sim_banded_random_custom(oracle, seed):
    rng = np.random.default_rng(50 + seed)
    pick = lambda: int(rng.choice([MIN, 0, 0, -2, -9, -40]))
    go, ge = int(rng.choice([0, -1, -5, -13])), int(rng.choice([0, -1, -2]))
    s, _ = oracle.make_scoring(go, ge, int(rng.choice([1, 2, 3])), int(rng.choice([-1, -3, -5])), None,
                               pick(), pick(), pick(), pick(), has_match_scores=int(seed % 2))
    batch = _mutated_window_batch(seed, 25, 60, 150, sub=0.08, indel=0.04) if seed % 2 else \
        synth.ragged_pairs(seed, 40, 50, 70, alphabet=b"AC", min_len=0)
    _compare_batch(oracle, "custom", s, int(rng.choice([4, 6, 8])), int(rng.choice([3, 5, 9])), batch,
                   f"banded custom seed={seed}")


def test_sim_banded_refuses_more_than_max_cells(oracle):
    rng = np.random.default_rng(3)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    x, y = bytes(alpha[rng.integers(0, 4, 500)]), bytes(alpha[rng.integers(0, 4, 10000)])
    s, _ = oracle.make_scoring(-5, -1, 1, -1)
    got, ops, _ = sim_util.banded_batch(MODES["semiglobal"], s, 32, 32, *_one(x, y))
    assert int(got["score"][0]) == MIN and ops[0] == [] and int(got["num_cells"][0]) == 501 * 10001


def _window_pair(rng, xlen=90, ylen=220, nsub=5):  # noqa: E302
    y = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, ylen)])
    st = int(rng.integers(0, ylen - xlen))
    x = bytearray(y[st:st + xlen])
    for p in rng.integers(0, xlen, nsub):
        x[p] = b"ACGT"[rng.integers(0, 4)]
    if rng.integers(0, 2):  # one indel
        q = int(rng.integers(5, xlen - 5))
        x = x[:q] + x[q + 1:] if rng.integers(0, 2) else x[:q] + b"G" + x[q:]
    return bytes(x), y


@pytest.mark.parametrize("seed", range(8))
def test_sim_banded_caller_supplied_band_inputs(oracle, seed):
    """custom_with_matches / custom_with_expanded_matches / custom_with_match_path (banded.rs:313-401): the host
    build of K4 (matches, expand_kmer_matches, lcskpp union, given path) + K3 against the oracle."""
    rng = np.random.default_rng(900 + seed)
    pick = lambda: int(rng.choice([MIN, 0, 0, -2, -9]))
    s, _ = oracle.make_scoring(int(rng.choice([-1, -5])), int(rng.choice([0, -1])), int(rng.choice([1, 2])),
                               int(rng.choice([-1, -3])), None, pick(), pick(), pick(), pick(),
                               has_match_scores=int(seed % 2))
    n_checked = 0
    for trial in range(10):
        x, y = _window_pair(rng)
        k, w = int(rng.choice([5, 6, 8])), int(rng.choice([3, 6]))
        m = oracle.find_kmer_matches(x, y, k)
        if len(m) > 3 and trial % 3 == 0:  # the caller may pass any sorted subset
            m = [mt for i, mt in enumerate(m) if i % 3 != 1]
        variants = [dict(), dict(allowed_mismatches=0), dict(allowed_mismatches=int(rng.integers(1, 3))),
                    dict(use_lcskpp_union=True), dict(allowed_mismatches=1, use_lcskpp_union=True)]
        if m:
            path, _ = oracle.lcskpp(m, k)
            variants.append(dict(path=path))
            variants.append(dict(path=[int(v) for v in sorted(rng.choice(len(m), size=min(3, len(m)), replace=False))]))
        for kw in variants:
            want = oracle.banded_align_hinted(s, k, w, x, y, m, **kw)
            got = sim_util.banded_hinted_one(s, k, w, x, y, m, **kw)
            if want is None:
                assert got is None, (seed, trial, kw, "the reference panics but the device code returned a result")
                continue
            assert got is not None, (seed, trial, kw)
            assert got[2] == want[2], (seed, trial, kw, "band cells")
            assert got[0] == {f: want[0][f] for f in got[0]} and got[1] == want[1], (seed, trial, kw, x, y)
            n_checked += 1
    assert n_checked > 30


def test_sim_banded_caller_inputs_reference_panics(oracle):
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    x, y = b"ACGTACGTTGCAACGT", b"TTACGTACGTTGCAACGTAA"
    m = oracle.find_kmer_matches(x, y, 6)
    assert len(m) >= 3
    assert sim_util.banded_hinted_one(s, 6, 3, x, y, m[::-1]) is None                       # not ascending
    assert sim_util.banded_hinted_one(s, 6, 3, x, y, m, path=[0, len(m)]) is None           # index out of range
    assert sim_util.banded_hinted_one(s, 6, 3, x, y, m, path=[]) is None                    # path[0] on an empty path
    assert sim_util.banded_hinted_one(s, 6, 3, x, y, [(2, 500)]) is None                    # outside the matrix
    got = sim_util.banded_hinted_one(s, 6, 3, x, y, [])                                     # no matches: full matrix
    want = oracle.banded_align_hinted(s, 6, 3, x, y, [])
    assert got is not None and got[0]["score"] == want[0]["score"] and got[1] == want[1]


# ---------------------------------------------------------------------------------------------------------------
# The same kernels as the GPU instantiates them (W = 32), with 32 host threads standing in for the lanes of a warp
# (tests/sim: B2A_HOST_WARP): the prefix-maximum I chain, chunk carries, the per-column arg-max, the runs of plain
# columns, the cooperative sorts, the segmented k-mer probe and the lane-parallel band geometry -- none of which
# the W = 1 build above exercises.

@pytest.mark.parametrize("case", G["cases"], ids=lambda c: c["name"])
def test_warp32_banded_known_answers(oracle, case):
    s, keep = _scoring(oracle, case["scoring"])
    x, y = case["x"].encode(), case["y"].encode()
    got = sim_util.banded_warp32_one(MODES[case["mode"]], s, case["k"], case["w"], x, y)
    ref, ref_ops = oracle.banded_align(case["mode"], s, case["k"], case["w"], x, y)
    assert got is not None and got[1] == ref_ops
    assert got[0] == {f: ref[f] for f in got[0]}


@pytest.mark.parametrize("mode", ["semiglobal", "local", "global", "custom"])
def test_warp32_banded_mutated_windows(oracle, mode):
    rng = np.random.default_rng({"semiglobal": 1, "local": 2, "global": 3, "custom": 4}[mode])
    pick = lambda: int(rng.choice([MIN, 0, 0, -2, -9]))
    n_ok = 0
    for trial in range(24):
        go, ge = int(rng.choice([0, -1, -5])), int(rng.choice([0, -1, -2]))  # incl. gap_open > gap_extend
        clips = (pick(), pick(), pick(), pick()) if mode == "custom" else (MIN, MIN, MIN, MIN)
        s, _ = oracle.make_scoring(go, ge, int(rng.choice([1, 2])), int(rng.choice([-1, -3])), None, *clips,
                                   has_match_scores=int(trial % 2))
        x, y = _window_pair(rng, int(rng.integers(20, 140)), int(rng.integers(150, 420)), nsub=int(rng.integers(0, 9)))
        k, w = int(rng.choice([4, 6, 9])), int(rng.choice([2, 5, 11]))
        got = sim_util.banded_warp32_one(MODES[mode], s, k, w, x, y, want_ranges=True)
        try:
            ref, ref_ops = oracle.banded_align(mode, s, k, w, x, y)
        except RuntimeError:  # the reference itself panics / hangs: the device code must flag it
            assert got is None
            continue
        want_rng, cells = oracle.band_create(mode, s, k, w, x, y)
        assert got is not None, (mode, trial)
        assert got[3] == want_rng and got[2] == cells, (mode, trial, "band")
        assert got[0] == {f: ref[f] for f in got[0]} and got[1] == ref_ops, (mode, trial, x, y)
        n_ok += 1
    assert n_ok >= 14


def test_warp32_c4_shaped_pairs(oracle):
    """BASELINE config 4's shape (500 x 10,000, k = 32, w = 32): long runs of plain columns, ~110-row bands."""
    blob, xo, xl, yo, yl = _mutated_window_batch(4, 3, 500, 10000, sub=0.05, indel=0.01)
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    for p in range(3):
        x = bytes(blob[int(xo[p]):int(xo[p]) + int(xl[p])])
        y = bytes(blob[int(yo[p]):int(yo[p]) + int(yl[p])])
        got = sim_util.banded_warp32_one(MODES["semiglobal"], s, 32, 32, x, y, want_ranges=True)
        ref, ref_ops = oracle.banded_align("semiglobal", s, 32, 32, x, y)
        want_rng, cells = oracle.band_create("semiglobal", s, 32, 32, x, y)
        assert got is not None and got[3] == want_rng and got[2] == cells
        assert got[0] == {f: ref[f] for f in got[0]} and got[1] == ref_ops


def test_warp32_caller_supplied_band_inputs(oracle):
    rng = np.random.default_rng(321)
    s, _ = oracle.make_scoring(-5, -1, 1, -1, None, -3, MIN, 0, -4, has_match_scores=1)
    n_ok = 0
    for trial in range(6):
        x, y = _window_pair(rng)
        k, w = int(rng.choice([5, 7])), int(rng.choice([3, 6]))
        m = oracle.find_kmer_matches(x, y, k)
        variants = [dict(), dict(allowed_mismatches=1), dict(use_lcskpp_union=True),
                    dict(allowed_mismatches=2, use_lcskpp_union=True)]
        if m:
            variants.append(dict(path=oracle.lcskpp(m, k)[0]))
        for kw in variants:
            want = oracle.banded_align_hinted(s, k, w, x, y, m, **kw)
            got = sim_util.banded_warp32_one(0, s, k, w, x, y, matches=m, **kw)
            if want is None:
                assert got is None
                continue
            assert got is not None and got[2] == want[2]
            assert got[0] == {f: want[0][f] for f in got[0]} and got[1] == want[1], (trial, kw)
            n_ok += 1
    assert n_ok >= 15
    x, y = b"ACGTACGTTGCAACGT", b"TTACGTACGTTGCAACGTAA"
    m = oracle.find_kmer_matches(x, y, 6)
    assert sim_util.banded_warp32_one(0, s, 6, 3, x, y, matches=m[::-1]) is None
    assert sim_util.banded_warp32_one(0, s, 6, 3, x, y, matches=m, path=[0, len(m)]) is None


# ---- the register-resident K3 column loop (banded_columns_fast: a lane owns 5 consecutive rows, S/D of the
# previous column and the row tracker in registers, one circular prefix maximum per column) on 32 emulated lanes.
# banded_warp32_one takes it whenever banded_fast_ok says the pair's band qualifies, as the device does.

@pytest.mark.parametrize("mode", ["semiglobal", "local", "global", "custom"])
def test_warp32_fast_column_loop_vs_literal_vs_oracle(oracle, mode, monkeypatch):
    rng = np.random.default_rng({"semiglobal": 11, "local": 12, "global": 13, "custom": 14}[mode])
    pick = lambda: int(rng.choice([MIN, 0, 0, -2, -9]))
    L = sim_util.lib()
    n_fast = n_ok = 0
    for trial in range(40):
        go, ge = int(rng.choice([0, -1, -5])), int(rng.choice([0, -1, -2]))
        clips = (pick(), pick(), pick(), pick()) if mode == "custom" else (MIN, MIN, MIN, MIN)
        s, _ = oracle.make_scoring(go, ge, int(rng.choice([1, 2])), int(rng.choice([-1, -3])), None, *clips,
                                   has_match_scores=int(trial % 2))
        xl = int(rng.integers(20, 200))
        x, y = _window_pair(rng, xl, xl + int(rng.integers(30, 400)), nsub=int(rng.integers(0, 9)))
        k, w = int(rng.choice([4, 6, 9])), int(rng.choice([2, 5, 11, 20, 45]))  # w = 45: bands taller than 32 x 5 rows
        got = sim_util.banded_warp32_one(MODES[mode], s, k, w, x, y, want_ranges=True)
        fast = L.sim_last_banded_fast()
        try:
            ref, ref_ops = oracle.banded_align(mode, s, k, w, x, y)
        except RuntimeError:
            assert got is None
            continue
        assert got is not None and got[0] == {f: ref[f] for f in got[0]} and got[1] == ref_ops, (mode, trial, fast)
        n_ok += 1
        n_fast += fast
        if fast and trial % 4 == 0:  # the literal loop on the same pair gives the same answer
            monkeypatch.setenv("B2A_SIM_BANDED_LITERAL", "1")
            lit = sim_util.banded_warp32_one(MODES[mode], s, k, w, x, y, want_ranges=True)
            monkeypatch.delenv("B2A_SIM_BANDED_LITERAL")
            assert L.sim_last_banded_fast() == 0 and lit[0] == got[0] and lit[1] == got[1]
    assert n_ok >= 30 and n_fast >= 20


def test_warp32_fast_column_loop_blosum_and_c4_shape(oracle):
    """LUT scoring (a tabulated MatchFunc) through the fast loop, and BASELINE config 4's shape on it."""
    from rust_bio_b200 import scores
    table = scores.matrix_table256("blosum62")
    rng = np.random.default_rng(77)
    L = sim_util.lib()
    alpha = np.frombuffer(synth.PROTEIN, dtype=np.uint8)
    s, keep = oracle.make_scoring(-10, -1, 0, 0, table)
    n_fast = 0
    for trial in range(10):
        y = alpha[rng.integers(0, 20, 300)].copy()
        st = int(rng.integers(0, 150))
        x = y[st:st + 120].copy()
        x[rng.integers(0, 120, 10)] = alpha[rng.integers(0, 20, 10)]
        x, y = bytes(x), bytes(y)
        for mode in ("local", "semiglobal"):
            got = sim_util.banded_warp32_one(MODES[mode], s, 5, 7, x, y)
            n_fast += L.sim_last_banded_fast()
            ref, ref_ops = oracle.banded_align(mode, s, 5, 7, x, y)
            assert got is not None and got[0] == {f: ref[f] for f in got[0]} and got[1] == ref_ops, (trial, mode)
    assert n_fast >= 10
    batch = synth.mutated_window_pairs(synth.BASES["C4"], 0, 3, 500, 10000)
    blob, xo, xl, yo, yl = batch
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    for p in range(3):
        x = bytes(blob[int(xo[p]):int(xo[p]) + 500])
        y = bytes(blob[int(yo[p]):int(yo[p]) + 10000])
        got = sim_util.banded_warp32_one(MODES["semiglobal"], s, 32, 32, x, y)
        assert L.sim_last_banded_fast() == 1
        ref, ref_ops = oracle.banded_align("semiglobal", s, 32, 32, x, y)
        assert got is not None and got[0] == {f: ref[f] for f in got[0]} and got[1] == ref_ops


# ---- the strip-wavefront banded fill (b2a_banded_strip.cuh: K1's packed cell on fixed 128-row strips, band mask, 4-bit
# traceback) + its finish pass, as ONE emulated warp-task of up to four pairs with different shapes and windows

@pytest.mark.parametrize("mode", ["semiglobal", "custom_y", "global_nocoln", "custom_xy", "local", "custom_xsuffix", "global"])
def test_warp32_strip_fill_vs_oracle(oracle, mode):
    rng = np.random.default_rng({"semiglobal": 21, "custom_y": 22, "global_nocoln": 23, "custom_xy": 24, "local": 25,
                                 "custom_xsuffix": 26, "global": 27}[mode])
    n_strip = n_tot = 0
    for trial in range(24):
        go, ge = int(rng.choice([0, -1, -5, -5])), int(rng.choice([0, -1, -1, -2]))
        if mode == "semiglobal":
            omode, clips = "semiglobal", (MIN, MIN, MIN, MIN)
        elif mode == "custom_y":     # y clips live with different penalties, x global
            omode, clips = "custom", (MIN, MIN, int(rng.choice([0, -2, -7])), int(rng.choice([0, -1, -6])))
        elif mode == "global":         # the band holds cells of column n: the finish pass runs the literal loop there
            omode, clips = "global", (MIN, MIN, MIN, MIN)
        elif mode == "local":          # every clip live at 0: row and column trackers, both prefix clip terms
            omode, clips = "local", (MIN, MIN, MIN, MIN)
        elif mode == "custom_xsuffix":  # the column tracker with its own penalty, y prefix live (S stays real)
            omode, clips = "custom", (int(rng.choice([MIN, 0, -3])), int(rng.choice([0, -2, -5])), int(rng.choice([0, -1])),
                                      int(rng.choice([MIN, 0, -4])))
        elif mode == "global_nocoln":  # y prefix dead, y suffix dead, x prefix live: sentinel-derived band cells
            omode, clips = "custom", (int(rng.choice([0, -3])), MIN, MIN, MIN)
        else:                          # both prefix clips live (the shared priority code), y suffix live
            omode, clips = "custom", (int(rng.choice([0, -2, -8])), MIN, int(rng.choice([0, -3])), int(rng.choice([0, -4])))
        s, _ = oracle.make_scoring(go, ge, int(rng.choice([1, 2])), int(rng.choice([-1, -3])), None, *clips,
                                   has_match_scores=int(trial % 2))
        k, w = int(rng.choice([4, 6, 9])), int(rng.choice([2, 5, 11, 20, 45]))
        pairs = []
        for q in range(int(rng.integers(1, 5))):
            xl = int(rng.integers(12, 420))      # one to four strips of 128 rows, ragged inside the task
            pairs.append(_window_pair(rng, xl, xl + int(rng.integers(30, 500)), nsub=int(rng.integers(0, 9))))
        got = sim_util.banded_strip_task(MODES[omode], s, k, w, pairs)
        for (x, y), g in zip(pairs, got):
            n_tot += 1
            if g is None:
                continue
            n_strip += 1
            ref, ref_ops = oracle.banded_align(omode, s, k, w, x, y)
            assert g[0] == {f: ref[f] for f in g[0]} and g[1] == ref_ops, (mode, trial, len(x), len(y), k, w)
    # the path must actually be exercised (non-zero clip penalties pull the band into the corner (m, n): column n is
    # then in the band and the pair stays with the K3 loops)
    assert n_strip >= {"semiglobal": n_tot // 3, "custom_y": 8, "custom_xy": 4, "global_nocoln": 0, "local": n_tot // 3,
                       "custom_xsuffix": 4, "global": n_tot // 3}[mode], (n_strip, n_tot)


def test_warp32_strip_fill_c4_shape(oracle):
    batch = synth.mutated_window_pairs(synth.BASES["C4"], 0, 4, 500, 10000)
    blob, xo, xl, yo, yl = batch
    s, _ = oracle.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    pairs = [(bytes(blob[int(xo[p]):int(xo[p]) + 500]), bytes(blob[int(yo[p]):int(yo[p]) + 10000])) for p in range(4)]
    got = sim_util.banded_strip_task(MODES["semiglobal"], s, 32, 32, pairs)
    for (x, y), g in zip(pairs, got):
        assert g is not None, "BASELINE config 4's pairs are strip-eligible"
        ref, ref_ops = oracle.banded_align("semiglobal", s, 32, 32, x, y)
        assert g[0] == {f: ref[f] for f in g[0]} and g[1] == ref_ops


def test_warp32_strip_fill_blosum62(oracle):
    """A tabulated MatchFunc (BLOSUM62) through the strip fill: sequence bytes mapped to LUT codes as they are loaded,
    scores from K1's scaled LUT; local and semiglobal, tasks of one to four protein pairs."""
    from rust_bio_b200 import scores
    table = scores.matrix_table256("blosum62")
    rng = np.random.default_rng(91)
    alpha = np.frombuffer(synth.PROTEIN, dtype=np.uint8)
    n_strip = n_tot = 0
    for trial in range(10):
        s, keep = oracle.make_scoring(int(rng.choice([-10, -5])), -1, 0, 0, table)
        pairs = []
        for q in range(int(rng.integers(1, 5))):
            yl, xl = int(rng.integers(260, 520)), int(rng.integers(40, 250))
            y = alpha[rng.integers(0, 20, yl)].copy()
            st = int(rng.integers(0, yl - xl))
            x = y[st:st + xl].copy()
            x[rng.integers(0, xl, max(1, xl // 12))] = alpha[rng.integers(0, 20, max(1, xl // 12))]
            pairs.append((bytes(x), bytes(y)))
        for mode in ("local", "semiglobal"):
            got = sim_util.banded_strip_task(MODES[mode], s, 5, 7, pairs)
            for (x, y), g in zip(pairs, got):
                n_tot += 1
                if g is None:
                    continue
                n_strip += 1
                ref, ref_ops = oracle.banded_align(mode, s, 5, 7, x, y)
                assert g[0] == {f: ref[f] for f in g[0]} and g[1] == ref_ops, (trial, mode, len(x), len(y))
    assert n_strip * 2 >= n_tot, (n_strip, n_tot)
