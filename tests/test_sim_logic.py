"""Host logic: the kernels' per-lane code (fill_lane<1,R,*> + walk_pair), compiled for the CPU by
tests/sim, against the oracle -- golden vectors, ragged/edge shapes, random custom clip penalties."""
import numpy as np
import pytest

import sim_util
from golden_util import load_cases, parse_ops, scoring_fields
from parity_util import MODES, assert_same, oracle_batch
from rust_bio_b200 import synth

CASES = load_cases()
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
                            f["xclip_prefix"], f["xclip_suffix"], f["yclip_prefix"], f["yclip_suffix"])


@pytest.mark.parametrize("R", [4, 16])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_sim_golden(oracle, case, R):
    s, keep = _scoring(oracle, case["scoring"])
    got, ops = sim_util.align_batch(MODES[case["mode"]], s, *_one(case["x"].encode(), case["y"].encode()), R=R)
    exp = case["expect"]
    for k in ("score", "xstart", "xend", "ystart", "yend"):
        if k in exp:
            assert int(got[k][0]) == exp[k], (case["name"], k)
    if "ops" in exp:
        assert ops[0] == parse_ops(exp["ops"])
    assert got["status"][0] == 0


@pytest.mark.parametrize("mode", ["local", "global", "semiglobal"])
@pytest.mark.parametrize("R,general,no_pack,no_lut", [(4, 0, 0, 0), (8, 1, 0, 0), (16, 0, 0, 0), (8, 0, 1, 0),
                                                      (4, 1, 1, 1), (16, 0, 0, 1)])
def test_sim_ragged_presets(oracle, mode, R, general, no_pack, no_lut):
    """Every kernel variant: specialised / general flags, packed / unpacked trackers, LUT / compare scores."""
    batch = synth.ragged_pairs(11 + R, 300, 70, 90)
    s, _ = oracle.make_scoring(-5, -1, 1, -1)
    ref, ref_ops = oracle_batch(oracle, mode, s, batch)
    got, ops = sim_util.align_batch(MODES[mode], s, *batch, R=R, force_general=general, no_pack=no_pack,
                                    no_lut=no_lut)
    assert_same(got, ops, ref, ref_ops, batch, f"{mode} R={R}")


def test_sim_uniform_150_local(oracle):
    batch = synth.uniform_pairs(synth.BASES["C1"], 0, 96, 150, 150)
    s, _ = oracle.make_scoring(-5, -1, 1, -1)
    ref, ref_ops = oracle_batch(oracle, "local", s, batch)
    got, ops = sim_util.align_batch(MODES["local"], s, *batch, R=16)
    assert_same(got, ops, ref, ref_ops, batch, "C1-shape local")


def test_sim_tiny_shapes_all_modes(oracle):
    """Empty and 1-2 symbol sequences (m == 0 / n == 0 alias S[k][m] with S[k][0], mod.rs:551 note)."""
    xs, ys = [], []
    for m in range(0, 4):
        for n in range(0, 4):
            for rep in range(4):
                xs.append(m)
                ys.append(n)
    rng = np.random.default_rng(5)
    total = sum(xs) + sum(ys)
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 2, size=total + 1)]
    lens = np.array([v for pair in zip(xs, ys) for v in pair], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    batch = (blob, offs[0::2].copy(), np.array(xs, dtype=np.uint32), offs[1::2].copy(), np.array(ys, dtype=np.uint32))
    for mode in ("custom", "local", "global", "semiglobal"):
        for clips in [(MIN, MIN, MIN, MIN), (0, 0, 0, 0), (-1, 0, MIN, -2), (0, MIN, MIN, 0), (MIN, 0, 0, MIN)]:
            s, _ = oracle.make_scoring(-2, -1, 2, -1, None, *clips)
            ref, ref_ops = oracle_batch(oracle, mode, s, batch)
            got, ops = sim_util.align_batch(MODES[mode], s, *batch, R=4)
            assert_same(got, ops, ref, ref_ops, batch, f"tiny {mode} {clips}")


@pytest.mark.parametrize("seed", range(12))
def test_sim_random_custom_clips(oracle, seed):
    """Arbitrary live/dead mixes of the four clip penalties, zero gap costs included."""
    rng = np.random.default_rng(100 + seed)
    pick = lambda: int(rng.choice([MIN, 0, 0, -1, -3, -7, -20]))
    go, ge = int(rng.choice([0, -1, -2, -5, -6])), int(rng.choice([0, -1, -1, -2]))
    ma, mi = int(rng.choice([1, 2, 4])), int(rng.choice([-1, -3, -7, 0]))
    s, _ = oracle.make_scoring(go, ge, ma, mi, None, pick(), pick(), pick(), pick())
    batch = synth.ragged_pairs(seed, 200, 40, 45, alphabet=b"AC" if seed % 2 else b"ACGT")
    ref, ref_ops = oracle_batch(oracle, "custom", s, batch)
    for R, no_pack in ((4, 0), (8, 1)):
        got, ops = sim_util.align_batch(MODES["custom"], s, *batch, R=R, no_pack=no_pack, no_lut=seed % 3 == 0)
        assert_same(got, ops, ref, ref_ops, batch, f"custom seed={seed} R={R}")


def test_sim_blosum62_protein(oracle):
    from rust_bio_b200 import scores
    table = scores.matrix_table256("blosum62")
    batch = synth.ragged_pairs(3, 120, 60, 60, alphabet=synth.PROTEIN, min_len=1)
    for mode, go in (("local", -10), ("global", -5), ("semiglobal", -11)):
        s, keep = oracle.make_scoring(go, -1, 0, 0, table)
        ref, ref_ops = oracle_batch(oracle, mode, s, batch)
        got, ops = sim_util.align_batch(MODES[mode], s, *batch, R=8)
        assert_same(got, ops, ref, ref_ops, batch, f"blosum62 {mode}")


# The fill shapes whose lanes hand rows to each other (G lanes per pair, anti-diagonal wavefront, masked last
# strips, the pair-major boundary row of the warp-per-pair shape): 32 host contexts in lock-step stand in for the
# warp (tests/sim B2A_HOST_WARP), the same fill_lane<G,R,FLAGS> the GPU instantiates.
WAVE_SHAPES = [(4, 16), (8, 8), (32, 8)]


@pytest.mark.parametrize("G,R", WAVE_SHAPES)
@pytest.mark.parametrize("mode", ["local", "global", "semiglobal"])
def test_sim_wavefront_shapes_ragged(oracle, mode, G, R):
    batch = synth.ragged_pairs(300 + G, 40, 1, 3 * G * R + 37)  # several strips, ragged last ones, tiny pairs too
    s, _ = oracle.make_scoring(-5, -1, 1, -1)
    ref, ref_ops = oracle_batch(oracle, mode, s, batch, threads=4)
    got, ops = sim_util.align_batch(MODES[mode], s, *batch, R=R, G=G)
    assert_same(got, ops, ref, ref_ops, batch, f"wavefront {G}x{R} {mode}")


@pytest.mark.parametrize("G,R", WAVE_SHAPES)
def test_sim_wavefront_shapes_uniform_and_custom(oracle, G, R):
    rng = np.random.default_rng(7 * G + R)
    batch = synth.uniform_pairs(99, 0, 33, 2 * G * R + 5, 70)  # uniform block: the unmasked strip variant
    pick = lambda: int(rng.choice([MIN, 0, -2, -9]))
    s, _ = oracle.make_scoring(-3, -1, 2, -2, None, pick(), pick(), pick(), pick())
    ref, ref_ops = oracle_batch(oracle, "custom", s, batch, threads=4)
    for no_pack in (0, 1):
        got, ops = sim_util.align_batch(MODES["custom"], s, *batch, R=R, G=G, no_pack=no_pack)
        assert_same(got, ops, ref, ref_ops, batch, f"wavefront custom {G}x{R} no_pack={no_pack}")


@pytest.mark.parametrize("G,R", [(8, 8), (8, 20)])
def test_sim_uniform_masked_strips_every_capture_row(oracle, G, R):
    """Uniform blocks run their masked (last) strip with the capture row of the boundary hand-off dispatched at
    compile time by row-quad, and with LUT scoring the padded rows read the poison LUT row instead of masking the
    column tracker: every number of valid rows in the partial lane (incl. none), LUT and MatchParams scoring,
    packed and explicit trackers, local / custom clips."""
    rng = np.random.default_rng(G * R)
    span = G * R
    ms = sorted(set(list(range(span - R, span + 3)) + [span + R // 2, 2 * span, 2 * span + 1, 2, 3, R + 1, R + 2]))
    for k, m in enumerate(ms):
        batch = synth.uniform_pairs(1000 + m, 0, 5, m, 37 + (k % 5))
        clips = [int(rng.choice([MIN, 0, -3])) for _ in range(4)] if k % 2 else [0, 0, 0, 0]
        s, _ = oracle.make_scoring(-5 if k % 3 else 0, -1, 1, -1, None, *clips)
        ref, ref_ops = oracle_batch(oracle, "custom", s, batch, threads=4)
        got, ops = sim_util.align_batch(MODES["custom"], s, *batch, R=R, G=G, no_pack=k % 2, no_lut=(k % 4 == 3))
        assert_same(got, ops, ref, ref_ops, batch, f"uniform masked {G}x{R} m={m}")


@pytest.mark.parametrize("mode", ["local", "global", "custom"])
def test_sim_strip_pipelined_warp_per_pair(oracle, mode):
    """The warp-per-pair shape as the engine schedules it: one emulated warp per (pair, strip) task, the strips of a
    pair running side by side and handing the boundary row over through progress words (G=132 selects it)."""
    rng = np.random.default_rng(len(mode))
    batch = synth.ragged_pairs(500 + len(mode), 9, 200, 700)  # 1-3 strips of 256 rows, ragged block
    clips = [int(rng.choice([MIN, 0, -4])) for _ in range(4)] if mode == "custom" else [MIN] * 4
    s, _ = oracle.make_scoring(-5, -1, 1, -1, None, *clips)
    ref, ref_ops = oracle_batch(oracle, mode, s, batch, threads=4)
    got, ops = sim_util.align_batch(MODES[mode], s, *batch, R=8, G=132)
    assert_same(got, ops, ref, ref_ops, batch, f"strip-pipelined 32x8 {mode}")
    # and a uniform block (the unmasked strip variant), long enough for several 16-column publishes
    ubatch = synth.uniform_pairs(5, 0, 3, 600, 300)
    ref, ref_ops = oracle_batch(oracle, mode, s, ubatch, threads=4)
    got, ops = sim_util.align_batch(MODES[mode], s, *ubatch, R=8, G=132)
    assert_same(got, ops, ref, ref_ops, ubatch, f"strip-pipelined uniform {mode}")


# ---- the warp-per-pair K2 (walk_pair_coop<32>: row m, the fix-up passes as 32-wide prefix maxima) on 32 emulated
# lanes, against the oracle.  Same inputs as the one-lane walk above.

@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_sim_warp_walk_golden(oracle, case):
    s, keep = _scoring(oracle, case["scoring"])
    got, ops = sim_util.align_batch(MODES[case["mode"]], s, *_one(case["x"].encode(), case["y"].encode()), R=4,
                                    warp_walk=1)
    exp = case["expect"]
    for k in ("score", "xstart", "xend", "ystart", "yend"):
        if k in exp:
            assert int(got[k][0]) == exp[k], (case["name"], k)
    if "ops" in exp:
        assert ops[0] == parse_ops(exp["ops"])
    assert got["status"][0] == 0


@pytest.mark.parametrize("mode", ["local", "global", "semiglobal"])
@pytest.mark.parametrize("R,G,no_pack", [(4, 1, 0), (16, 1, 1), (16, 4, 0), (8, 32, 0)])
def test_sim_warp_walk_ragged_presets(oracle, mode, R, G, no_pack):
    batch = synth.ragged_pairs(31 + R, 160, 100, 130)
    s, _ = oracle.make_scoring(-5, -1, 1, -1)
    ref, ref_ops = oracle_batch(oracle, mode, s, batch)
    got, ops = sim_util.align_batch(MODES[mode], s, *batch, R=R, G=G, no_pack=no_pack, warp_walk=1)
    assert_same(got, ops, ref, ref_ops, batch, f"warp walk {mode} R={R} G={G}")


@pytest.mark.parametrize("seed", range(10))
def test_sim_warp_walk_random_custom_clips(oracle, seed):
    rng = np.random.default_rng(900 + seed)
    pick = lambda: int(rng.choice([MIN, 0, 0, -1, -3, -7, -20]))
    go, ge = int(rng.choice([0, -1, -2, -5, -6])), int(rng.choice([0, -1, -1, -2]))
    ma, mi = int(rng.choice([1, 2, 4])), int(rng.choice([-1, -3, -7, 0]))
    clips = (pick(), pick(), pick(), pick())
    s, _ = oracle.make_scoring(go, ge, ma, mi, None, *clips)
    batch = synth.ragged_pairs(seed, 150, 75, 95, alphabet=b"AC" if seed % 2 else b"ACGT")
    ref, ref_ops = oracle_batch(oracle, "custom", s, batch)
    got, ops = sim_util.align_batch(MODES["custom"], s, *batch, R=8, warp_walk=1)
    assert_same(got, ops, ref, ref_ops, batch, f"warp walk custom seed={seed} clips={clips}")


def test_sim_warp_walk_tiny_shapes(oracle):
    xs, ys = [], []
    for m in range(0, 6):
        for n in range(0, 6):
            for rep in range(2):
                xs.append(m)
                ys.append(n)
    for m, n in [(33, 1), (1, 33), (2, 64), (64, 2), (32, 32), (33, 33), (65, 31)]:
        xs.append(m)
        ys.append(n)
    rng = np.random.default_rng(5)
    total = sum(xs) + sum(ys)
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 2, size=total + 1)]
    lens = np.array([v for pair in zip(xs, ys) for v in pair], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    batch = (blob, offs[0::2].copy(), np.array(xs, dtype=np.uint32), offs[1::2].copy(), np.array(ys, dtype=np.uint32))
    for mode in ("custom", "local", "global", "semiglobal"):
        for clips in [(MIN, MIN, MIN, MIN), (0, 0, 0, 0), (-1, 0, MIN, -2), (0, MIN, MIN, 0)]:
            s, _ = oracle.make_scoring(-2, -1, 2, -1, None, *clips)
            ref, ref_ops = oracle_batch(oracle, mode, s, batch)
            got, ops = sim_util.align_batch(MODES[mode], s, *batch, R=4, warp_walk=1)
            assert_same(got, ops, ref, ref_ops, batch, f"warp walk tiny {mode} {clips}")


def test_sim_relative_packed_trackers_long_sequence_form():
    """F_PACKREL (what BASELINE config 5's 10k x 10k pairs run): exercised on small inputs by a host build whose
    column chunks are 32 columns long -- its own process, the chunk length is a compile-time constant."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, B2A_SIM_KREL_BITS="5")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "workers", "sim_relpack_worker.py")], env=env,
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and r.stdout.strip().startswith("OK"), r.stdout[-2000:] + r.stderr[-4000:]


def test_tracker_form_by_shape_and_score_range():
    """scoring_flags (b2a_plan.h): trackers as packed keys over absolute indices for reads (m, n <= 4095, scores
    < 2^17), over chunk- / strip-relative indices for longer sequences (scores < 2^18: BASELINE config 5), as explicit
    (value, index) pairs beyond; no tracker flags at all in global mode."""
    import ctypes as C
    L = sim_util.lib()
    L.sim_scoring_flags.restype = C.c_int
    MINS = -858993459
    F_TR, F_TC, F_CX, F_LUT, F_PK, F_RELU, F_PR = 1, 2, 4, 8, 16, 32, 128
    f = lambda clips, alpha, bound, m, n: L.sim_scoring_flags(*[C.c_int32(c) for c in clips], C.c_int32(alpha), C.c_int64(bound),
                                                           C.c_uint32(m), C.c_uint32(n))
    local, glob, semi = (0, 0, 0, 0), (MINS,) * 4, (MINS, MINS, 0, 0)
    assert f(local, 4, 1500, 150, 150) == F_TR | F_TC | F_CX | F_LUT | F_RELU | F_PK       # C2: reads
    assert f(local, 25, 220032, 10000, 10000) == F_TR | F_TC | F_CX | F_LUT | F_RELU | F_PR  # C5: 10k x 10k BLOSUM62
    assert f(local, 25, 1 << 19, 10000, 10000) == F_TR | F_TC | F_CX | F_LUT | F_RELU        # scores beyond 2^18: explicit pairs
    assert f(local, 4, 1 << 17, 150, 150) & (F_PK | F_PR) == F_PR                              # short reads, wide scores
    assert f(glob, 4, 10000, 1000, 1000) == F_LUT                                               # C3: no trackers
    assert f(semi, 0, 3000, 150, 5000) == F_TR | F_PR and f(semi, 0, 3000, 150, 4000) == F_TR | F_PK
