"""Host-side logic of the mirrors (no GPU): what is sent across the C ABI."""
import numpy as np

from rust_bio_b200 import dist as bdist, scores, synth
from rust_bio_b200.engine import pack_pairs
from rust_bio_b200.pairwise import Scoring, _symbols_present


def test_closure_is_tabulated_over_sequence_bytes_only():
    """pack_pairs pads every sequence with zero bytes; the reference only ever calls match_fn on sequence
    bytes (mod.rs:733), so a closure that cannot score byte 0 (bio::scores::blosum62 indexes out of bounds
    there, scores/mod.rs:22-35) must still work, and byte 0 must not take an alphabet slot."""
    batch = pack_pairs([(b"ACDEFGHIKLMNPQRS", b"ACDEFG"), (b"WY", b"")])
    cs, keep = Scoring.new(-5, -1, lambda a, b: scores.blosum62(a, b)).to_c(batch)
    alpha = bytes(keep[1])
    assert alpha == bytes(sorted(set(b"ACDEFGHIKLMNPQRSWY")))
    assert cs.alphabet_len == len(alpha)
    table = keep[0].reshape(256, 256)
    assert table[ord("A"), ord("A")] == 4 and table[ord("W"), ord("W")] == 11
    # a dict-based closure: KeyError on any byte it was not built for
    d = {(a, b): (2 if a == b else -3) for a in b"ACGT" for b in b"ACGT"}
    batch = pack_pairs([(b"ACGTTGCA", b"GGTTAACC")])
    cs, keep = Scoring.new(-5, -1, lambda a, b: d[(a, b)]).to_c(batch)
    assert bytes(keep[1]) == b"ACGT"


def test_symbols_present_ignores_gaps_between_sequences():
    blob = np.frombuffer(b"zzACGTzzzzTTzz", dtype=np.uint8)
    batch = (blob, np.array([2], np.uint64), np.array([4], np.uint32), np.array([10], np.uint64), np.array([2], np.uint32))
    assert bytes(_symbols_present(batch)) == b"ACGT"
    empty = (blob, np.array([0], np.uint64), np.array([0], np.uint32), np.array([0], np.uint64), np.array([0], np.uint32))
    assert len(_symbols_present(empty)) == 0


def test_shard_batch_sends_only_the_bytes_of_the_shard():
    """A blob laid out as all x, then all y: every rank's shard is gathered into a compact blob instead of
    re-sending (almost) the whole blob to every rank; contiguous layouts are cut to the shard's span."""
    n, m = 256, 40000
    rng = np.random.default_rng(3)
    xs = rng.integers(65, 69, size=(n, m), dtype=np.uint8)
    ys = rng.integers(65, 69, size=(n, m), dtype=np.uint8)
    blob = np.concatenate([xs.reshape(-1), ys.reshape(-1)])
    x_off = (np.arange(n, dtype=np.uint64) * np.uint64(m))
    y_off = x_off + np.uint64(n * m)
    lens = np.full(n, m, dtype=np.uint32)
    batch = (blob, x_off, lens, y_off, lens)
    for rank in range(4):
        (b, xo, xl, yo, yl), lo, hi = bdist.shard_batch(batch, 4, rank)
        assert len(b) < len(blob) // 3
        for i in range(hi - lo):
            assert np.array_equal(b[int(xo[i]):int(xo[i]) + m], xs[lo + i])
            assert np.array_equal(b[int(yo[i]):int(yo[i]) + m], ys[lo + i])
    batch = synth.uniform_pairs(1, 0, 100, 30, 50)
    (b, xo, xl, yo, yl), lo, hi = bdist.shard_batch(batch, 4, 2)
    assert (lo, hi) == (50, 75) and len(b) <= 25 * (32 + 64)
    assert np.array_equal(b[int(yo[3]):int(yo[3]) + 50], batch[0][int(batch[3][53]):int(batch[3][53]) + 50])
