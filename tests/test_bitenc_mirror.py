"""Host mirrors of bio::data_structures::bitenc::BitEnc and bio::alphabets::{Alphabet, RankTransform} against the
reference's own doc examples (src/data_structures/bitenc.rs:20-43, 118-126, 160-173; src/alphabets/mod.rs:60-73,
250-262, 275-280, 420-428)."""
import numpy as np
import pytest

from rust_bio_b200.alphabets import Alphabet, RankTransform
from rust_bio_b200.data_structures import BitEnc


def test_bitenc_doc_examples():
    b = BitEnc.new(2)
    for v in (0, 2, 1):
        b.push(v)
    assert list(b.iter()) == [0, 2, 1] and b.nr_blocks() == 1 and b.nr_symbols() == 3
    b = BitEnc.new(4)
    for v in (0b0000, 0b1000, 0b1010):
        b.push(v)
    assert list(b.iter()) == [0b0000, 0b1000, 0b1010] and int(b.storage[0]) == 0b1010_1000_0000
    b = BitEnc.new(8)
    b.push_values(4, 0b101010)  # bitenc.rs:160-173
    b.push_values(2, 0b111111)
    assert b.nr_blocks() == 2 and list(b.iter())[:4] == [0b101010] * 4 and b.get(5) == 0b111111 and b.get(6) is None
    b.set(1, 7)
    assert b.get(1) == 7 and b.get(0) == 0b101010
    with pytest.raises(AssertionError, match="widths up to 8"):
        BitEnc.new(9)


@pytest.mark.parametrize("width", [1, 2, 3, 5, 7, 8])
def test_bitenc_addressing_and_bulk_constructor(width):
    """width 3, 5, 7: 32 % width != 0, so blocks have unused top bits (usable_bits_per_block, bitenc.rs:82)."""
    rng = np.random.default_rng(width)
    vals = rng.integers(0, 1 << width, size=301).astype(np.uint8)
    a = BitEnc.new(width)
    for v in vals:
        a.push(int(v))
    b = BitEnc.from_values(width, vals)
    assert a == b and a.nr_blocks() == -(-301 // ((32 - 32 % width) // width))
    assert np.array_equal(b.to_values(), vals) and [a.get(i) for i in (0, 17, 300)] == [int(vals[i]) for i in (0, 17, 300)]


def test_alphabet_and_rank_transform_doc_examples():
    dna = Alphabet.new(b"ACGTacgt")
    assert not dna.is_word(b"N")
    dna.insert(78)
    assert dna.is_word(b"N") and dna.is_word(b"GAttACA") and not dna.is_word(b"42")
    assert Alphabet.new(b"acgtACGT").max_symbol() == 116 and Alphabet.new(b"acgtACGT").len() == 8
    ranks = RankTransform.new(Alphabet.new(b"acgtACGT"))
    assert ranks.get(65) == 0 and ranks.get(116) == 7
    assert list(RankTransform.new(Alphabet.new(b"ACGTacgt")).transform(b"aAcCgGtT")) == [4, 0, 5, 1, 6, 2, 7, 3]
    assert RankTransform.new(Alphabet.new(b"ACGT")).get_width() == 2
    assert RankTransform.new(Alphabet.new(b"ACGTN")).get_width() == 3
    with pytest.raises(KeyError, match="Unexpected character"):
        ranks.get(ord("N"))
    enc = RankTransform.new(Alphabet.new(b"ACGT")).bitenc(b"GATTACA")
    assert enc.width == 2 and list(enc.iter()) == [2, 0, 3, 3, 0, 1, 0]
