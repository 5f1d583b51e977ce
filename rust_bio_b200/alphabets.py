"""Mirror of `bio::alphabets::{Alphabet, RankTransform}` (reference src/alphabets/mod.rs:30-435), the part of it that
sits in front of the alignment path: texts -> ranks -> BitEnc (-> b2a_align_batch_packed)."""
from __future__ import annotations

import math

import numpy as np

from .data_structures.bitenc import BitEnc


class Alphabet:
    """mod.rs:30-215: a set of byte symbols (iteration in ascending symbol order, like the reference's BitSet)."""

    def __init__(self, symbols=b""):
        self.symbols = set(int(c) for c in bytes(symbols))

    @staticmethod
    def new(symbols) -> "Alphabet":
        return Alphabet(symbols)

    def insert(self, a: int) -> None:
        self.symbols.add(int(a))

    def is_word(self, text) -> bool:
        return all(int(c) in self.symbols for c in bytes(text))

    def max_symbol(self):
        return max(self.symbols) if self.symbols else None

    def len(self) -> int:
        return len(self.symbols)

    def __len__(self) -> int:
        return len(self.symbols)

    def is_empty(self) -> bool:
        return not self.symbols

    def intersection(self, other: "Alphabet") -> "Alphabet":
        return Alphabet(bytes(sorted(self.symbols & other.symbols)))

    def difference(self, other: "Alphabet") -> "Alphabet":
        return Alphabet(bytes(sorted(self.symbols - other.symbols)))

    def union(self, other: "Alphabet") -> "Alphabet":
        return Alphabet(bytes(sorted(self.symbols | other.symbols)))


class RankTransform:
    """mod.rs:220-435: symbol -> rank among the alphabet's symbols in ascending order."""

    def __init__(self, alphabet: Alphabet):
        self.ranks = {c: r for r, c in enumerate(sorted(alphabet.symbols))}
        self._table = np.full(256, 255, dtype=np.uint8)
        for c, r in self.ranks.items():
            self._table[c] = r

    @staticmethod
    def new(alphabet: Alphabet) -> "RankTransform":
        return RankTransform(alphabet)

    def get(self, a: int) -> int:
        if int(a) not in self.ranks:
            raise KeyError("Unexpected character.")  # the reference panics with this message (mod.rs:264)
        return self.ranks[int(a)]

    def transform(self, text) -> np.ndarray:
        t = np.frombuffer(bytes(text), dtype=np.uint8)
        if len(t) and not all(int(c) in self.ranks for c in np.unique(t)):
            raise KeyError("Unexpected character.")
        return self._table[t]

    def alphabet(self) -> Alphabet:
        return Alphabet(bytes(sorted(self.ranks)))

    def get_width(self) -> int:  # mod.rs:430-432
        return int(math.ceil(math.log2(len(self.ranks)))) if self.ranks else 0

    def bitenc(self, text) -> BitEnc:
        """transform(text) bit-encoded at get_width() bits per symbol (how the reference's docs combine the two)."""
        return BitEnc.from_values(max(1, self.get_width()), self.transform(text))
