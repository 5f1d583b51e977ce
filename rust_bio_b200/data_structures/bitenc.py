"""Mirror of `bio::data_structures::bitenc::BitEnc` (reference src/data_structures/bitenc.rs:45-434): a sequence of
`width`-bit values in 32-bit blocks, 32 - 32 % width usable bits per block.  The engine takes this storage as it
is (`Aligner.*_batch_bitenc`, b2a_align_batch_packed): the packed blocks are what crosses PCIe.

Same method names as the reference; `storage` is a numpy uint32 array so that a batch of BitEncs can be handed to
the C ABI without repacking.
"""
from __future__ import annotations

import numpy as np


class BitEnc:
    def __init__(self, width: int, capacity: int = 0):
        assert width <= 8, "Only encoding widths up to 8 supported"
        assert width >= 1
        self.width = int(width)
        self.mask = (1 << width) - 1
        self.usable_bits_per_block = 32 - 32 % width
        self._len = 0
        self._storage = np.zeros(max(1, capacity * width // 32 + 1), dtype=np.uint32)
        self._blocks = 0

    # bitenc.rs:75, 99
    @staticmethod
    def new(width: int) -> "BitEnc":
        return BitEnc(width)

    @staticmethod
    def with_capacity(width: int, n: int) -> "BitEnc":
        return BitEnc(width, n)

    @staticmethod
    def from_values(width: int, values) -> "BitEnc":
        """All of `values` at once (vectorised `push` for every element)."""
        b = BitEnc(width)
        v = np.asarray(values, dtype=np.uint32) & np.uint32(b.mask)
        n = len(v)
        per = b.usable_bits_per_block // width
        nblocks = (n + per - 1) // per
        pad = np.zeros(nblocks * per, dtype=np.uint32)
        pad[:n] = v
        shifts = (np.arange(per, dtype=np.uint32) * np.uint32(width))
        b._storage = np.bitwise_or.reduce(pad.reshape(nblocks, per) << shifts[None, :], axis=1).astype(np.uint32) \
            if nblocks else np.zeros(1, dtype=np.uint32)
        b._blocks = nblocks
        b._len = n
        return b

    def _addr(self, i: int):  # bitenc.rs:332-338
        k = i * self.width
        return k // self.usable_bits_per_block, k % self.usable_bits_per_block

    def _grow(self, blocks: int):
        if blocks > len(self._storage):
            new = np.zeros(max(blocks, 2 * len(self._storage)), dtype=np.uint32)
            new[:self._blocks] = self._storage[:self._blocks]
            self._storage = new

    def push(self, value: int) -> None:  # bitenc.rs:127-134
        block, bit = self._addr(self._len)
        if bit == 0:
            self._grow(self._blocks + 1)
            self._storage[self._blocks] = 0
            self._blocks += 1
        self._set_by_addr(block, bit, value)
        self._len += 1

    def push_values(self, n: int, value: int) -> None:  # bitenc.rs:175-230 (same resulting storage)
        for _ in range(n):
            self.push(value)

    def set(self, i: int, value: int) -> None:  # bitenc.rs:246-249
        block, bit = self._addr(i)
        self._set_by_addr(block, bit, value)

    def get(self, i: int):  # bitenc.rs:266-273
        if i >= self._len:
            return None
        block, bit = self._addr(i)
        return int((int(self._storage[block]) >> bit) & self.mask)

    def _set_by_addr(self, block: int, bit: int, value: int) -> None:  # bitenc.rs:324-329
        cur = int(self._storage[block])
        cur &= ~(self.mask << bit) & 0xFFFFFFFF
        cur |= (int(value) & self.mask) << bit
        self._storage[block] = cur

    def iter(self):
        return (self.get(i) for i in range(self._len))

    def __iter__(self):
        return self.iter()

    def clear(self) -> None:
        self._len = 0
        self._blocks = 0

    def len(self) -> int:
        return self._len

    def __len__(self) -> int:
        return self._len

    def nr_blocks(self) -> int:
        return self._blocks

    def nr_symbols(self) -> int:
        return self._len

    def is_empty(self) -> bool:
        return self._len == 0

    @property
    def storage(self) -> np.ndarray:
        """The 32-bit blocks (a view of nr_blocks() entries)."""
        return self._storage[:self._blocks]

    def to_values(self) -> np.ndarray:
        per = self.usable_bits_per_block // self.width
        shifts = (np.arange(per, dtype=np.uint32) * np.uint32(self.width))
        v = (self.storage[:, None] >> shifts[None, :]) & np.uint32(self.mask)
        return v.reshape(-1)[:self._len].astype(np.uint8)

    def __eq__(self, other) -> bool:
        return (isinstance(other, BitEnc) and self.width == other.width and self._len == other._len
                and np.array_equal(self.storage, other.storage))
