"""Deterministic synthetic batches: splitmix64, one draw per symbol.

state starts at the seed; each draw does state += 0x9E3779B97F4A7C15 and mixes;
symbol = alphabet[(z >> 32) % len(alphabet)]; pair p uses seed BASE+2p for x
and BASE+2p+1 for y.  Vectorised: draw k of a sequence sees state seed+(k+1)*G.
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
DNA = b"ACGT"
PROTEIN = b"ACDEFGHIKLMNPQRSTVWY"

BASES = {"C1": 0xB2000001, "C2": 0xB2000002, "C3": 0xB2000003, "C4": 0xB2000004, "C5": 0xB2000005}


def _mix(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def draws(seeds: np.ndarray, length: int) -> np.ndarray:
    """uint64 [len(seeds), length] of splitmix64 outputs."""
    with np.errstate(over="ignore"):
        k = (np.arange(1, length + 1, dtype=np.uint64) * GOLDEN)[None, :]
        return _mix(seeds.astype(np.uint64)[:, None] + k)


def random_seqs(seeds: np.ndarray, length: int, alphabet: bytes) -> np.ndarray:
    """uint8 [len(seeds), length]."""
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    out = np.empty((len(seeds), length), dtype=np.uint8)
    step = max(1, (1 << 22) // max(1, length))
    for lo in range(0, len(seeds), step):
        z = draws(seeds[lo:lo + step], length)
        out[lo:lo + step] = alpha[((z >> np.uint64(32)) % np.uint64(len(alpha))).astype(np.int64)]
    return out


def uniform_pairs(base: int, first_pair: int, n_pairs: int, m: int, n: int, alphabet: bytes = DNA,
                  align: int = 16):
    """Batch in the C-ABI input layout: (blob, x_off, x_len, y_off, y_len); each sequence starts on an
    `align`-byte boundary (x block then y block per pair)."""
    p = np.arange(first_pair, first_pair + n_pairs, dtype=np.uint64)
    with np.errstate(over="ignore"):
        xs = random_seqs(np.uint64(base) + np.uint64(2) * p, m, alphabet)
        ys = random_seqs(np.uint64(base) + np.uint64(2) * p + np.uint64(1), n, alphabet)
    ms = -(-m // align) * align
    ns = -(-n // align) * align
    blob = np.zeros((n_pairs, ms + ns), dtype=np.uint8)
    blob[:, :m] = xs
    blob[:, ms:ms + n] = ys
    stride = ms + ns
    x_off = (np.arange(n_pairs, dtype=np.uint64) * np.uint64(stride))
    y_off = x_off + np.uint64(ms)
    x_len = np.full(n_pairs, m, dtype=np.uint32)
    y_len = np.full(n_pairs, n, dtype=np.uint32)
    return blob.reshape(-1), x_off, x_len, y_off, y_len


def ragged_pairs(seed: int, n_pairs: int, max_m: int, max_n: int, alphabet: bytes = DNA, min_len: int = 0):
    """Ragged batch (lengths uniform in [min_len, max]) for edge-case parity tests; unaligned offsets."""
    rng = np.random.default_rng(seed)
    x_len = rng.integers(min_len, max_m + 1, size=n_pairs).astype(np.uint32)
    y_len = rng.integers(min_len, max_n + 1, size=n_pairs).astype(np.uint32)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    total = int(x_len.sum() + y_len.sum())
    blob = alpha[rng.integers(0, len(alpha), size=total + 1)]
    lens = np.stack([x_len, y_len], axis=1).reshape(-1).astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    return blob, offs[0::2].copy(), x_len, offs[1::2].copy(), y_len
