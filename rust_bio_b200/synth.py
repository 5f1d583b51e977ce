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


def mutated_window_pairs(base: int, first_pair: int, n_pairs: int, m: int, n: int, sub: float = 0.05,
                         ins: float = 0.005, dele: float = 0.005, alphabet: bytes = DNA, align: int = 16):
    """The C4 generator: y uniform over `alphabet` (seed BASE+2p+1); x = a window of y at a random
    start with `sub` substitutions, `ins` insertions and `dele` deletions per source position, cut to exactly m
    symbols.  All randomness is splitmix64 of seed BASE+2p: draw 0 picks the window start, draws 1..L the event of
    each source position, draws L+1..2L the inserted / substituted symbol.  Vectorised over pairs."""
    A = len(alphabet)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    code_of = np.zeros(256, dtype=np.int64)
    code_of[alpha] = np.arange(A)
    L = m + max(32, m // 8)
    if n < L:
        raise ValueError("y must be at least m + max(32, m/8) long")
    ms = -(-m // align) * align
    ns = -(-n // align) * align
    stride = ms + ns
    blob = np.zeros((n_pairs, stride), dtype=np.uint8)
    step = max(1, (1 << 21) // (2 * L + 1))
    for lo in range(0, n_pairs, step):
        hi = min(n_pairs, lo + step)
        p = np.arange(first_pair + lo, first_pair + hi, dtype=np.uint64)
        with np.errstate(over="ignore"):
            ys = random_seqs(np.uint64(base) + np.uint64(2) * p + np.uint64(1), n, alphabet)
            z = draws(np.uint64(base) + np.uint64(2) * p, 2 * L + 1)
        c = hi - lo
        start = ((z[:, 0] >> np.uint64(32)) % np.uint64(n - L + 1)).astype(np.int64)
        src = np.take_along_axis(ys, start[:, None] + np.arange(L, dtype=np.int64)[None, :], axis=1)
        u = (z[:, 1:L + 1] >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        r = ((z[:, L + 1:] >> np.uint64(32)) % np.uint64(A)).astype(np.int64)
        is_sub = u < sub
        is_ins = (u >= sub) & (u < sub + ins)
        is_del = (u >= sub + ins) & (u < sub + ins + dele)
        # a substitution always changes the symbol: code + 1 + (r mod (A-1))
        sub_sym = alpha[(code_of[src] + 1 + r % max(1, A - 1)) % A]
        sym = np.where(is_sub, sub_sym, src)
        cnt = 1 + is_ins.astype(np.int64) - is_del.astype(np.int64)
        pos = np.cumsum(cnt, axis=1) - cnt
        xs = np.zeros((c, m + 2), dtype=np.uint8)
        rows = np.broadcast_to(np.arange(c)[:, None], (c, L))
        keep = (~is_del) & (pos + is_ins < m)
        xs[rows[keep], (pos + is_ins)[keep]] = sym[keep]
        insm = is_ins & (pos < m)
        xs[rows[insm], pos[insm]] = alpha[r[insm]]
        total = pos[:, -1] + cnt[:, -1]
        if np.any(total < m):  # cannot happen for sane rates (needs > L - m deletions); pad to stay well-formed
            for q in np.nonzero(total < m)[0]:
                xs[q, int(total[q]):m] = alpha[0]
        blob[lo:hi, :m] = xs[:, :m]
        blob[lo:hi, ms:ms + n] = ys
    x_off = np.arange(n_pairs, dtype=np.uint64) * np.uint64(stride)
    y_off = x_off + np.uint64(ms)
    return (blob.reshape(-1), x_off, np.full(n_pairs, m, dtype=np.uint32), y_off,
            np.full(n_pairs, n, dtype=np.uint32))
