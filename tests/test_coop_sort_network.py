"""The compare-exchange network of coop_sort_u64 (rust_bio_b200/csrc/b2a_banded.cuh), restated step by step in
numpy: the all-ascending bitonic form with pairs that reach past n skipped must sort every length, not only
powers of two (the claim the device code relies on instead of padding its arrays)."""
import numpy as np


def network_steps(n):
    """The (mask) sequence the device code walks: per merge size k the mirror step k-1, then k/4, k/8, ..., 1."""
    k = 2
    while (k >> 1) < n:
        yield k - 1
        j = k >> 2
        while j > 0:
            yield j
            j >>= 1
        k <<= 1


def run_network(a):
    a = a.copy()
    n = len(a)
    for mask in network_steps(n):
        i = np.arange(n)
        j = i ^ mask
        sel = (j > i) & (j < n)
        lo, hi = i[sel], j[sel]
        # every index takes part in at most one pair per step, so the step is one vectorised exchange
        assert len(set(lo.tolist()) & set(hi.tolist())) == 0
        u, v = a[lo].copy(), a[hi].copy()
        swap = u > v
        a[lo] = np.where(swap, v, u)
        a[hi] = np.where(swap, u, v)
    return a


def test_network_sorts_every_length():
    rng = np.random.default_rng(0)
    for n in list(range(0, 140)) + [255, 256, 257, 1000, 1023, 1025]:
        for _ in range(3 if n > 140 else 6):
            a = rng.integers(0, 50 if n % 2 else 2 ** 62, size=n, dtype=np.uint64)  # with and without duplicates
            assert np.array_equal(run_network(a), np.sort(a)), n
    assert np.array_equal(run_network(np.arange(37, dtype=np.uint64)[::-1]), np.arange(37, dtype=np.uint64))
