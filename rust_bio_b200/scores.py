"""Substitution matrices of the path: mirror of `bio::scores` (reference src/scores/).

`lookup` follows src/scores/mod.rs:22-35; `blosum62(a, b)` follows
src/scores/blosum62.rs:54-58 (MAT[27*a + b]); the other matrices have the same
shape (src/scores/{blosum30,blosum45,pam40,pam120,pam200,pam250}.rs).
"""
from __future__ import annotations

import numpy as np

from ._score_tables import MATRICES, SYMBOLS

__all__ = ["lookup", "blosum30", "blosum45", "blosum62", "pam40", "pam120", "pam200", "pam250",
           "tabulate", "matrix_table256"]


def lookup(a: int) -> int:
    """Letter -> matrix index (scores/mod.rs:22-35). Bytes outside A-Z/'*' index out of bounds there."""
    if a == ord("Y"):
        return 23
    if a == ord("Z"):
        return 24
    if a == ord("X"):
        return 25
    if a == ord("*"):
        return 26
    idx = a - 65
    if not 0 <= idx < 27:
        raise IndexError("bio::scores::lookup: byte %d outside the matrix alphabet (the reference panics)" % a)
    return idx


def _flat(name: str) -> np.ndarray:
    rows = MATRICES[name]
    return np.array([rows[s] for s in SYMBOLS], dtype=np.int32)


_FLAT = {n: _flat(n) for n in MATRICES}


def _make(name: str):
    mat = _FLAT[name]

    def score(a: int, b: int) -> int:
        return int(mat[lookup(a), lookup(b)])

    score.__name__ = name
    score.__doc__ = f"{name.upper()} score of [a, b] (src/scores/{name}.rs)."
    score.matrix_name = name
    return score


blosum30 = _make("blosum30")
blosum45 = _make("blosum45")
blosum62 = _make("blosum62")
pam40 = _make("pam40")
pam120 = _make("pam120")
pam200 = _make("pam200")
pam250 = _make("pam250")


def matrix_table256(name: str) -> np.ndarray:
    """256x256 int32 table of a named matrix: entries for bytes outside the alphabet are 0 and must not be used."""
    t = np.zeros((256, 256), dtype=np.int32)
    valid = [b for b in range(256) if b == ord("*") or 65 <= b <= 90]
    idx = np.array([lookup(b) for b in valid])
    t[np.ix_(valid, valid)] = _FLAT[name][np.ix_(idx, idx)]
    return t


def tabulate(match_fn, symbols) -> np.ndarray:
    """Tabulate an arbitrary `Fn(u8,u8)->i32` MatchFunc (mod.rs:221-228) over the symbols present."""
    name = getattr(match_fn, "matrix_name", None)
    if name is not None:
        return matrix_table256(name)
    t = np.zeros((256, 256), dtype=np.int32)
    syms = sorted(set(int(s) for s in symbols))
    for a in syms:
        for b in syms:
            t[a, b] = int(match_fn(a, b))
    return t
