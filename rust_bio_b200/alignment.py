"""Result types of the pairwise path: mirror of `bio_types::alignment`.

The reference re-exports these at src/alignment/mod.rs:14 and builds them at
src/alignment/pairwise/mod.rs:911-921.  bio-types is an external crate that is
not vendored in the reference tree; the variant orders used for the integer
codes follow bio-types 1.0.x.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

__all__ = [
    "AlignmentOperation", "AlignmentMode", "Alignment",
    "Match", "Subst", "Del", "Ins", "Xclip", "Yclip",
]

_OP_NAMES = ("Match", "Subst", "Del", "Ins", "Xclip", "Yclip")


@dataclass(frozen=True)
class AlignmentOperation:
    """`AlignmentOperation::{Match,Subst,Del,Ins,Xclip(usize),Yclip(usize)}`."""
    code: int
    len: int = 0

    def __repr__(self) -> str:
        n = _OP_NAMES[self.code]
        return f"{n}({self.len})" if self.code >= 4 else n


Match = AlignmentOperation(0)
Subst = AlignmentOperation(1)
Del = AlignmentOperation(2)
Ins = AlignmentOperation(3)


def Xclip(n: int) -> AlignmentOperation:
    return AlignmentOperation(4, int(n))


def Yclip(n: int) -> AlignmentOperation:
    return AlignmentOperation(5, int(n))


class AlignmentMode:
    """`AlignmentMode::{Custom,Global,Semiglobal,Local}` (codes == B2A_MODE_*)."""
    Custom = 0
    Global = 1
    Semiglobal = 2
    Local = 3
    names = ("Custom", "Global", "Semiglobal", "Local")


@dataclass
class Alignment:
    """`bio_types::alignment::Alignment` (fields as built at mod.rs:911-921)."""
    score: int
    ystart: int
    xstart: int
    yend: int
    xend: int
    ylen: int
    xlen: int
    operations: List[AlignmentOperation] = field(default_factory=list)
    mode: int = AlignmentMode.Custom

    def filter_clip_operations(self) -> None:
        """Drop Xclip/Yclip (call sites mod.rs:974,1006)."""
        self.operations = [o for o in self.operations if o.code < 4]

    def x_aln_len(self) -> int:
        return self.xend - self.xstart

    def y_aln_len(self) -> int:
        return self.yend - self.ystart

    def path(self) -> List[Tuple[int, int, AlignmentOperation]]:
        """(x_i, y_i, op) tuples, 1-based end coordinates, like bio-types `Alignment::path`."""
        out = []
        x_i, y_i = (self.xlen, self.ylen) if self.mode == AlignmentMode.Custom else (self.xend, self.yend)
        for op in reversed(self.operations):
            if op.code in (0, 1):
                out.append((x_i, y_i, op)); x_i -= 1; y_i -= 1
            elif op.code == 2:
                out.append((x_i, y_i, op)); y_i -= 1
            elif op.code == 3:
                out.append((x_i, y_i, op)); x_i -= 1
            elif op.code == 4:
                out.append((x_i, y_i, op)); x_i -= op.len
            else:
                out.append((x_i, y_i, op)); y_i -= op.len
        out.reverse()
        return out

    def cigar(self, hard_clip: bool = False) -> str:
        """CIGAR of x against y (bio-types `Alignment::cigar`; Semiglobal/Custom-global-x only there)."""
        clip = "H" if hard_clip else "S"
        sym = {0: "=", 1: "X", 2: "D", 3: "I"}
        parts: List[str] = []
        if self.xstart > 0:
            parts.append(f"{self.xstart}{clip}")
        run, last = 0, None
        for op in self.operations:
            if op.code >= 4:
                continue
            if op.code == last:
                run += 1
            else:
                if last is not None:
                    parts.append(f"{run}{sym[last]}")
                last, run = op.code, 1
        if last is not None:
            parts.append(f"{run}{sym[last]}")
        if self.xlen > self.xend:
            parts.append(f"{self.xlen - self.xend}{clip}")
        return "".join(parts)
