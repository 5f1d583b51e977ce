"""Mirror of `bio::alignment::pairwise::banded` (reference src/alignment/pairwise/banded.rs:122-1004).

`Aligner::{new, with_capacity, with_capacity_and_scoring, with_scoring}` take the k-mer length `k`
and the band half-width `w` like the reference (banded.rs:150-267); `custom / global_ / semiglobal /
local` (banded.rs:282, 872, 901, 975) and their `*_batch` forms run K4 (k-mer matches -> SDPk++ chain
-> Band) and K3 (banded fill + walk) on the GPU.  A band with more than MAX_CELLS = 5,000,000 cells
returns the reference's empty alignment (score MIN_SCORE, no operations, xlen = ylen = 0; banded.rs:407-420).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from ._lib import MIN_SCORE, MODE_CUSTOM, MODE_GLOBAL, MODE_LOCAL, MODE_SEMIGLOBAL
from .alignment import Alignment, AlignmentMode, AlignmentOperation
from .engine import Engine, default_engine, pack_pairs
from .pairwise import DEFAULT_ALIGNER_CAPACITY, MatchFunc, Scoring, _check_scoring

MAX_CELLS = 5_000_000          # banded.rs:104
DEFAULT_MATCH_SCORE = 2        # banded.rs:105


class Aligner:
    def __init__(self, scoring: Scoring, k: int, w: int, engine: Optional[Engine] = None):
        self.scoring, self.k, self.w, self._engine = scoring, int(k), int(w), engine

    @staticmethod
    def new(gap_open: int, gap_extend: int, match_fn: MatchFunc, k: int, w: int, engine=None) -> "Aligner":
        return Aligner.with_capacity(DEFAULT_ALIGNER_CAPACITY, DEFAULT_ALIGNER_CAPACITY, gap_open, gap_extend,
                                     match_fn, k, w, engine)

    @staticmethod
    def with_capacity(m: int, n: int, gap_open: int, gap_extend: int, match_fn: MatchFunc, k: int, w: int,
                      engine=None) -> "Aligner":
        return Aligner(Scoring.new(gap_open, gap_extend, match_fn), k, w, engine)

    @staticmethod
    def with_capacity_and_scoring(m: int, n: int, scoring: Scoring, k: int, w: int, engine=None) -> "Aligner":
        _check_scoring(scoring)
        return Aligner(scoring, k, w, engine)

    @staticmethod
    def with_scoring(scoring: Scoring, k: int, w: int, engine=None) -> "Aligner":
        return Aligner.with_capacity_and_scoring(DEFAULT_ALIGNER_CAPACITY, DEFAULT_ALIGNER_CAPACITY, scoring, k, w,
                                                 engine)

    def get_mut_scoring(self) -> Scoring:
        return self.scoring

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = default_engine(0)
        return self._engine

    def _batch(self, mode: int, pairs: Sequence[Tuple[bytes, bytes]]) -> List[Alignment]:
        batch = pack_pairs(pairs)
        cs, keep = self.scoring.to_c(batch[0])
        res = self.engine.align_batch_banded(mode, cs, self.k, self.w, batch)
        out = []
        for i, (x, y) in enumerate(pairs):
            ops = [AlignmentOperation(c, l) for c, l in res.ops_of(i)]
            refused = int(res.score[i]) == MIN_SCORE and not ops  # MAX_CELLS refusal, banded.rs:407-420
            out.append(Alignment(int(res.score[i]), int(res.ystart[i]), int(res.xstart[i]), int(res.yend[i]),
                                 int(res.xend[i]), 0 if refused else len(y), 0 if refused else len(x), ops,
                                 AlignmentMode.Custom if refused else mode))
        return out

    def custom_batch(self, pairs):
        return self._batch(MODE_CUSTOM, pairs)

    def global_batch(self, pairs):
        return self._batch(MODE_GLOBAL, pairs)

    def semiglobal_batch(self, pairs):
        return self._batch(MODE_SEMIGLOBAL, pairs)

    def local_batch(self, pairs):
        return self._batch(MODE_LOCAL, pairs)

    def custom(self, x: bytes, y: bytes) -> Alignment:
        return self._batch(MODE_CUSTOM, [(x, y)])[0]

    def global_(self, x: bytes, y: bytes) -> Alignment:
        return self._batch(MODE_GLOBAL, [(x, y)])[0]

    def semiglobal(self, x: bytes, y: bytes) -> Alignment:
        return self._batch(MODE_SEMIGLOBAL, [(x, y)])[0]

    def local(self, x: bytes, y: bytes) -> Alignment:
        return self._batch(MODE_LOCAL, [(x, y)])[0]


setattr(Aligner, "global", Aligner.global_)
