"""Records -> the batch buffers of `b2a_pairs` (include/b200align.h): one byte blob with 16-byte aligned sequences,
offsets and lengths; pinned host memory when asked (the e2e path copies straight from it)."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

from ..engine import pack_pairs


def pairs_from_records(queries: Iterable, targets: Iterable) -> List[Tuple[bytes, bytes]]:
    """Zip two record streams (e.g. reads and the windows they are aligned to) into (x, y) byte pairs."""
    return [(q.seq(), t.seq()) for q, t in zip(queries, targets)]


def records_to_batch(queries: Sequence, targets: Sequence, pinned: bool = False):
    """-> (batch, keep): `batch` = (blob, x_off, x_len, y_off, y_len) as the engine takes it; with pinned=True the
    arrays live in page-locked memory (torch) and `keep` holds the tensors that own it."""
    if len(queries) != len(targets):
        raise ValueError("one target per query")
    batch = pack_pairs(pairs_from_records(queries, targets))
    if not pinned:
        return batch, None
    import torch
    keep = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in batch]
    return tuple(t.numpy() for t in keep), keep
