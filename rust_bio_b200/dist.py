"""Multi-GPU form of the path: split the pair list over ranks, one all-gather of result records.

The reference has no distributed layer; pairs are independent, so the only exchange is
the reassembly of per-pair results (BASELINE north_star: "a single NCCL allgather over NVLink only
to reassemble per-pair scores/CIGARs").  One process per GPU under torch.distributed; records are the
fixed-stride layout of b2a_batch_records (include/b200align.h).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np

RECORD_HEAD = 40


def shard_range(n_pairs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split with equal counts (the last shards may be one shorter / empty)."""
    per = -(-n_pairs // world)
    lo = min(n_pairs, rank * per)
    return lo, min(n_pairs, lo + per)


def shard_batch(batch, world: int, rank: int):
    blob, x_off, x_len, y_off, y_len = batch
    lo, hi = shard_range(len(x_len), world, rank)
    return (blob, x_off[lo:hi].copy(), x_len[lo:hi].copy(), y_off[lo:hi].copy(), y_len[lo:hi].copy()), lo, hi


def record_stride(max_m: int, max_n: int) -> int:
    """== b2a_record_stride"""
    return RECORD_HEAD + ((max_m + max_n + 4 + 15) & ~15)


def encode_records(fields: dict, ops_lists, stride: int) -> np.ndarray:
    """Host-side encoder of the record layout (used by tests to fake a rank's device output)."""
    n = len(ops_lists)
    rec = np.zeros((n, stride), dtype=np.uint8)
    head = rec[:, :RECORD_HEAD].view(np.uint32)
    head[:, 0] = np.asarray(fields["score"]).astype(np.int32).view(np.uint32)
    for k, name in enumerate(("xstart", "xend", "ystart", "yend")):
        head[:, 1 + k] = fields[name]
    for p, ops in enumerate(ops_lists):
        head[p, 5] = len(ops)
        clips = [l for c, l in ops if c >= 4]
        for k, l in enumerate(clips[:4]):
            head[p, 6 + k] = l
        rec[p, RECORD_HEAD:RECORD_HEAD + len(ops)] = [c for c, _ in ops]
    return rec.reshape(-1)


def decode_records(rec: np.ndarray, stride: int, n: int):
    """Pure-numpy decoder (the C ABI has b2a_records_decode for the same job)."""
    rec = rec[:n * stride].reshape(n, stride)
    head = rec[:, :RECORD_HEAD].view(np.uint32)
    fields = {"score": head[:, 0].view(np.int32).copy(), "xstart": head[:, 1].copy(), "xend": head[:, 2].copy(),
              "ystart": head[:, 3].copy(), "yend": head[:, 4].copy()}
    ops_lists = []
    for p in range(n):
        codes = rec[p, RECORD_HEAD:RECORD_HEAD + int(head[p, 5])]
        out, k = [], 0
        for c in codes:
            c = int(c)
            if c >= 4:
                out.append((c, int(head[p, 6 + k])))
                k += 1
            else:
                out.append((c, 0))
        ops_lists.append(out)
    return fields, ops_lists


def all_gather_records(local_rec, n_local: int, n_total: int, stride: int, world: int):
    """local_rec: torch uint8 tensor [per * stride] on this rank's device (padded to the common shard
    size `per`).  Returns the gathered [world * per * stride] tensor (one collective)."""
    import torch
    import torch.distributed as dist
    per = -(-n_total // world)
    assert local_rec.numel() == per * stride
    out = torch.empty(world * per * stride, dtype=torch.uint8, device=local_rec.device)
    dist.all_gather_into_tensor(out, local_rec)
    return out


def align_sharded(batch, stride: int, run_local: Callable, device="cpu"):
    """Shard `batch` over the ranks of the default process group, run `run_local(shard)` -> uint8 torch
    tensor of that shard's records (n_shard * stride bytes, on `device`), all-gather once, and return
    (fields, ops_lists) for the WHOLE batch in the caller's order, on every rank."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    n = len(batch[2])
    shard, lo, hi = shard_batch(batch, world, rank)
    per = -(-n // world)
    local = torch.zeros(per * stride, dtype=torch.uint8, device=device)
    if hi > lo:
        rec = run_local(shard)
        local[:(hi - lo) * stride] = rec
    allrec = all_gather_records(local, hi - lo, n, stride, world)
    host = allrec.cpu().numpy()
    # rank r's shard occupies [r*per, r*per + len_r): contiguous split => caller order is preserved
    return decode_records(host, stride, n)
