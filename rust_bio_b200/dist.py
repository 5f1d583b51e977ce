"""Multi-GPU form of the path: split the pair list over ranks, one all-gather of result records.

The reference has no distributed layer; pairs are independent, so the only exchange is
the reassembly of per-pair results (BASELINE north_star: "a single NCCL allgather over NVLink only
to reassemble per-pair scores/CIGARs").  One process per GPU under torch.distributed.  Two wire formats
(include/b200align.h): fixed-stride records (b2a_batch_records) and the compact segment
(b2a_batch_compact_*: per-rank arrays + dense ops, ranks first agree on the largest segment with a MAX
all-reduce of one integer) which is what bench.py exchanges -- short alignments travel at their real length.
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
    """This rank's contiguous share of the pair list, with only the bytes it needs: the blob is cut to the
    span its sequences cover (offsets rebased), or gathered into a compact blob when they are scattered over
    a much larger span (e.g. a blob laid out as all x, then all y)."""
    blob, x_off, x_len, y_off, y_len = batch
    lo, hi = shard_range(len(x_len), world, rank)
    xo, xl, yo, yl = x_off[lo:hi].copy(), x_len[lo:hi].copy(), y_off[lo:hi].copy(), y_len[lo:hi].copy()
    if hi <= lo:
        return (blob[:0], xo, xl, yo, yl), lo, hi
    bmin = int(min(xo.min(), yo.min()))
    bmax = int(max((xo + xl.astype(np.uint64)).max(), (yo + yl.astype(np.uint64)).max()))
    need = int(xl.astype(np.uint64).sum() + yl.astype(np.uint64).sum())
    if bmax - bmin > 2 * need + (1 << 20):
        pad = lambda v: (v.astype(np.uint64) + np.uint64(15)) // np.uint64(16) * np.uint64(16)
        sizes = np.stack([pad(xl), pad(yl)], axis=1).reshape(-1)
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        out = np.zeros(int(offs[-1]) + 16, dtype=np.uint8)
        nxo, nyo = offs[0:-1:2].copy(), offs[1::2].copy()
        for i in range(hi - lo):
            out[int(nxo[i]):int(nxo[i]) + int(xl[i])] = blob[int(xo[i]):int(xo[i]) + int(xl[i])]
            out[int(nyo[i]):int(nyo[i]) + int(yl[i])] = blob[int(yo[i]):int(yo[i]) + int(yl[i])]
        return (out, nxo, xl, nyo, yl), lo, hi
    return (blob[bmin:bmax], xo - np.uint64(bmin), xl, yo - np.uint64(bmin), yl), lo, hi


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


# ---------------------------------------------------------------- compact segments (b2a_batch_compact_*)
COMPACT_HEAD = 64


def encode_compact(fields: dict, ops_lists) -> np.ndarray:
    """Host-side encoder of one rank's compact segment (used by tests to fake a rank's device output)."""
    n = len(ops_lists)
    codes = [np.array([c for c, _ in ops], dtype=np.uint8) for ops in ops_lists]
    total = int(sum(len(c) for c in codes))
    seg = np.zeros(COMPACT_HEAD + 40 * n + total, dtype=np.uint8)
    seg[:16].view(np.uint64)[:] = (n, total)
    a = seg[COMPACT_HEAD:COMPACT_HEAD + 40 * n].view(np.uint32)
    a[0:n] = np.asarray(fields["score"]).astype(np.int32).view(np.uint32)
    for k, name in enumerate(("xstart", "xend", "ystart", "yend")):
        a[(1 + k) * n:(2 + k) * n] = fields[name]
    a[5 * n:6 * n] = [len(c) for c in codes]
    clip = a[6 * n:10 * n].reshape(n, 4)
    for p, ops in enumerate(ops_lists):
        for k, l in enumerate([l for c, l in ops if c >= 4][:4]):
            clip[p, k] = l
    if total:
        seg[COMPACT_HEAD + 40 * n:] = np.concatenate(codes)
    return seg


def decode_compact(host: np.ndarray, segment_bytes: int, world: int):
    """Pure-numpy decoder of `world` gathered segments (the C ABI has b2a_compact_decode for the same job)."""
    fields = {k: [] for k in ("score", "xstart", "xend", "ystart", "yend")}
    ops_lists = []
    for r in range(world):
        seg = host[r * segment_bytes:(r + 1) * segment_bytes]
        n, total = (int(v) for v in seg[:16].view(np.uint64))
        a = seg[COMPACT_HEAD:COMPACT_HEAD + 40 * n].view(np.uint32)
        fields["score"].append(a[0:n].view(np.int32).copy())
        for k, name in enumerate(("xstart", "xend", "ystart", "yend")):
            fields[name].append(a[(1 + k) * n:(2 + k) * n].copy())
        nops = a[5 * n:6 * n]
        clip = a[6 * n:10 * n].reshape(n, 4)
        ops = seg[COMPACT_HEAD + 40 * n:COMPACT_HEAD + 40 * n + total]
        off = 0
        for p in range(n):
            out, k = [], 0
            for c in ops[off:off + int(nops[p])]:
                c = int(c)
                if c >= 4:
                    out.append((c, int(clip[p, k])))
                    k += 1
                else:
                    out.append((c, 0))
            ops_lists.append(out)
            off += int(nops[p])
    return {k: np.concatenate(v) if v else np.zeros(0, np.uint32) for k, v in fields.items()}, ops_lists


def align_sharded_compact(batch, run_local: Callable, device="cpu"):
    """Like align_sharded, with compact segments: `run_local(shard)` -> uint8 torch tensor holding that
    shard's segment (on `device`).  One MAX all-reduce of the segment size, one all-gather of the segments."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    shard, lo, hi = shard_batch(batch, world, rank)
    seg = run_local(shard)
    size = torch.tensor([seg.numel()], dtype=torch.int64, device=device)
    dist.all_reduce(size, op=dist.ReduceOp.MAX)
    seg_bytes = (int(size.item()) + 255) // 256 * 256
    local = torch.zeros(seg_bytes, dtype=torch.uint8, device=device)
    local[:seg.numel()] = seg
    out = torch.empty(world * seg_bytes, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, local)
    return decode_compact(out.cpu().numpy(), seg_bytes, world)
