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


# ---------------------------------------------------------------- one rank of the multi-process form, end to end
class ShardedAligner:
    """Host buffers in -> the whole batch's results in host memory on rank 0, one process per GPU (torch.distributed,
    NCCL).  What `bench.py --gpus N` times as `e2e`.

    Every rank aligns its contiguous share of the pair list: the share is cut into `chunks` pieces, each on its own
    engine and stream, so the H2D copy of piece c+1 runs under the kernels of piece c; every piece leaves a
    fixed-capacity result segment (b2a_batch_compact_fixed: no size agreement); ONE all-gather moves all of them,
    and rank 0 decodes the gathered buffer straight into the caller's host arrays (b2a_gathered_fetch).  The segment
    capacity is sized on the first call (one MAX all-reduce) and reused; a later batch that does not fit is
    reported by the decoder (B2A_E_CAPACITY) and re-sized."""

    WEIGHTS = (1, 2, 2)

    def __init__(self, device: int, chunks: int = 3):
        import torch
        from .engine import Engine
        self.torch = torch
        self.device = device
        w = list(self.WEIGHTS[:chunks]) + [2] * max(0, chunks - len(self.WEIGHTS))
        self.weights = w
        self.engs = [Engine(device) for _ in w]
        self.streams = [torch.cuda.Stream(device=device) for _ in w]
        for e, s in zip(self.engs, self.streams):
            e.set_stream(s.cuda_stream)
            e.set_pipeline(0)
        self.events = [torch.cuda.Event() for _ in w]
        # the exchange has its own (non-default) stream: the engine treats stream handle 0 as "use your own stream",
        # so the decode could otherwise not be ordered after an all-gather issued on the legacy default stream
        self.xstream = torch.cuda.Stream(device=device)
        self._pin = {}
        self.cap = 0
        self.local = self.all = None

    def close(self):
        for e in self.engs:
            e.close()

    def _cuts(self, n):
        tot = float(sum(self.weights))
        cuts, acc = [0], 0.0
        for wv in self.weights:
            acc += wv / tot * n
            cuts.append(min(n, (int(acc) + 31) // 32 * 32))
        cuts[-1] = n
        return cuts

    def _slice(self, c, batch, lo, hi):
        """Piece c of the shard: a view of the blob's span and offsets rebased into PINNED scratch (a copy from
        pageable memory first waits for the stream and then blocks the host)."""
        blob, xo, xl, yo, yl = batch
        if hi <= lo:
            return (blob[:0], xo[lo:lo], xl[lo:lo], yo[lo:lo], yl[lo:lo])
        xo_, yo_, xl_, yl_ = xo[lo:hi], yo[lo:hi], xl[lo:hi], yl[lo:hi]
        bmin = int(min(xo_.min(), yo_.min()))
        bmax = int(max((xo_ + xl_.astype(np.uint64)).max(), (yo_ + yl_.astype(np.uint64)).max()))
        n = hi - lo
        if c not in self._pin or self._pin[c][0].numel() < n:
            self._pin[c] = (self.torch.empty(n + n // 8 + 16, dtype=self.torch.int64).pin_memory(),
                            self.torch.empty(n + n // 8 + 16, dtype=self.torch.int64).pin_memory())
        px = self._pin[c][0].numpy().view(np.uint64)[:n]
        py = self._pin[c][1].numpy().view(np.uint64)[:n]
        np.subtract(xo_, np.uint64(bmin), out=px)
        np.subtract(yo_, np.uint64(bmin), out=py)
        return (blob[bmin:bmax], px, xl_, py, yl_)

    def align(self, mode: int, cscoring, shard, results=None):
        """`shard` = this rank's share (dist.shard_batch); `results` = rust_bio_b200.engine.Results for the WHOLE
        batch on rank 0 (None elsewhere).  Returns the number of pairs decoded on rank 0, else 0."""
        import torch.distributed as dist
        torch = self.torch
        world, rank = dist.get_world_size(), dist.get_rank()
        n = len(shard[2])
        cuts = self._cuts(n)
        k = len(self.engs)
        for attempt in range(2):
            sizing = self.cap == 0
            for reuse in (True, False):
                cs_piece, alpha_keep = cscoring, None
                for c, (e, s) in enumerate(zip(self.engs, self.streams)):
                    sub = self._slice(c, shard, cuts[c], cuts[c + 1])
                    e.stage(mode, cs_piece, sub)   # H2D of the piece (returns when the copies have landed) ...
                    e.run()                        # ... its kernels run under the next piece's copies
                    if reuse and c == 0 and not cscoring.alphabet and not cscoring.table and cuts[1] > cuts[0]:
                        # later pieces reuse the alphabet the first one found: their stage then has no discovery
                        # pass (a kernel + a sync that would wait behind this piece's persistent fill)
                        import ctypes as _C
                        from ._lib import CScoring as _CS
                        alpha_keep = e.last_alphabet()
                        cs_piece = _CS.from_buffer_copy(bytes(cscoring))
                        cs_piece.alphabet = alpha_keep.ctypes.data_as(_C.c_void_p)
                        cs_piece.alphabet_len = len(alpha_keep)
                try:
                    for e in self.engs:
                        e.fetch(None)              # waits for the piece; reports a byte outside the alphabet
                    break
                except Exception as ex:            # a later piece held a byte the first did not: discover per piece
                    if not reuse or "alphabet" not in str(ex):
                        raise
            if sizing:
                need = max(e.compact_bytes() for e in self.engs)
                t = torch.tensor([need], dtype=torch.int64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                self.cap = (int(int(t.item()) * 1.15) + (1 << 16) - 1) >> 16 << 16
                self.local = torch.empty(k * self.cap, dtype=torch.uint8, device="cuda")
                self.all = torch.empty(world * k * self.cap, dtype=torch.uint8, device="cuda")
            cur = self.xstream
            for c, (e, s, ev) in enumerate(zip(self.engs, self.streams, self.events)):
                e.compact_fixed(self.local.data_ptr() + c * self.cap, self.cap)
                ev.record(s)
                cur.wait_event(ev)
            with torch.cuda.stream(cur):
                dist.all_gather_into_tensor(self.all, self.local)
            got = 0
            ok = torch.ones(1, dtype=torch.int32, device="cuda")
            if rank == 0:
                try:
                    self.engs[0].set_stream(cur.cuda_stream)
                    got, self.d2h = self.engs[0].gathered_fetch(self.all.data_ptr(), self.cap, world * k, results)
                except Exception as ex:  # a segment did not fit the capacity fixed earlier: size again
                    if "CAPACITY" not in str(ex) or attempt:
                        raise
                    ok[0] = 0
                finally:
                    self.engs[0].set_stream(self.streams[0].cuda_stream)
            else:
                cur.synchronize()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()):
                return got
            self.cap = 0  # every rank redoes the batch with a fresh sizing
        return 0
