"""torchrun worker of tests/test_gpu_multi.py: every rank aligns its shard of one batch on its own GPU, the compact
result segments are all-gathered over NCCL, and EVERY rank decodes the gathered segments and checks the whole
batch against the oracle (both wire formats: fixed-capacity segments + b2a_gathered_fetch, and the size-agreed
segments + the host decoder)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as orc  # noqa: E402
from parity_util import assert_same, oracle_batch  # noqa: E402
from rust_bio_b200 import dist as bdist, synth  # noqa: E402
from rust_bio_b200._lib import CScoring, MIN_SCORE  # noqa: E402
from rust_bio_b200.engine import Engine, Results  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for mode_name, mode, n_pairs, gen in (("local", 3, 5000, lambda: synth.uniform_pairs(synth.BASES["C1"], 0, 5000, 150, 150)),
                                          ("global", 1, 1201, lambda: synth.ragged_pairs(9, 1201, 260, 300)),
                                          ("semiglobal", 2, 7, lambda: synth.ragged_pairs(10, 7, 60, 90))):
        batch = gen()
        cs = CScoring(-5, -1, MIN_SCORE, MIN_SCORE, MIN_SCORE, MIN_SCORE, 1, -1, 1, None, None, 0)
        shard, lo, hi = bdist.shard_batch(batch, world, rank)
        eng.stage(mode, cs, shard)
        eng.run()
        # (1) fixed-capacity segments: no size agreement
        cap = 64 + 40 * (hi - lo) + int(shard[2].astype(np.int64).sum() + shard[4].astype(np.int64).sum()) + 4 * (hi - lo)
        capt = torch.tensor([cap], dtype=torch.int64, device="cuda")
        dist.all_reduce(capt, op=dist.ReduceOp.MAX)  # only because shard sizes differ by one pair here
        seg = (int(capt.item()) + 255) // 256 * 256
        local_buf = torch.zeros(seg, dtype=torch.uint8, device="cuda")
        allbuf = torch.zeros(world * seg, dtype=torch.uint8, device="cuda")
        eng.compact_fixed(local_buf.data_ptr(), seg)
        dist.all_gather_into_tensor(allbuf, local_buf)
        torch.cuda.synchronize()
        res = Results(n_pairs, int(Engine.default_ops_capacity(batch)))
        got_n, _ = eng.gathered_fetch(allbuf.data_ptr(), seg, world, res)
        assert got_n == n_pairs, (got_n, n_pairs)
        s, _ = orc.make_scoring(-5, -1, 1, -1)
        ref, ref_ops = oracle_batch(orc, mode_name, s, batch, threads=8)
        assert_same(res.as_dict(), [res.ops_of(i) for i in range(n_pairs)], ref, ref_ops, batch,
                    f"rank {rank}/{world} {mode_name} gathered_fetch")
        # (2) size-agreed compact segments + the pure-host decoders
        nb = torch.tensor([eng.compact_bytes()], dtype=torch.int64, device="cuda")
        dist.all_reduce(nb, op=dist.ReduceOp.MAX)
        seg2 = (int(nb.item()) + 255) // 256 * 256
        l2 = torch.zeros(seg2, dtype=torch.uint8, device="cuda")
        a2 = torch.zeros(world * seg2, dtype=torch.uint8, device="cuda")
        eng.compact_into(l2.data_ptr(), seg2)
        dist.all_gather_into_tensor(a2, l2)
        torch.cuda.synchronize()
        host = a2.cpu().numpy()
        res2 = eng.decode_compact(host, seg2, world, n_pairs, int(Engine.default_ops_capacity(batch)))
        assert_same(res2.as_dict(), [res2.ops_of(i) for i in range(n_pairs)], ref, ref_ops, batch,
                    f"rank {rank}/{world} {mode_name} decode_compact")
        fields, ops_lists = bdist.decode_compact(host, seg2, world)
        assert ops_lists == ref_ops and np.array_equal(fields["score"], ref["score"])
    eng.close()
    # (3) the end-to-end form bench.py times at N > 1: ShardedAligner (pieces pipelined per rank, one all-gather,
    # rank 0 decodes the whole batch); called three times: sizing call, reuse, and a batch that no longer fits the
    # capacity fixed by the first (the decoder reports it and every rank re-sizes)
    sh = bdist.ShardedAligner(local, chunks=3)
    cs = CScoring(-5, -1, MIN_SCORE, MIN_SCORE, MIN_SCORE, MIN_SCORE, 1, -1, 1, None, None, 0)
    s, _ = orc.make_scoring(-5, -1, 1, -1)
    for mode_name, mode, batch in (("local", 3, synth.uniform_pairs(synth.BASES["C1"], 0, 6000, 150, 150)),
                                   ("local", 3, synth.uniform_pairs(synth.BASES["C1"], 6000, 6000, 150, 150)),
                                   ("global", 1, synth.uniform_pairs(synth.BASES["C1"], 0, 6000, 150, 150))):
        shard, lo, hi = bdist.shard_batch(batch, world, rank)
        res = Results(6000, int(Engine.default_ops_capacity(batch))) if rank == 0 else None
        if rank == 0:
            print("ShardedAligner", mode_name, "capacity before", sh.cap, flush=True)
        got = sh.align(mode, cs, shard, res)
        if rank == 0:
            assert got == 6000
            ref, ref_ops = oracle_batch(orc, mode_name, s, batch, threads=8)
            assert_same(res.as_dict(), [res.ops_of(i) for i in range(6000)], ref, ref_ops, batch, f"ShardedAligner {mode_name}")
    # a byte that only the last piece of the last rank holds: the pieces that reused the first piece's alphabet
    # are redone with their own discovery pass
    batch = list(synth.uniform_pairs(synth.BASES["C1"], 0, 6000, 150, 150))
    blob = batch[0].copy()
    blob[int(batch[3][5999]) + 7] = ord("N")
    batch[0] = blob
    batch = tuple(batch)
    shard, lo, hi = bdist.shard_batch(batch, world, rank)
    res = Results(6000, int(Engine.default_ops_capacity(batch))) if rank == 0 else None
    got = sh.align(3, cs, shard, res)
    if rank == 0:
        ref, ref_ops = oracle_batch(orc, "local", s, batch, threads=8)
        assert_same(res.as_dict(), [res.ops_of(i) for i in range(6000)], ref, ref_ops, batch, "ShardedAligner, late new symbol")
    sh.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_SHARD_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
