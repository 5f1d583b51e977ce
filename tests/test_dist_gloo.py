"""N>1 host logic on CPU: world_size-2 gloo run of the sharding + single all-gather + reassembly
(rust_bio_b200/dist.py).  Each rank's 'device output' is faked from the oracle with the same
fixed-stride record layout the CUDA records kernel writes."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch, torch.distributed as dist
    from rust_bio_b200 import synth, dist as bdist
    from oracle import oracle as orc
    from parity_util import oracle_batch, assert_same
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    batch = synth.ragged_pairs(21, 37, 40, 50)          # 37 pairs: shards of 19 and 18
    s, _ = orc.make_scoring(-5, -1, 1, -1)
    stride = bdist.record_stride(int(batch[2].max()), int(batch[4].max()))
    calls = []
    def run_local(shard):
        calls.append(len(shard[2]))
        ref, ops = oracle_batch(orc, "local", s, shard, threads=1)
        return torch.from_numpy(bdist.encode_records(ref, ops, stride))
    fields, ops = bdist.align_sharded(batch, stride, run_local)
    ref, ref_ops = oracle_batch(orc, "local", s, batch, threads=1)
    assert_same(fields, ops, ref, ref_ops, batch, "gloo rank %d" % rank)
    # the compact wire format (what bench.py exchanges): same answer
    def run_local_compact(shard):
        ref, ops = oracle_batch(orc, "local", s, shard, threads=1)
        return torch.from_numpy(bdist.encode_compact(ref, ops))
    fields2, ops2 = bdist.align_sharded_compact(batch, run_local_compact)
    assert_same(fields2, ops2, ref, ref_ops, batch, "gloo compact rank %d" % rank)
    lo, hi = bdist.shard_range(37, world, rank)
    assert calls == [hi - lo] and (lo, hi) == ((0, 19) if rank == 0 else (19, 37))
    dist.barrier(); dist.destroy_process_group()
    print("rank %d ok" % rank, flush=True)  # one string: the two ranks share the pipe
''')


def test_world2_gloo_shard_allgather_reassemble(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_shard_ranges_cover_exactly():
    from rust_bio_b200 import dist as bdist
    for n in (0, 1, 7, 8, 1000, 1_000_001):
        for world in (1, 2, 4, 8):
            spans = [bdist.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_record_codec_roundtrip():
    from rust_bio_b200 import dist as bdist
    fields = {"score": np.array([5, -7], dtype=np.int32), "xstart": np.array([1, 0], dtype=np.uint32),
              "xend": np.array([4, 3], dtype=np.uint32), "ystart": np.array([0, 2], dtype=np.uint32),
              "yend": np.array([3, 9], dtype=np.uint32)}
    ops = [[(4, 1), (0, 0), (1, 0), (0, 0), (5, 6)], [(2, 0), (3, 0), (0, 0)]]
    stride = bdist.record_stride(10, 12)
    rec = bdist.encode_records(fields, ops, stride)
    f2, o2 = bdist.decode_records(rec, stride, 2)
    assert o2 == ops and all(np.array_equal(f2[k], fields[k]) for k in fields)
