"""N > 1 on real GPUs: b2a_batch_compact_* -> NCCL all-gather -> decode across ranks, checked against the
oracle on every rank (the N > 1 results had only ever been verified with a single-rank round trip).
Needs >= 2 GPUs on the box (`gpurun --gpus 2`); skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_all_gather_of_result_segments_matches_the_oracle(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "workers", "nccl_shard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"NCCL_SHARD_OK world={world}" in r.stdout
