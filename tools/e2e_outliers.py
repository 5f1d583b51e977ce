"""Dev tool: find what a slow b2a_align_batch call spends its time on (per-chunk host timeline on stderr)."""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rust_bio_b200 import synth
from rust_bio_b200._lib import CScoring, MIN_SCORE
from rust_bio_b200.engine import Engine, Results
P = 1_000_000
batch = synth.uniform_pairs(synth.BASES["C2"], 0, P, 150, 150)
keep = [torch.from_numpy(a).pin_memory() for a in batch]
batch = tuple(k.numpy() for k in keep)
eng = Engine(0)
cs = CScoring(-5, -1, MIN_SCORE, MIN_SCORE, MIN_SCORE, MIN_SCORE, 1, -1, 1, None, None, 0)
outs = {k: torch.empty(n, dtype=dt).pin_memory() for k, n, dt in (("score", P, torch.int32), ("xstart", P, torch.int32), ("xend", P, torch.int32), ("ystart", P, torch.int32), ("yend", P, torch.int32), ("ops_off", P + 1, torch.int64), ("ops", 64 * P, torch.uint8), ("clip_len", 4 * P, torch.int32))}
views = {k: (v.numpy().view(np.uint32) if k in ("xstart", "xend", "ystart", "yend", "clip_len") else v.numpy().view(np.uint64) if k == "ops_off" else v.numpy()) for k, v in outs.items()}
res = Results(P, 64 * P, out=views)
os.environ["B2A_DEBUG_TIMING"] = "1"
ts = []
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 80):
    sys.stderr.write("ITER %d\n" % it); sys.stderr.flush()
    t0 = time.perf_counter(); eng.align_batch(3, cs, batch, results=res); t1 = time.perf_counter()
    ts.append((t1 - t0) * 1e3)
    sys.stderr.write("ITER %d took %.2f ms\n" % (it, ts[-1])); sys.stderr.flush()
print("e2e ms:", [round(t, 1) for t in ts])
print("median %.2f  mean %.2f  max %.2f" % (sorted(ts)[len(ts) // 2], sum(ts) / len(ts), max(ts)))
