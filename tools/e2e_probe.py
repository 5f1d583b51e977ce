import os, sys, time
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
for chunks in (5, 4, 0):
    eng.set_pipeline(chunks)
    for it in range(4):
        if it == 3: os.environ["B2A_DEBUG_TIMING"] = "1"
        t0 = time.perf_counter(); eng.align_batch(3, cs, batch, results=res); t1 = time.perf_counter()
        os.environ.pop("B2A_DEBUG_TIMING", None)
    print("chunks", chunks, "e2e ms", round((t1 - t0) * 1e3, 2), "fill", round(eng.stats.fill_ms, 2), "walk", round(eng.stats.walk_ms, 2), flush=True)
# staged pieces
eng.set_pipeline(0)
t0 = time.perf_counter(); eng.stage(3, cs, batch); t1 = time.perf_counter(); eng.run(); t2 = time.perf_counter(); eng.fetch(res); t3 = time.perf_counter()
print("monolithic: stage %.2f run(launch) %.2f fetch(incl. wait) %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
