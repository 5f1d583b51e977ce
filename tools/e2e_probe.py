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
WEIGHTS = os.environ.get("PROBE_WEIGHTS", "1,2,2,2,1;1,3,6,6;1,3,4;1,2,4,4;1,1,2,4;2,3,3;1,2,3,3,3").split(";")
for wts in WEIGHTS:
    os.environ["B2A_PIPE_WEIGHTS"] = wts
    eng.set_pipeline(5)
    ts = []
    for it in range(14):
        t0 = time.perf_counter(); eng.align_batch(3, cs, batch, results=res); t1 = time.perf_counter()
        ts.append((t1 - t0) * 1e3)
    print("weights", wts, "e2e ms", [round(t, 1) for t in ts[4:]], "median", round(sorted(ts[4:])[5], 2), "fill", round(eng.stats.fill_ms, 2), "walk", round(eng.stats.walk_ms, 2), flush=True)
os.environ.pop("B2A_PIPE_WEIGHTS", None)
# staged pieces
eng.set_pipeline(0)
t0 = time.perf_counter(); eng.stage(3, cs, batch); t1 = time.perf_counter(); eng.run(); t2 = time.perf_counter(); eng.fetch(res); t3 = time.perf_counter()
print("monolithic: stage %.2f run(launch) %.2f fetch(incl. wait) %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
