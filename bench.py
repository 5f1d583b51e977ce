#!/usr/bin/env python
"""bench.py -- GCUPS of the batched pairwise-alignment hot path on N B200s (BASELINE.json metric).

A step = one pass of the hot path (K0 pack -> K1 fill -> K2 row-m/fix-ups/walk -> ops compaction,
plus the single NCCL all-gather of result records when N > 1) over one batch of synthetic pairs.
Workload at every N: BASELINE config 2 -- 1M pairs of 150x150 uniform random DNA per GPU, local
affine (match 1, mismatch -1, gap_open -5, gap_extend -1); weak scaling (pairs split by rank).

  value  : whole-job GCUPS with the batch already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e    : same metric through b2a_align_batch with pinned HOST buffers (H2D + kernels + D2H inside)
  roofline / int32_alu : the K1 fill kernel against measured HBM bandwidth and the measured int32 ALU peak
  cpu_baseline : the oracle (C++ restatement of rust-bio 4.0.1) on the box's host cores, bounded sample
  --impl reference : the reference's CPU implementation of the path (the oracle port: rust-bio itself
                     cannot be compiled in this image), all host threads, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

M, N_LEN = 150, 150
SCORING = dict(gap_open=-5, gap_extend=-1, match=1, mismatch=-1)
WORKLOAD = "C2: 1M pairs/GPU of 150x150 uniform random DNA, Aligner::local, match 1 mismatch -1 gap_open -5 gap_extend -1"
OPS_PER_CELL_LOCAL = 25  # SURVEY 8d accounting convention


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(",") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                r = [c.strip() for c in r]
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                                  ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            busy = sorted(sm)[len(sm) // 2:]  # upper half: samples taken under load
            out["sm_mhz"] = float(np.median(busy))
            out["sm_max_mhz"] = mx
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def oracle_scoring(orc):
    s, _ = orc.make_scoring(SCORING["gap_open"], SCORING["gap_extend"], SCORING["match"], SCORING["mismatch"])
    return s


def cpu_sample(orc, threads: int, target_s: float):
    """Time the oracle on a bounded sample of the workload; returns (gcups, n_pairs, seconds)."""
    from rust_bio_b200 import synth
    s = oracle_scoring(orc)
    probe = synth.uniform_pairs(synth.BASES["C2"], 0, 64 * threads, M, N_LEN)
    _, _, _, t = orc.align_batch("local", s, *probe, threads=threads, want_ops=False)
    rate = 64 * threads / max(t, 1e-6)
    n = int(max(threads * 64, min(400_000, rate * target_s)))
    batch = synth.uniform_pairs(synth.BASES["C2"], 0, n, M, N_LEN)
    _, _, _, t = orc.align_batch("local", s, *batch, threads=threads, want_ops=False)
    return n * M * N_LEN / t / 1e9, n, t


def run_reference(args):
    """The reference arm: rust-bio's CPU path (oracle port) on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = orc.hardware_threads()
    vals, sample = [], None
    for it in range(args.warmup + args.steps):
        g, n, t = cpu_sample(orc, threads, target_s=4.0 if it >= args.warmup else 1.0)
        if it >= args.warmup:
            vals.append((g, n, t))
            sample = n
    cells = sum(v[1] for v in vals) * M * N_LEN
    secs = sum(v[2] for v in vals)
    value = cells / secs / 1e9
    line = {
        "impl": "reference", "metric": "GCUPS", "value": round(value, 4), "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(secs / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "rust-bio cannot be built here (no rustc); this is the C++ "
                   "restatement of rust-bio 4.0.1 pinned to the reference's known-answer vectors"},
        "cpu_baseline": {"value": round(value, 4), "unit": "GCUPS", "cores": threads, "kind": "port",
                         "sample": f"{sample} pairs of {M}x{N_LEN} per step, {threads} threads"},
        "e2e": {"value": round(value, 4), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from rust_bio_b200 import synth
    from rust_bio_b200._lib import CScoring, MIN_SCORE, load
    from rust_bio_b200.engine import Engine, Results

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = args.pairs
    cells_rank = P * M * N_LEN
    steps, warm = args.steps, max(args.warmup, 0)

    # ---- synthetic inputs of the named shape, in PINNED host memory (this rank's shard of the pair list)
    blob, x_off, x_len, y_off, y_len = synth.uniform_pairs(synth.BASES["C2"], rank * P, P, M, N_LEN)

    def pin(a):
        t = torch.from_numpy(a).pin_memory()
        return t, t.numpy()

    keep = [pin(a) for a in (blob, x_off, x_len, y_off, y_len)]
    batch = tuple(k[1] for k in keep)
    ops_cap = 64 * P  # local alignments of random DNA are ~15 ops; capacity is checked by the ABI
    out_t = {"score": torch.empty(P, dtype=torch.int32).pin_memory(),
             "xstart": torch.empty(P, dtype=torch.int32).pin_memory(),
             "xend": torch.empty(P, dtype=torch.int32).pin_memory(),
             "ystart": torch.empty(P, dtype=torch.int32).pin_memory(),
             "yend": torch.empty(P, dtype=torch.int32).pin_memory(),
             "ops_off": torch.empty(P + 1, dtype=torch.int64).pin_memory(),
             "ops": torch.empty(ops_cap, dtype=torch.uint8).pin_memory(),
             "clip_len": torch.empty(4 * P, dtype=torch.int32).pin_memory()}
    views = {"score": out_t["score"].numpy(), "xstart": out_t["xstart"].numpy().view(np.uint32),
             "xend": out_t["xend"].numpy().view(np.uint32), "ystart": out_t["ystart"].numpy().view(np.uint32),
             "yend": out_t["yend"].numpy().view(np.uint32), "ops_off": out_t["ops_off"].numpy().view(np.uint64),
             "ops": out_t["ops"].numpy(), "clip_len": out_t["clip_len"].numpy().view(np.uint32)}
    results = Results(P, ops_cap, out=views)

    eng = Engine(local)
    # a real (non-legacy-default) torch stream: the engine launches on it, torch.cuda.Event times it,
    # and NCCL collectives issued under torch.cuda.stream(stream) are ordered on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    cs = CScoring(SCORING["gap_open"], SCORING["gap_extend"], MIN_SCORE, MIN_SCORE, MIN_SCORE, MIN_SCORE,
                  SCORING["match"], SCORING["mismatch"], 1, None, None, 0)
    MODE_LOCAL = 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: stage once, then time K passes of the hot path
    eng.stage(MODE_LOCAL, cs, batch)
    # N > 1: the one exchange of the path is the all-gather that reassembles the per-pair results on every
    # rank.  Ranks exchange compact segments (include/b200align.h: 40 B of fields per pair + the ops at
    # their real length); a MAX all-reduce of one integer makes the segments equally long first.
    seg_size = torch.zeros(1, dtype=torch.int64, device="cuda") if world > 1 else None
    seg_buf = {"cap": 0, "local": None, "all": None}
    gathered = {"bytes": 0}

    def step_resident():
        eng.run()
        if world > 1:
            seg_size[0] = eng.compact_bytes()  # waits for this rank's batch
            dist.all_reduce(seg_size, op=dist.ReduceOp.MAX)
            seg = (int(seg_size.item()) + (1 << 20) - 1) >> 20 << 20
            if seg > seg_buf["cap"]:
                seg_buf["cap"] = seg
                seg_buf["local"] = torch.empty(seg, dtype=torch.uint8, device="cuda")
                seg_buf["all"] = torch.empty(world * seg, dtype=torch.uint8, device="cuda")
            eng.compact_into(seg_buf["local"].data_ptr(), seg)
            dist.all_gather_into_tensor(seg_buf["all"][:world * seg], seg_buf["local"][:seg])
            gathered["bytes"] = world * seg

    for _ in range(warm):
        step_resident()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step_resident()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    ms_step = max_over_ranks(ms_total / steps)
    eng.fetch(None)
    st = eng.stats
    launches_step = int(st.kernel_launches)  # the exchange adds D2D copies and NCCL's kernels, none of ours
    # kernel-level numbers over instrumented passes (engine CUDA events on the same stream)
    fills, walks, packs = [], [], []
    for _ in range(max(3, steps)):
        eng.run()
        eng.fetch(None)
        fills.append(eng.stats.fill_ms)
        walks.append(eng.stats.walk_ms)
        packs.append(eng.stats.pack_ms)
    fill_ms = float(np.mean(fills))
    value = world * cells_rank / (ms_step * 1e-3) / 1e9

    # ---- end-to-end arm: the public batch call with pinned host buffers, copies inside the timed region
    for _ in range(max(warm, 3)):
        eng.align_batch(MODE_LOCAL, cs, batch, results=results)
    barrier()
    t0 = time.perf_counter()
    e2e_each = []
    for _ in range(steps):
        t1 = time.perf_counter()
        eng.align_batch(MODE_LOCAL, cs, batch, results=results)  # returns with the results in host memory
        e2e_each.append(round((time.perf_counter() - t1) * 1e3, 2))
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / steps * 1e3)
    h2d, d2h = int(eng.stats.h2d_bytes), int(eng.stats.d2h_bytes)
    e2e_value = world * cells_rank / (e2e_ms * 1e-3) / 1e9
    G, R = int(eng.stats.fill_lanes_per_pair), int(eng.stats.fill_rows_per_lane)
    tb_bytes = int(eng.stats.traceback_bytes)

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        # algorithmic bytes of one K1 launch (DESIGN.md "K1 bytes"): staged sequences in, 4-bit traceback
        # out, strip-boundary rows out+in, row trackers and last column out
        nstrips = -(-(M - 1) // (G * R))
        per_pair = (M + N_LEN) + 16 * N_LEN * (2 * nstrips - 1) + 20 * (M - 1)
        fill_bytes = P * per_pair + tb_bytes
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "fill_traffic.json")
        if os.path.exists(tpath) and (G, R) == (1, 16):
            with open(tpath) as f:
                traffic = int(json.load(f)["dram_bytes_per_pair"] * P)  # ncu dram read+write per pair x pairs of this launch
        roof = {"bound": "hbm", "kernel": f"fill_kernel<G={G},R={R},local>",
                "achieved": round(fill_bytes / (fill_ms * 1e-3) / 1e9, 2), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(fill_bytes / (fill_ms * 1e-3) / 1e9 / hbm_peak, 4), "traffic": traffic,
                "traffic_source": "profiles/fill_traffic.json (ncu --set full capture at 200k pairs, scaled per pair)" if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(fill_bytes),
                "kernel_ms": round(fill_ms, 4),
                "note": "integer DP: the binding roof is int32 ALU issue (see int32_alu), not HBM"}
        L = load()
        import ctypes as C
        fa, fb, fc = C.c_float(), C.c_float(), C.c_float()
        L.b2a_util_int32_peak(local, C.byref(fa), C.byref(fb), C.byref(fc))
        p_int = max(fa.value, fb.value, fc.value)
        fill_gcups = cells_rank / (fill_ms * 1e-3) / 1e9
        alu = {"ops_per_cell": OPS_PER_CELL_LOCAL, "fill_gcups": round(fill_gcups, 2),
               "peak_tera_lane_ops": {"add": round(fa.value, 2), "minmax": round(fb.value, 2),
                                      "add_max_mix": round(fc.value, 2)},
               "frac": round(fill_gcups * OPS_PER_CELL_LOCAL / 1e3 / p_int, 4),
               "frac_whole_step": round(value / world * OPS_PER_CELL_LOCAL / 1e3 / p_int, 4),
               "peak_source": "measured in this run: b2a_util_int32_peak (independent register chains, all SMs)"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as orc
            orc.build()
            threads = orc.hardware_threads()
            g, n, t = cpu_sample(orc, threads, target_s=12.0)
            g1, n1, t1 = cpu_sample(orc, 1, target_s=4.0)
            cpu = {"value": round(g, 4), "unit": "GCUPS", "cores": threads, "kind": "port",
                   "sample": f"{n} pairs of {M}x{N_LEN} ({t:.1f} s, {threads} threads); 1 thread: {g1:.4f} GCUPS on {n1} pairs",
                   "single_thread_value": round(g1, 4)}
        line = {
            "metric": "GCUPS", "value": round(value, 2), "unit": "GCUPS", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": P, "m": M, "n": N_LEN,
                       "fill_shape": {"lanes_per_pair": G, "rows_per_lane": R},
                       "l2": "inputs larger than L2: 320 MB sequence blob + ~12 GB traceback stream per step (126 MB L2)",
                       "parallelism": (f"pair list sharded over {world} GPU(s); per step one MAX all-reduce of the segment size and "
                                       f"one NCCL all-gather of compact result segments ({gathered['bytes']} B received per rank)")
                       if world > 1 else "single GPU"},
            # value = work / wall time of the K calls (contract); the median call is listed beside it because a
            # call now and then stalls 10-30 ms in a host-side CUDA API call on these shared boxes
            # (tools/e2e_outliers.py: the GPU is idle during those stalls)
            "e2e": {"value": round(e2e_value, 2), "unit": "GCUPS", "ms_per_step": round(e2e_ms, 3),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_each_step": e2e_each,
                    "ms_median_step": float(np.median(e2e_each)),
                    "value_at_median_step": round(world * cells_rank / (float(np.median(e2e_each)) * 1e-3) / 1e9, 2)},
            "gpu_launches": launches_step * steps,
            "kernel_ms": {"pack": round(float(np.mean(packs)), 4), "fill": round(fill_ms, 4),
                          "walk_and_compact": round(float(np.mean(walks)), 4)},
            "roofline": roof, "int32_alu": alu, "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
