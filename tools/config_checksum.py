"""Checksum (wrapping i64 sum of scores) of a BASELINE config's first N pairs as the ORACLE computes them: the
number rust_bio_b200/rust/examples/configs_bench.rs prints for rust-bio itself must equal it.
  python tools/config_checksum.py C2 20000"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from rust_bio_b200 import scores, synth

name, n = sys.argv[1], int(sys.argv[2])
orc.build()
if name in ("C1", "C2"):
    b = synth.uniform_pairs(synth.BASES[name], 0, n, 150, 150)
    s, _ = orc.make_scoring(-5, -1, 1, -1)
    ref, *_ = orc.align_batch("local", s, *b, threads=orc.hardware_threads(), want_ops=False)
elif name == "C3":
    b = synth.uniform_pairs(synth.BASES[name], 0, n, 1000, 1000)
    s, _ = orc.make_scoring(-5, -1, 1, -1)
    ref, *_ = orc.align_batch("global", s, *b, threads=orc.hardware_threads(), want_ops=False)
elif name == "C5":
    b = synth.uniform_pairs(synth.BASES[name], 0, n, 10000, 10000, alphabet=synth.PROTEIN)
    s, keep = orc.make_scoring(-10, -1, 0, 0, scores.matrix_table256("blosum62"))
    ref, *_ = orc.align_batch("local", s, *b, threads=orc.hardware_threads(), want_ops=False)
else:
    b = synth.mutated_window_pairs(synth.BASES["C4"], 0, n, 500, 10000)
    s, _ = orc.make_scoring(-5, -1, 1, -1, has_match_scores=1)
    ref, *_ = orc.banded_align_batch("semiglobal", s, 32, 32, *b, threads=orc.hardware_threads(), want_ops=False)
print({"config": name, "pairs": n, "checksum": int(ref["score"].astype(np.int64).sum())})
