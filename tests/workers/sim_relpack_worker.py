"""Runs in its own process with B2A_SIM_KREL_BITS set (a host build of the kernel sources whose relative packed
trackers use SHORT column chunks): the long-sequence tracker form (F_PACKREL: row-tracker columns relative to a chunk
and flushed to the rows arena per chunk, column-tracker rows relative to the strip and made absolute where the strip
hands the boundary on) against the oracle, on every fill shape the host emulation runs."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np

import sim_util
from oracle import oracle
from parity_util import MODES, assert_same, oracle_batch
from rust_bio_b200 import scores, synth

assert os.environ.get("B2A_SIM_KREL_BITS"), "run with B2A_SIM_KREL_BITS set"
oracle.build()
MIN = -858993459
rng = np.random.default_rng(2024)
n_cases = 0
for G, R in ((1, 16), (1, 4), (8, 8), (4, 16), (32, 8), (132, 8)):
    for trial in range(4):
        mode = ["local", "semiglobal", "custom", "custom"][trial]
        clips = (MIN,) * 4
        if mode == "custom":
            clips = tuple(int(rng.choice([MIN, 0, -2, -7])) for _ in range(4))
        s, _ = oracle.make_scoring(int(rng.choice([-5, -1, 0])), int(rng.choice([-1, -2, 0])), 2, -2, None, *clips)
        span = (32 if G == 132 else G) * R
        if trial % 2:
            batch = synth.uniform_pairs(900 + trial, 0, 6, 2 * span + 7, 150 + 13 * trial)  # uniform: the unmasked strips
        else:
            batch = synth.ragged_pairs(700 + 10 * G + trial, 9, 1, 2 * span + 37, min_len=1)
            # long y: several column chunks per strip
            batch = synth.ragged_pairs(700 + 10 * G + trial, 9, 2 * span + 37, 170, min_len=3)
        ref, ref_ops = oracle_batch(oracle, mode, s, batch, threads=4)
        got, ops = sim_util.align_batch(MODES[mode], s, *batch, R=R, G=G, rel_pack=1, force_general=trial == 3)
        assert_same(got, ops, ref, ref_ops, batch, f"relative packed trackers {G}x{R} {mode} trial {trial}")
        n_cases += 1
# BLOSUM62 (LUT) local, the shape class of BASELINE config 5
table = scores.matrix_table256("blosum62")
s, keep = oracle.make_scoring(-10, -1, 0, 0, table)
batch = synth.ragged_pairs(31, 6, 300, 200, alphabet=synth.PROTEIN, min_len=40)
ref, ref_ops = oracle_batch(oracle, "local", s, batch, threads=4)
for G, R in ((132, 8), (8, 8)):
    got, ops = sim_util.align_batch(MODES["local"], s, *batch, R=R, G=G, rel_pack=1)
    assert_same(got, ops, ref, ref_ops, batch, f"relative packed trackers blosum62 {G}x{R}")
    n_cases += 1
print("OK", n_cases)
