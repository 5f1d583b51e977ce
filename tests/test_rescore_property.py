"""The fuzz target's self-consistency property (fuzz/fuzz_targets/banded_aligner.rs:10-56) as a test helper:
checked here on the ORACLE's results (not-gpu), so the GPU tests can rely on the helper itself."""
import numpy as np
import pytest

from parity_util import oracle_batch, rescore_path
from rust_bio_b200 import synth

MIN = -858993459


@pytest.mark.parametrize("mode", ["global", "semiglobal", "local", "custom"])
def test_oracle_paths_rescore_to_their_score(oracle, mode):
    rng = np.random.default_rng(11)
    for trial in range(6):
        go, ge = int(rng.choice([0, -1, -2, -5, -6])), int(rng.choice([0, -1, -1, -2]))
        ge = max(ge, go)  # the affine model the property is stated for: opening costs at least as much as extending
        ma, mi = int(rng.choice([1, 2, 4])), int(rng.choice([-1, -3, -7]))
        clips = (MIN, MIN, MIN, MIN)
        if mode == "custom":
            clips = tuple(int(rng.choice([MIN, 0, -1, -3, -7, -20])) for _ in range(4))
        s, _ = oracle.make_scoring(go, ge, ma, mi, None, *clips)
        batch = synth.ragged_pairs(trial, 120, 70, 90, alphabet=b"AC" if trial % 2 else b"ACGT")
        ref, ref_ops = oracle_batch(oracle, mode, s, batch)
        blob, xo, xl, yo, yl = batch
        eff = {"global": (MIN,) * 4, "semiglobal": (MIN, MIN, 0, 0), "local": (0, 0, 0, 0)}.get(mode, clips)
        for p in range(len(xl)):
            x = bytes(blob[int(xo[p]):int(xo[p]) + int(xl[p])])
            y = bytes(blob[int(yo[p]):int(yo[p]) + int(yl[p])])
            f = {k: ref[k][p] for k in ("xstart", "xend", "ystart", "yend")}
            got = rescore_path(x, y, ref_ops[p], f, mode, go, ge, lambda a, b: ma if a == b else mi, eff)
            assert got == int(ref["score"][p]), (mode, trial, p, clips, x, y, ref_ops[p])
