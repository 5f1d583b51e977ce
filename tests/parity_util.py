"""Shared comparison helpers: an engine-like result (dict of arrays + op lists) vs the oracle."""
import numpy as np

MODES = {"custom": 0, "global": 1, "semiglobal": 2, "local": 3}


def oracle_batch(orc, mode, scoring, batch, threads=4):
    blob, xo, xl, yo, yl = batch
    out, ops, off, _ = orc.align_batch(mode, scoring, blob, xo, xl, yo, yl, threads=threads)
    lists = []
    for p in range(len(xl)):
        seg = ops[int(off[p]):int(off[p]) + int(out["n_ops"][p])]
        lists.append([(int(v) & 7, int(v) >> 3) for v in seg])
    return out, lists


def assert_same(got, got_ops, ref, ref_ops, batch, what=""):
    blob, xo, xl, yo, yl = batch
    for f in ("score", "xstart", "xend", "ystart", "yend"):
        bad = np.nonzero(np.asarray(got[f]).astype(np.int64) != ref[f].astype(np.int64))[0]
        if len(bad):
            p = int(bad[0])
            x = bytes(blob[int(xo[p]):int(xo[p]) + int(xl[p])])
            y = bytes(blob[int(yo[p]):int(yo[p]) + int(yl[p])])
            raise AssertionError(f"{what}: field {f} differs for {len(bad)} pairs; first pair {p}: "
                                 f"got {got[f][p]} ref {ref[f][p]} x={x} y={y}")
    for p in range(len(xl)):
        if got_ops[p] != ref_ops[p]:
            x = bytes(blob[int(xo[p]):int(xo[p]) + int(xl[p])])
            y = bytes(blob[int(yo[p]):int(yo[p]) + int(yl[p])])
            raise AssertionError(f"{what}: ops differ for pair {p}: got {got_ops[p]} ref {ref_ops[p]} x={x} y={y}")
    if "status" in got:
        assert not np.any(got["status"]), what


def rescore_path(x, y, ops, fields, mode, go, ge, score_fn, clips=(0, 0, 0, 0)):
    """Independent of the oracle: re-score a returned path with the 4.0 gap model (gap_open for the first
    symbol of a gap, gap_extend for each further one; reference mod.rs:9-15) and the clip penalties, and check
    it walks from (xstart, ystart) to (xend, yend).  Structure of fuzz/fuzz_targets/banded_aligner.rs:10-56
    (whose first-gap cost go+ge is the pre-4.0 convention).  Returns the path's score.
    ops: [(code, len)] in alignment order (clips present only in custom mode).
    When gap_open > gap_extend (opening cheaper than extending) the recurrence re-opens instead of extending
    (D = max(D + ge, S + go) with S == D), so a further gap symbol costs max(gap_extend, gap_open)."""
    xp, xs, yp, ys = clips
    ge = max(ge, go)
    m, n = len(x), len(y)
    i, j = int(fields["xstart"]), int(fields["ystart"])
    score, last = 0, None
    if mode == "custom":  # fuzz target: a clip is charged when the alignment leaves that end unaligned
        if i > 0:
            score += xp
        if j > 0:
            score += yp
        if int(fields["xend"]) < m:
            score += xs
        if int(fields["yend"]) < n:
            score += ys
    for c, _ in ops:
        if c in (0, 1):
            assert (x[i] == y[j]) == (c == 0), "Match/Subst must follow byte equality (mod.rs:762)"
            score += score_fn(x[i], y[j])
            i += 1
            j += 1
        elif c == 2:
            score += ge if last == 2 else go
            j += 1
        elif c == 3:
            score += ge if last == 3 else go
            i += 1
        else:
            continue  # clips move nothing inside [start, end)
        last = c
    assert (i, j) == (int(fields["xend"]), int(fields["yend"])), "ops do not span [start, end)"
    if mode == "global":
        assert (int(fields["xstart"]), int(fields["ystart"]), i, j) == (0, 0, m, n)
    if mode == "semiglobal":
        assert (int(fields["xstart"]), i) == (0, m)
    return score
