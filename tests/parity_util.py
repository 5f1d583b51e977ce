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
