"""ctypes wrapper around oracle/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
MIN_SCORE = -858993459
MODES = {"custom": 0, "global": 1, "semiglobal": 2, "local": 3}


class OrcScoring(C.Structure):
    _fields_ = [("gap_open", C.c_int32), ("gap_extend", C.c_int32),
                ("xclip_prefix", C.c_int32), ("xclip_suffix", C.c_int32),
                ("yclip_prefix", C.c_int32), ("yclip_suffix", C.c_int32),
                ("match_score", C.c_int32), ("mismatch_score", C.c_int32),
                ("has_match_scores", C.c_int32),
                ("table", C.POINTER(C.c_int32)),
                ("alphabet", C.c_void_p), ("alphabet_len", C.c_uint32)]


class OrcAlignment(C.Structure):
    _fields_ = [("score", C.c_int32), ("ystart", C.c_uint32), ("xstart", C.c_uint32),
                ("yend", C.c_uint32), ("xend", C.c_uint32), ("ylen", C.c_uint32),
                ("xlen", C.c_uint32), ("mode", C.c_uint32), ("n_ops", C.c_uint32)]


ALN_DTYPE = np.dtype([("score", "<i4"), ("ystart", "<u4"), ("xstart", "<u4"), ("yend", "<u4"),
                      ("xend", "<u4"), ("ylen", "<u4"), ("xlen", "<u4"), ("mode", "<u4"),
                      ("n_ops", "<u4")])


def build(force: bool = False) -> str:
    """Compile the oracle if needed (g++ only; no reference sources are involved)."""
    srcs = [os.path.join(_HERE, f) for f in ("pairwise_oracle.cpp", "banded_oracle.cpp")]
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_align.restype = C.c_int
        L.orc_align_batch.restype = C.c_double
        L.orc_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def make_scoring(gap_open, gap_extend, match=0, mismatch=0, table=None, xclip_prefix=MIN_SCORE,
                 xclip_suffix=MIN_SCORE, yclip_prefix=MIN_SCORE, yclip_suffix=MIN_SCORE,
                 has_match_scores=0):
    s = OrcScoring(gap_open, gap_extend, xclip_prefix, xclip_suffix, yclip_prefix, yclip_suffix,
                   match, mismatch, has_match_scores, None, None, 0)
    keep = None
    if table is not None:
        keep = np.ascontiguousarray(table, dtype=np.int32).reshape(256 * 256)
        s.table = keep.ctypes.data_as(C.POINTER(C.c_int32))
    return s, keep


def align(mode, scoring: OrcScoring, x: bytes, y: bytes):
    """One pair -> (dict of Alignment fields, [(code, len), ...])."""
    mode = MODES.get(mode, mode)
    m, n = len(x), len(y)
    out = OrcAlignment()
    ops = (C.c_uint32 * (m + n + 4))()
    rc = lib().orc_align(int(mode), C.byref(scoring), x, C.c_uint32(m), y, C.c_uint32(n),
                         C.byref(out), ops)
    if rc != 0:
        raise RuntimeError("oracle: traceback panic path (mod.rs:905)")
    d = {f: getattr(out, f) for f, _ in OrcAlignment._fields_}
    return d, [(int(v) & 7, int(v) >> 3) for v in ops[:out.n_ops]]


def align_batch(mode, scoring: OrcScoring, blob, x_off, x_len, y_off, y_len, threads=1,
                want_ops=True):
    """Batch -> (structured array of Alignment fields, ops uint32 flat, ops_off uint64, seconds)."""
    mode = MODES.get(mode, mode)
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    x_off = np.ascontiguousarray(x_off, dtype=np.uint64)
    y_off = np.ascontiguousarray(y_off, dtype=np.uint64)
    x_len = np.ascontiguousarray(x_len, dtype=np.uint32)
    y_len = np.ascontiguousarray(y_len, dtype=np.uint32)
    n = len(x_len)
    out = np.zeros(n, dtype=ALN_DTYPE)
    if want_ops:
        cap = x_len.astype(np.uint64) + y_len.astype(np.uint64) + np.uint64(4)
        ops_off = np.concatenate([[0], np.cumsum(cap)]).astype(np.uint64)
        ops = np.zeros(int(ops_off[-1]), dtype=np.uint32)
        ops_p, off_p = ops.ctypes.data_as(C.c_void_p), ops_off.ctypes.data_as(C.c_void_p)
    else:
        ops, ops_off, ops_p, off_p = None, None, None, None
    secs = lib().orc_align_batch(
        int(mode), C.byref(scoring), blob.ctypes.data_as(C.c_void_p),
        x_off.ctypes.data_as(C.c_void_p), x_len.ctypes.data_as(C.c_void_p),
        y_off.ctypes.data_as(C.c_void_p), y_len.ctypes.data_as(C.c_void_p),
        C.c_uint64(n), out.ctypes.data_as(C.c_void_p), ops_p, off_p, C.c_int(int(threads)))
    return out, ops, ops_off, secs


def hardware_threads() -> int:
    """Host threads worth starting: the CPUs in this process's affinity mask, capped by the cgroup CPU quota when
    there is one (a container that sees 128 CPUs but may only use 16 CPUs' worth of time runs 128 threads no
    faster than 16, and much less predictably)."""
    n = int(lib().orc_hardware_threads())
    q = cpu_quota()
    if q is not None:
        n = max(1, min(n, int(q + 0.999)))
    return n


def cpu_quota():
    """CPUs' worth of time the cgroup allows (cgroup v2 cpu.max, v1 cfs quota), or None if unlimited/unknown."""
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            a, b = f.read().split()[:2]
        if a != "max":
            return float(a) / float(b)
        return None
    except Exception:
        pass
    try:
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
            q = float(f.read())
        with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
            p = float(f.read())
        return q / p if q > 0 else None
    except Exception:
        return None


# ------------------------------------------------------------------ banded::Aligner + sparse pieces
def banded_align(mode, scoring: OrcScoring, k: int, w: int, x: bytes, y: bytes):
    mode = MODES.get(mode, mode)
    m, n = len(x), len(y)
    out = OrcAlignment()
    ops = (C.c_uint32 * (m + n + 8))()
    L = lib()
    L.orc_banded_align.restype = C.c_int
    rc = L.orc_banded_align(int(mode), C.byref(scoring), C.c_uint32(k), C.c_uint32(w), x, C.c_uint32(m), y,
                            C.c_uint32(n), C.byref(out), ops)
    if rc != 0:
        raise RuntimeError("banded oracle: the reference would panic on this input")
    d = {f: getattr(out, f) for f, _ in OrcAlignment._fields_}
    return d, [(int(v) & 7, int(v) >> 3) for v in ops[:out.n_ops]]


def banded_align_batch(mode, scoring: OrcScoring, k, w, blob, x_off, x_len, y_off, y_len, threads=1,
                       want_ops=True):
    """-> (fields, ops uint32 flat, ops_off, seconds, band cells)"""
    mode = MODES.get(mode, mode)
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    x_off = np.ascontiguousarray(x_off, dtype=np.uint64)
    y_off = np.ascontiguousarray(y_off, dtype=np.uint64)
    x_len = np.ascontiguousarray(x_len, dtype=np.uint32)
    y_len = np.ascontiguousarray(y_len, dtype=np.uint32)
    n = len(x_len)
    out = np.zeros(n, dtype=ALN_DTYPE)
    if want_ops:
        cap = x_len.astype(np.uint64) + y_len.astype(np.uint64) + np.uint64(8)
        ops_off = np.concatenate([[0], np.cumsum(cap)]).astype(np.uint64)
        ops = np.zeros(int(ops_off[-1]), dtype=np.uint32)
        ops_p, off_p = ops.ctypes.data_as(C.c_void_p), ops_off.ctypes.data_as(C.c_void_p)
    else:
        ops, ops_off, ops_p, off_p = None, None, None, None
    cells = C.c_uint64(0)
    L = lib()
    L.orc_banded_align_batch.restype = C.c_double
    secs = L.orc_banded_align_batch(
        int(mode), C.byref(scoring), C.c_uint32(k), C.c_uint32(w), blob.ctypes.data_as(C.c_void_p),
        x_off.ctypes.data_as(C.c_void_p), x_len.ctypes.data_as(C.c_void_p),
        y_off.ctypes.data_as(C.c_void_p), y_len.ctypes.data_as(C.c_void_p), C.c_uint64(n),
        out.ctypes.data_as(C.c_void_p), ops_p, off_p, C.byref(cells), C.c_int(int(threads)))
    return out, ops, ops_off, secs, int(cells.value)


def band_create(mode, scoring: OrcScoring, k, w, x: bytes, y: bytes):
    """Band::create -> (ranges [(start, end)] per column, num_cells)"""
    mode = MODES.get(mode, mode)
    n = len(y)
    ranges = (C.c_uint64 * (2 * (n + 1)))()
    cells = C.c_uint64(0)
    rc = lib().orc_band_create(int(mode), C.byref(scoring), C.c_uint32(k), C.c_uint32(w), x, C.c_uint32(len(x)),
                               y, C.c_uint32(n), ranges, C.byref(cells))
    if rc != 0:
        raise RuntimeError("banded oracle: Band::create would panic")
    return [(int(ranges[2 * j]), int(ranges[2 * j + 1])) for j in range(n + 1)], int(cells.value)


def band_ops(m, n, ops):
    """ops: [("entry"|"kmer", r, c, k, w)] applied to Band::new(m, n) -> ranges"""
    flat = (C.c_uint32 * (5 * len(ops)))()
    for t, (kind, r, c, k, w) in enumerate(ops):
        flat[5 * t:5 * t + 5] = [0 if kind == "entry" else 1, r, c, k, w]
    ranges = (C.c_uint64 * (2 * (n + 1)))()
    rc = lib().orc_band_ops(C.c_uint32(m), C.c_uint32(n), flat, C.c_uint32(len(ops)), ranges)
    assert rc == 0
    return [[int(ranges[2 * j]), int(ranges[2 * j + 1])] for j in range(n + 1)]


def find_kmer_matches(x: bytes, y: bytes, k: int):
    L = lib()
    L.orc_find_kmer_matches.restype = C.c_uint64
    cap = 1 << 16
    buf = (C.c_uint32 * (2 * cap))()
    cnt = L.orc_find_kmer_matches(x, C.c_uint32(len(x)), y, C.c_uint32(len(y)), C.c_uint32(k), buf, C.c_uint64(cap))
    assert cnt <= cap
    return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(cnt)]


def sdpkpp(matches, k, match_score, gap_open, gap_extend):
    n = len(matches)
    xy = (C.c_uint32 * (2 * max(1, n)))()
    for i, (a, b) in enumerate(matches):
        xy[2 * i], xy[2 * i + 1] = a, b
    path = (C.c_uint64 * max(1, n))()
    npath, score = C.c_uint64(0), C.c_uint32(0)
    rc = lib().orc_sdpkpp(xy, C.c_uint64(n), C.c_uint32(k), C.c_uint32(match_score), C.c_int32(gap_open),
                          C.c_int32(gap_extend), path, C.byref(npath), C.byref(score))
    assert rc == 0
    return [int(path[i]) for i in range(npath.value)], int(score.value)


def _xy(matches):
    n = len(matches)
    xy = (C.c_uint32 * (2 * max(1, n)))()
    for i, (a, b) in enumerate(matches):
        xy[2 * i], xy[2 * i + 1] = a, b
    return xy, n


def lcskpp(matches, k):
    """sparse::lcskpp (sparse.rs:67-143) -> (path, score)"""
    xy, n = _xy(matches)
    path = (C.c_uint64 * max(1, n))()
    npath, score = C.c_uint64(0), C.c_uint32(0)
    rc = lib().orc_lcskpp(xy, C.c_uint64(n), C.c_uint32(k), path, C.byref(npath), C.byref(score))
    if rc != 0:
        raise RuntimeError("lcskpp oracle: the reference would panic on this input")
    return [int(path[i]) for i in range(npath.value)], int(score.value)


def sdpkpp_union_lcskpp_path(matches, k, match_score, gap_open, gap_extend):
    """sparse::sdpkpp_union_lcskpp_path (sparse.rs:297-330) -> path"""
    xy, n = _xy(matches)
    path = (C.c_uint64 * max(1, 2 * n))()
    npath = C.c_uint64(0)
    rc = lib().orc_sdpkpp_union_lcskpp_path(xy, C.c_uint64(n), C.c_uint32(k), C.c_uint32(match_score),
                                            C.c_int32(gap_open), C.c_int32(gap_extend), path, C.byref(npath))
    if rc != 0:
        raise RuntimeError("union path oracle: the reference would panic on this input")
    return [int(path[i]) for i in range(npath.value)]


def expand_kmer_matches(x: bytes, y: bytes, k: int, matches, allowed_mismatches: int):
    """sparse::expand_kmer_matches (sparse.rs:404-498) -> sorted expanded matches"""
    xy, n = _xy(matches)
    L = lib()
    L.orc_expand_kmer_matches.restype = C.c_uint64
    cap = max(16, 4 * (n + 1) * (allowed_mismatches + 2))
    while True:
        out = (C.c_uint32 * (2 * cap))()
        got = L.orc_expand_kmer_matches(x, C.c_uint32(len(x)), y, C.c_uint32(len(y)), C.c_uint32(k), xy,
                                        C.c_uint64(n), C.c_uint32(allowed_mismatches), out, C.c_uint64(cap))
        if got == 0xFFFFFFFFFFFFFFFF:
            raise RuntimeError("expand_kmer_matches oracle: the reference would panic on this input")
        if got <= cap:
            return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(got)]
        cap = int(got)


def banded_align_hinted(scoring: OrcScoring, k: int, w: int, x: bytes, y: bytes, matches, path=None,
                        allowed_mismatches=None, use_lcskpp_union=False):
    """banded::Aligner::custom_with_matches / custom_with_expanded_matches / custom_with_match_path
    (banded.rs:313-401) -> (fields, ops, band cells); None when the reference would panic."""
    m, n = len(x), len(y)
    xy, nm = _xy(matches)
    pp = (C.c_uint64 * max(1, len(path or [])))(*(path or []))
    out = OrcAlignment()
    ops = (C.c_uint32 * (m + n + 8))()
    cells = C.c_uint64(0)
    L = lib()
    L.orc_banded_align_hinted.restype = C.c_int
    rc = L.orc_banded_align_hinted(C.byref(scoring), C.c_uint32(k), C.c_uint32(w), x, C.c_uint32(m), y, C.c_uint32(n),
                                   xy, C.c_uint64(nm), pp, C.c_uint64(len(path or [])),
                                   C.c_int32(1 if path is not None else 0),
                                   C.c_int32(-1 if allowed_mismatches is None else int(allowed_mismatches)),
                                   C.c_int32(1 if use_lcskpp_union else 0), C.byref(out), ops, C.byref(cells))
    if rc != 0:
        return None
    d = {f: getattr(out, f) for f, _ in OrcAlignment._fields_}
    return d, [(int(v) & 7, int(v) >> 3) for v in ops[:out.n_ops]], int(cells.value)
