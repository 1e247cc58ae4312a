"""ctypes binding of oracle/liboracle.so -- test infrastructure only (see oracle/oracle.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

MODES = {"global": 0, "local": 1, "semiglobal": 2, "fitting": 3, "overlap": 4}
SCORERS = {"blosum62": 0, "pam250": 1, "unit": 2, "table": 3}
OK, ERR_RANGE, ERR_SIZE, PANIC, HANG = 0, 1, 2, 3, 4
UNDEFINED = (PANIC, HANG)

_lib = None


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle.cpp", "oracle.h", "score_tables.inc", "Makefile")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        u8p, i32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
        L.orc_aligner_new.restype = C.c_void_p
        L.orc_aligner_free.argtypes = [C.c_void_p]
        L.orc_align.restype = C.c_int
        L.orc_align.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int,
                                C.c_void_p, C.c_int32, C.c_int32, i32p, C.c_char_p, C.c_char_p, C.c_size_t,
                                C.POINTER(C.c_size_t)]
        L.orc_align_lean.restype = C.c_int
        L.orc_align_lean.argtypes = L.orc_align.argtypes[1:]
        for f in (L.orc_edit_distance, L.orc_edit_distance_lean):
            f.restype = C.c_int
            f.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, u64p]
        L.orc_align_batch.restype = C.c_double
        L.orc_align_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int32,
                                      C.c_int32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
        L.orc_edit_distance_batch.restype = C.c_double
        L.orc_edit_distance_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        L.orc_hw_threads.restype = C.c_int
        L.orc_hamming_distance.restype = C.c_int
        L.orc_hamming_distance.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, u64p]
        L.orc_p_distance_matrix.restype = C.c_int
        L.orc_p_distance_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def _table_ptr(table):
    if table is None:
        return None, None
    t = np.ascontiguousarray(table, dtype=np.int32)
    assert t.shape == (256, 256)
    return t, t.ctypes.data


def align(mode, s1, s2, scorer, a, b, table=None, lean=False, aligner=None):
    """Returns (status, score, a_align, b_align). scorer: name or 'table' with a 256x256 table."""
    L = lib()
    s1, s2 = bytes(s1), bytes(s2)
    cap = len(s1) + len(s2) + 1
    o1, o2 = C.create_string_buffer(cap), C.create_string_buffer(cap)
    sc, ln = C.c_int32(0), C.c_size_t(0)
    keep, tp = _table_ptr(table)
    args = [MODES[mode], s1, len(s1), s2, len(s2), SCORERS[scorer], tp, a, b, C.byref(sc), o1, o2, cap, C.byref(ln)]
    if lean:
        st = L.orc_align_lean(*args)
    else:
        own = aligner is None
        h = L.orc_aligner_new() if own else aligner
        st = L.orc_align(h, *args)
        if own:
            L.orc_aligner_free(h)
    return st, sc.value, o1.raw[:ln.value], o2.raw[:ln.value]


def edit_distance(s1, s2, lean=False):
    L = lib()
    out = C.c_uint64(0)
    f = L.orc_edit_distance_lean if lean else L.orc_edit_distance
    f(bytes(s1), len(s1), bytes(s2), len(s2), C.byref(out))
    return out.value


def hamming_distance(s1, s2):
    """(status, distance): status ORC_ERR_SIZE (2) when the lengths differ (seq.rs:81)."""
    out = C.c_uint64(0)
    st = lib().orc_hamming_distance(bytes(s1), len(s1), bytes(s2), len(s2), C.byref(out))
    return st, out.value


def p_distance_matrix(rows):
    """rows: list of bytes -> float32 (R, R) array, the reference's arithmetic (stat.rs:138-152)."""
    res = np.frombuffer(b"".join(bytes(r) for r in rows) or b"\0", dtype=np.uint8).copy()
    off = np.zeros(len(rows) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in rows])
    out = np.zeros((len(rows), len(rows)), np.float32)
    st = lib().orc_p_distance_matrix(res.ctypes.data, off.ctypes.data, len(rows), out.ctypes.data)
    if st:
        raise RuntimeError("reference panics (empty matrix)")
    return out


def align_batch(mode, residues, seq_off, scorer, a, b, table=None, threads=1, lean=False, want_strings=True,
                fresh=False):
    """Batch driver. residues: uint8 array; seq_off: uint64[2n+1].
    Returns dict(score, status, len, hash, arena, out_off, seconds)."""
    L = lib()
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    n = (len(seq_off) - 1) // 2
    score = np.zeros(n, np.int32)
    status = np.zeros(n, np.uint8)
    ln = np.zeros(n, np.uint64)
    hs = np.zeros(n, np.uint64)
    keep, tp = _table_ptr(table)
    arena = out_off = None
    ap = op = None
    if want_strings:
        caps = (seq_off[2::2] - seq_off[0:-1:2]).astype(np.uint64)  # n_p + m_p
        out_off = np.zeros(n + 1, np.uint64)
        np.cumsum(2 * caps, out=out_off[1:])
        arena = np.zeros(int(out_off[-1]) + 1, np.uint8)
        ap, op = arena.ctypes.data, out_off.ctypes.data
    secs = L.orc_align_batch(MODES[mode], residues.ctypes.data, seq_off.ctypes.data, n, SCORERS[scorer], tp, a, b,
                             threads, int(lean) | (2 if fresh else 0), score.ctypes.data, status.ctypes.data, ln.ctypes.data,
                             hs.ctypes.data, ap, op)
    return dict(score=score, status=status, len=ln, hash=hs, arena=arena, out_off=out_off, seconds=secs)


def batch_strings(res, seq_off, p):
    """(a_align, b_align) of pair p from an align_batch(want_strings=True) result."""
    cap = int(seq_off[2 * p + 2] - seq_off[2 * p])
    o = int(res["out_off"][p]); ln = int(res["len"][p])
    return bytes(res["arena"][o:o + ln]), bytes(res["arena"][o + cap:o + cap + ln])


def edit_distance_batch(residues, seq_off, threads=1, lean=False):
    L = lib()
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    n = (len(seq_off) - 1) // 2
    out = np.zeros(n, np.uint64)
    secs = L.orc_edit_distance_batch(residues.ctypes.data, seq_off.ctypes.data, n, threads, int(lean), out.ctypes.data)
    return out, secs


def hw_threads():
    return lib().orc_hw_threads()
