"""ctypes binding of libbgalign.so (include/bgalign.h).

There is no fallback of any kind: if the shared library is missing, or there is no CUDA device,
every compute entry point raises.  The library is built in-tree by `__graft_entry__.build()` /
`make -C biogarden_b200/csrc`."""
import ctypes as C
import os

import numpy as np

from .error import EngineError, InvalidArgumentRange, InvalidInputSize

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbgalign.so")

BG_OK, BG_EINVAL_RANGE, BG_EINVAL_SIZE, BG_ECUDA, BG_ENOMEM, BG_EINVAL_ARG, BG_EINVAL_RESIDUE, BG_ENODEVICE, \
    BG_EUNSUPPORTED, BG_EINVAL_FASTA = range(10)
MODES = {"global": 0, "local": 1, "semiglobal": 2, "fitting": 3, "overlap": 4}
F_SCORE_ONLY = 1
ST_OK, ST_REF_UNDEFINED = 0, 1


class bg_batch(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("residues", C.c_void_p), ("seq_off", C.c_void_p),
                ("packing", C.c_uint32), ("reserved_", C.c_uint32), ("alphabet", C.c_void_p)]


PACK_NONE, PACK_2BIT, PACK_5BIT = 0, 2, 5


class bg_params(C.Structure):
    _fields_ = [("mode", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("flags", C.c_uint32),
                ("table", C.c_void_p), ("n_rows", C.c_int32), ("n_cols", C.c_int32),
                ("row_code", C.c_void_p), ("col_code", C.c_void_p)]


class bg_result(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("score", C.c_void_p), ("status", C.c_void_p), ("arena", C.c_void_p),
                ("off", C.c_void_p), ("owner_", C.c_void_p)]


class bg_ops_result(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("score", C.c_void_p), ("status", C.c_void_p), ("len", C.c_void_p),
                ("first", C.c_void_p), ("ops", C.c_void_p), ("ops_off", C.c_void_p), ("owner_", C.c_void_p)]


class bg_fasta(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("residues", C.c_void_p), ("seq_off", C.c_void_p),
                ("ids", C.c_void_p), ("id_off", C.c_void_p),
                ("packing", C.c_uint32), ("reserved_", C.c_uint32), ("alphabet", C.c_uint8 * 32)]


class bg_timing(C.Structure):
    _fields_ = [("encode_ms", C.c_double), ("fill_ms", C.c_double), ("walk_ms", C.c_double),
                ("compact_ms", C.c_double), ("total_ms", C.c_double), ("cells", C.c_uint64),
                ("launches", C.c_uint64), ("trace_bytes", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("cells_packed16", C.c_uint64), ("cells_bitparallel", C.c_uint64),
                ("fill_launches", C.c_uint64), ("cells_refilled", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/bgalign.h declares (tests check the library exports all of them)
SYMBOLS = ["bg_create", "bg_destroy", "bg_strerror", "bg_last_error", "bg_version", "bg_align_batch",
           "bg_result_free", "bg_edit_distance_batch", "bg_hamming_distance_batch", "bg_p_distance_matrix", "bg_batch_upload", "bg_dbatch_free", "bg_align_device",
           "bg_edit_distance_device", "bg_dresult_download", "bg_dresult_download_u64", "bg_dresult_free",
           "bg_sync", "bg_stream", "bg_device_ordinal", "bg_last_timing", "bg_batch_prepare", "bg_set_shape",
           "bg_set_trace_budget", "bg_set_long_trace_budget", "bg_set_host_plan", "bg_set_fine_pairs", "bg_fasta_parse", "bg_fasta_free", "bg_pin_host", "bg_unpin_host", "bg_score_table26", "bg_residue_histogram", "bg_ref_status",
           "bg_align_batch_ops", "bg_ops_result_free", "bg_expand_ops", "bg_expand_kind",
           "bg_fasta_parse_packed", "bg_packed_bytes", "bg_pack_residues", "bg_unpack_residues"]

_lib = None


def lib():
    """Load libbgalign.so; raises (never falls back) when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError("libbgalign.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C biogarden_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, ci = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_int
    L.bg_create.restype = ci; L.bg_create.argtypes = [C.POINTER(ci), ci, C.POINTER(vp)]
    L.bg_destroy.restype = None; L.bg_destroy.argtypes = [vp]
    L.bg_strerror.restype = C.c_char_p; L.bg_strerror.argtypes = [ci]
    L.bg_last_error.restype = C.c_char_p; L.bg_last_error.argtypes = [vp]
    L.bg_version.restype = ci
    L.bg_align_batch.restype = ci
    L.bg_align_batch.argtypes = [vp, C.POINTER(bg_batch), C.POINTER(bg_params), C.POINTER(bg_result)]
    L.bg_result_free.restype = None; L.bg_result_free.argtypes = [C.POINTER(bg_result)]
    L.bg_align_batch_ops.restype = ci
    L.bg_align_batch_ops.argtypes = [vp, C.POINTER(bg_batch), C.POINTER(bg_params), C.POINTER(bg_ops_result)]
    L.bg_ops_result_free.restype = None; L.bg_ops_result_free.argtypes = [C.POINTER(bg_ops_result)]
    L.bg_expand_ops.restype = ci; L.bg_expand_ops.argtypes = [vp, vp, vp, u64, vp, vp]
    L.bg_expand_ops_impl.restype = ci; L.bg_expand_ops_impl.argtypes = [ci, vp, vp, vp, u64, vp, vp]
    L.bg_expand_kind.restype = C.c_char_p; L.bg_expand_kind.argtypes = []
    L.bg_edit_distance_batch.restype = ci; L.bg_edit_distance_batch.argtypes = [vp, C.POINTER(bg_batch), vp]
    L.bg_batch_upload.restype = ci; L.bg_batch_upload.argtypes = [vp, ci, C.POINTER(bg_batch), C.POINTER(vp)]
    L.bg_dbatch_free.restype = None; L.bg_dbatch_free.argtypes = [vp]
    L.bg_align_device.restype = ci; L.bg_align_device.argtypes = [vp, vp, C.POINTER(bg_params), C.POINTER(vp)]
    L.bg_edit_distance_device.restype = ci; L.bg_edit_distance_device.argtypes = [vp, vp, C.POINTER(vp)]
    L.bg_dresult_download.restype = ci; L.bg_dresult_download.argtypes = [vp, vp, C.POINTER(bg_result)]
    L.bg_dresult_download_u64.restype = ci; L.bg_dresult_download_u64.argtypes = [vp, vp, vp]
    L.bg_dresult_free.restype = None; L.bg_dresult_free.argtypes = [vp]
    L.bg_sync.restype = ci; L.bg_sync.argtypes = [vp]
    L.bg_stream.restype = vp; L.bg_stream.argtypes = [vp, ci]
    L.bg_device_ordinal.restype = ci; L.bg_device_ordinal.argtypes = [vp, ci]
    L.bg_last_timing.restype = ci; L.bg_last_timing.argtypes = [vp, C.POINTER(bg_timing)]
    L.bg_batch_prepare.restype = ci; L.bg_batch_prepare.argtypes = [vp, vp, ci]
    L.bg_hamming_distance_batch.restype = ci; L.bg_hamming_distance_batch.argtypes = [vp, C.POINTER(bg_batch), vp]
    L.bg_p_distance_matrix.restype = ci; L.bg_p_distance_matrix.argtypes = [vp, vp, vp, u64, vp]
    L.bg_fasta_parse.restype = ci; L.bg_fasta_parse.argtypes = [vp, u64, ci, C.POINTER(bg_fasta)]
    L.bg_fasta_free.restype = None; L.bg_fasta_free.argtypes = [C.POINTER(bg_fasta)]
    L.bg_fasta_parse_packed.restype = ci; L.bg_fasta_parse_packed.argtypes = [vp, u64, ci, ci, C.POINTER(bg_fasta)]
    L.bg_packed_bytes.restype = u64; L.bg_packed_bytes.argtypes = [u64, ci]
    L.bg_pack_residues.restype = ci; L.bg_pack_residues.argtypes = [vp, u64, ci, ci, vp, vp]
    L.bg_unpack_residues.restype = ci; L.bg_unpack_residues.argtypes = [vp, ci, vp, u64, u64, vp]
    L.bg_pin_host.restype = ci; L.bg_pin_host.argtypes = [vp, u64]
    L.bg_unpin_host.restype = ci; L.bg_unpin_host.argtypes = [vp]
    L.bg_set_shape.restype = ci; L.bg_set_shape.argtypes = [vp, ci, ci]
    L.bg_set_trace_budget.restype = ci; L.bg_set_trace_budget.argtypes = [vp, u64]
    L.bg_set_long_trace_budget.restype = ci; L.bg_set_long_trace_budget.argtypes = [vp, u64]
    L.bg_set_host_plan.restype = ci; L.bg_set_host_plan.argtypes = [vp, ci]
    L.bg_set_fine_pairs.restype = ci; L.bg_set_fine_pairs.argtypes = [vp, ci]
    L.bg_debug_plan_compare.restype = ci; L.bg_debug_plan_compare.argtypes = [vp, C.POINTER(bg_batch), C.POINTER(bg_params), vp]
    L.bg_score_table26.restype = C.POINTER(C.c_int8); L.bg_score_table26.argtypes = [C.c_char_p]
    L.bg_residue_histogram.restype = ci; L.bg_residue_histogram.argtypes = [C.POINTER(bg_batch), vp, vp]
    L.bg_ref_status.restype = ci; L.bg_ref_status.argtypes = [ci, u64, u64, i32, ci]
    _lib = L
    return L


def check(rc, ctx=None):
    if rc == BG_OK:
        return
    L = lib()
    msg = L.bg_strerror(rc).decode()
    if ctx is not None:
        extra = L.bg_last_error(ctx).decode()
        if extra:
            msg += ": " + extra
    if rc == BG_EINVAL_RANGE:
        raise InvalidArgumentRange(msg)
    if rc == BG_EINVAL_SIZE:
        raise InvalidInputSize(msg)
    raise EngineError("bgalign error %d: %s" % (rc, msg))


def score_table26(name: str) -> np.ndarray:
    p = lib().bg_score_table26(name.encode())
    if not p:
        raise KeyError(name)
    return np.ctypeslib.as_array(p, shape=(26, 26)).astype(np.int32)


class Batch:
    """Host-side batch in the C ABI's layout (keeps the numpy arrays alive)."""

    def __init__(self, residues: np.ndarray, seq_off: np.ndarray, packing: int = PACK_NONE, alphabet=None):
        """packing != PACK_NONE: `residues` holds the packed bytes (bgalign.h BG_PACK_*), `alphabet` the code -> byte map."""
        self.residues = np.ascontiguousarray(residues, dtype=np.uint8)
        self.seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        assert self.seq_off.ndim == 1 and len(self.seq_off) % 2 == 1
        self.n_pairs = (len(self.seq_off) - 1) // 2
        self.packing = int(packing)
        self.alphabet = None if alphabet is None else np.ascontiguousarray(np.frombuffer(bytes(alphabet).ljust(32, b"\0"), np.uint8))
        self.c = bg_batch(self.n_pairs, self.residues.ctypes.data if self.residues.size else None,
                          self.seq_off.ctypes.data, self.packing, 0,
                          self.alphabet.ctypes.data if self.alphabet is not None else None)

    def pack(self, bits: int) -> "Batch":
        """The same pairs with the residues packed at `bits` per residue (bg_pack_residues)."""
        assert self.packing == PACK_NONE
        n = int(self.seq_off[-1]) if self.seq_off.size else 0
        assert int(self.seq_off[0]) == 0, "pack() expects offsets that start at 0"
        out = np.zeros(int(lib().bg_packed_bytes(n, bits)), np.uint8)
        alpha = np.zeros(32, np.uint8)
        check(lib().bg_pack_residues(self.residues.ctypes.data, n, bits, 0, out.ctypes.data, alpha.ctypes.data))
        return Batch(out, self.seq_off, bits, bytes(alpha))

    def unpacked_residues(self) -> np.ndarray:
        """One byte per residue, whatever the packing."""
        if self.packing == PACK_NONE:
            return self.residues
        n = int(self.seq_off[-1])
        out = np.zeros(max(1, n), np.uint8)
        check(lib().bg_unpack_residues(self.residues.ctypes.data, self.packing, self.alphabet.ctypes.data, 0, n, out.ctypes.data))
        return out[:n]

    @classmethod
    def from_sequences(cls, seqs):
        """seqs: flat list of bytes-like, pair p = (seqs[2p], seqs[2p+1])."""
        if len(seqs) % 2:
            raise InvalidInputSize("a Tile of pairs needs an even number of sequences")
        lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=len(seqs))
        off = np.zeros(len(seqs) + 1, np.uint64)
        np.cumsum(lens, out=off[1:])
        res = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8) if len(seqs) else np.zeros(0, np.uint8)
        return cls(res, off)

    def lengths(self):
        d = np.diff(self.seq_off)
        return d[0::2], d[1::2]

    def cells(self) -> int:
        n, m = self.lengths()
        return int(np.sum(n.astype(np.float64) * m.astype(np.float64)))

    def histograms(self):
        ha = np.zeros(256, np.uint64)
        hb = np.zeros(256, np.uint64)
        check(lib().bg_residue_histogram(C.byref(self.c), ha.ctypes.data, hb.ctypes.data))
        return ha, hb


class Params:
    """bg_params plus the arrays it points to."""

    def __init__(self, mode, gap_open, gap_extend, table, row_code, col_code, score_only=False):
        self.table = np.ascontiguousarray(table, dtype=np.int32)
        self.row_code = np.ascontiguousarray(row_code, dtype=np.uint8)
        self.col_code = np.ascontiguousarray(col_code, dtype=np.uint8)
        assert self.table.ndim == 2 and self.row_code.shape == (256,) and self.col_code.shape == (256,)
        m = MODES[mode] if isinstance(mode, str) else int(mode)
        self.c = bg_params(m, int(gap_open), int(gap_extend), F_SCORE_ONLY if score_only else 0,
                           self.table.ctypes.data, self.table.shape[0], self.table.shape[1],
                           self.row_code.ctypes.data, self.col_code.ctypes.data)


class Result:
    """Owns a bg_result; numpy views are valid until close()."""

    def __init__(self, c_res: bg_result):
        self._c = c_res
        n = int(c_res.n_pairs)
        self.n_pairs = n

        def view(ptr, dtype, count):
            if count == 0 or not ptr:
                return np.zeros(0, dtype)
            buf = (C.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=count)
        self.score = view(c_res.score, np.int32, n)
        self.status = view(c_res.status, np.uint8, n)
        self.off = view(c_res.off, np.uint64, 2 * n + 1)
        total = int(self.off[-1]) if n else 0
        self.arena = view(c_res.arena, np.uint8, total)

    def strings(self, p):
        o0, o1, o2 = int(self.off[2 * p]), int(self.off[2 * p + 1]), int(self.off[2 * p + 2])
        return bytes(self.arena[o0:o1]), bytes(self.arena[o1:o2])

    def close(self):
        if self._c is not None:
            lib().bg_result_free(C.byref(self._c))
            self._c = None
            self.score = self.status = self.off = self.arena = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _view(ptr, dtype, count):
    if count == 0 or not ptr:
        return np.zeros(0, dtype)
    buf = (C.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


class OpsResult:
    """Owns a bg_ops_result (compact results: 2-bit ops per alignment column); numpy views valid until close()."""

    def __init__(self, c_res: bg_ops_result, batch: "Batch"):
        self._c = c_res
        self.batch = batch
        n = int(c_res.n_pairs)
        self.n_pairs = n
        self.score = _view(c_res.score, np.int32, n)
        self.status = _view(c_res.status, np.uint8, n)
        self.len = _view(c_res.len, np.uint32, n)
        self.first = _view(c_res.first, np.uint32, 2 * n)
        self.ops_off = _view(c_res.ops_off, np.uint64, n + 1)
        self.ops = _view(c_res.ops, np.uint32, int(self.ops_off[-1]) if n else 0)

    def strings(self, p, impl=None):
        """(a_align, b_align) of pair p through bg_expand_ops (impl: None = the library's choice, 0 scalar, 1 AVX-512)."""
        ln = int(self.len[p])
        a = np.zeros(max(1, ln), np.uint8); b = np.zeros(max(1, ln), np.uint8)
        bt = self.batch
        s1 = bt.residues.ctypes.data + int(bt.seq_off[2 * p]) + int(self.first[2 * p])
        s2 = bt.residues.ctypes.data + int(bt.seq_off[2 * p + 1]) + int(self.first[2 * p + 1])
        ops = self.ops.ctypes.data + 4 * int(self.ops_off[p]) if self.ops.size else None
        if impl is None:
            check(lib().bg_expand_ops(s1, s2, ops, ln, a.ctypes.data, b.ctypes.data))
        else:
            check(lib().bg_expand_ops_impl(impl, s1, s2, ops, ln, a.ctypes.data, b.ctypes.data))
        return bytes(a[:ln]), bytes(b[:ln])

    def close(self):
        if self._c is not None:
            lib().bg_ops_result_free(C.byref(self._c))
            self._c = None
            self.score = self.status = self.len = self.first = self.ops = self.ops_off = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def expand_ops(s1: bytes, s2: bytes, ops: np.ndarray, length: int, impl=None):
    """bg_expand_ops on host buffers (s1 / s2 start at the first residue the alignment consumes)."""
    ops = np.ascontiguousarray(ops, np.uint32)
    a = np.zeros(max(1, length), np.uint8); b = np.zeros(max(1, length), np.uint8)
    b1 = np.frombuffer(s1 + b"\0", np.uint8); b2 = np.frombuffer(s2 + b"\0", np.uint8)
    if impl is None:
        check(lib().bg_expand_ops(b1.ctypes.data, b2.ctypes.data, ops.ctypes.data, length, a.ctypes.data, b.ctypes.data))
    else:
        check(lib().bg_expand_ops_impl(impl, b1.ctypes.data, b2.ctypes.data, ops.ctypes.data, length, a.ctypes.data, b.ctypes.data))
    return bytes(a[:length]), bytes(b[:length])


def expand_kind() -> str:
    return lib().bg_expand_kind().decode()


class Context:
    """bg_ctx: one engine instance (SequenceAligner::new, aligner.rs:44)."""

    def __init__(self, devices=None):
        L = lib()
        h = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            rc = L.bg_create(arr, len(devices), C.byref(h))
        else:
            rc = L.bg_create(None, 0, C.byref(h))
        if rc == BG_ENODEVICE:
            raise EngineError("no usable CUDA device: the alignment engine has no CPU path")
        check(rc)
        self.h = h
        self.n_devices = len(devices) if devices else 1

    def close(self):
        if getattr(self, "h", None):
            lib().bg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-buffer path ----
    def align_batch(self, batch: Batch, params: Params) -> Result:
        r = bg_result()
        check(lib().bg_align_batch(self.h, C.byref(batch.c), C.byref(params.c), C.byref(r)), self.h)
        return Result(r)

    def align_batch_ops(self, batch: Batch, params: Params) -> OpsResult:
        r = bg_ops_result()
        check(lib().bg_align_batch_ops(self.h, C.byref(batch.c), C.byref(params.c), C.byref(r)), self.h)
        return OpsResult(r, batch)

    def edit_distance_batch(self, batch: Batch) -> np.ndarray:
        out = np.zeros(batch.n_pairs, np.uint64)
        check(lib().bg_edit_distance_batch(self.h, C.byref(batch.c), out.ctypes.data), self.h)
        return out

    def hamming_distance_batch(self, batch: Batch) -> np.ndarray:
        """seq.rs:74-83 for every pair; raises InvalidInputSize when a pair has unequal lengths."""
        out = np.zeros(batch.n_pairs, np.uint64)
        check(lib().bg_hamming_distance_batch(self.h, C.byref(batch.c), out.ctypes.data), self.h)
        return out

    def p_distance_matrix(self, residues: np.ndarray, seq_off: np.ndarray) -> np.ndarray:
        """stat.rs:138-152: rows x rows float32."""
        residues = np.ascontiguousarray(residues, np.uint8)
        seq_off = np.ascontiguousarray(seq_off, np.uint64)
        rows = len(seq_off) - 1
        out = np.zeros((max(rows, 0), max(rows, 0)), np.float32)
        check(lib().bg_p_distance_matrix(self.h, residues.ctypes.data, seq_off.ctypes.data, max(rows, 0), out.ctypes.data), self.h)
        return out

    # ---- device-resident path ----
    def upload(self, batch: Batch, dev_index=0, prepare=None):
        h = C.c_void_p()
        check(lib().bg_batch_upload(self.h, dev_index, C.byref(batch.c), C.byref(h)), self.h)
        if prepare is not None:
            check(lib().bg_batch_prepare(self.h, h, 1 if prepare == "edit" else 0), self.h)
        return h

    def free_batch(self, h):
        lib().bg_dbatch_free(h)

    def align_device(self, dbatch, params: Params):
        h = C.c_void_p()
        check(lib().bg_align_device(self.h, dbatch, C.byref(params.c), C.byref(h)), self.h)
        return h

    def edit_distance_device(self, dbatch):
        h = C.c_void_p()
        check(lib().bg_edit_distance_device(self.h, dbatch, C.byref(h)), self.h)
        return h

    def download(self, dres) -> Result:
        r = bg_result()
        check(lib().bg_dresult_download(self.h, dres, C.byref(r)), self.h)
        return Result(r)

    def download_u64(self, dres, n) -> np.ndarray:
        out = np.zeros(n, np.uint64)
        check(lib().bg_dresult_download_u64(self.h, dres, out.ctypes.data), self.h)
        return out

    def free_result(self, dres):
        lib().bg_dresult_free(dres)

    def sync(self):
        check(lib().bg_sync(self.h), self.h)

    def stream(self, dev_index=0):
        return lib().bg_stream(self.h, dev_index)

    def timing(self) -> dict:
        t = bg_timing()
        check(lib().bg_last_timing(self.h, C.byref(t)), self.h)
        return t.as_dict()

    def set_shape(self, L_, C_):
        check(lib().bg_set_shape(self.h, L_, C_), self.h)

    def set_trace_budget(self, nbytes):
        check(lib().bg_set_trace_budget(self.h, nbytes), self.h)

    def set_fine_pairs(self, max_pairs: int):
        """Long-pair launches of at most max_pairs pairs run on the fine-grained wavefront kernel (0: never)."""
        check(lib().bg_set_fine_pairs(self.h, int(max_pairs)), self.h)

    def set_host_plan(self, on: bool):
        """True: launch plans are built on the host; False (default): pipeline chunks are planned on the device."""
        check(lib().bg_set_host_plan(self.h, int(bool(on))), self.h)

    def plan_compare(self, batch: Batch, params: Params):
        """(eligible, descriptors compared, differing, non-empty surplus slots): host planner vs device planner."""
        out = np.zeros(4, np.uint64)
        check(lib().bg_debug_plan_compare(self.h, C.byref(batch.c), C.byref(params.c), out.ctypes.data), self.h)
        return tuple(int(x) for x in out)

    def set_long_trace_budget(self, nbytes):
        """Trace memory per launch of the long-pair path; pairs that need more use bounded-memory traceback."""
        check(lib().bg_set_long_trace_budget(self.h, nbytes), self.h)


_synth = None
SYNTH_LIB_PATH = os.path.join(_HERE, "libbgsynth.so")


def synth_lib():
    """libbgsynth.so (include/bgsynth.h): the workload generator of tests and bench.py -- tooling, kept out of the
    product library so that a process that only generates inputs (bench.py --impl reference) never maps libbgalign.so."""
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise EngineError("libbgsynth.so not built (%s): run `make -C biogarden_b200/csrc`" % SYNTH_LIB_PATH)
        S = C.CDLL(SYNTH_LIB_PATH)
        S.bg_synth_pairs.restype = C.c_int
        S.bg_synth_pairs.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_char_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int,
                                     C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        _synth = S
    return _synth


def _check_synth(rc):
    if rc:
        raise EngineError("bg_synth_pairs: invalid argument")


def synth_pairs(seed, first_pair, n_pairs, alphabet: bytes, len_lo, len_hi, resize_b=True) -> Batch:
    """Deterministic synthetic workload (SURVEY 8d generator)."""
    L = synth_lib()
    off = np.zeros(2 * n_pairs + 1, np.uint64)
    tot = C.c_uint64(0)
    _check_synth(L.bg_synth_pairs(seed, first_pair, n_pairs, alphabet, len(alphabet), len_lo, len_hi, int(resize_b), None,
                                  off.ctypes.data, C.byref(tot)))
    res = np.zeros(max(1, tot.value), np.uint8)
    _check_synth(L.bg_synth_pairs(seed, first_pair, n_pairs, alphabet, len(alphabet), len_lo, len_hi, int(resize_b),
                                  res.ctypes.data, off.ctypes.data, C.byref(tot)))
    return Batch(res[:tot.value], off)
