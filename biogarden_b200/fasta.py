"""FASTA ingest that feeds the path (reference src/io/fasta.rs:95-136).

Same record grammar as the reference reader: a record starts at a line beginning with '>',
id = first whitespace-delimited token of the header, description = the rest, sequence =
the following lines concatenated after trim_end, up to the next '>' or EOF."""
import io
import os
from typing import Optional

from .sequence import Sequence, Tile


class Record:
    __slots__ = ("id", "desc", "seq")

    def __init__(self, id: str = "", desc: Optional[str] = None, seq: str = ""):
        self.id, self.desc, self.seq = id, desc, seq

    def is_empty(self):   # fasta.rs:232-234
        return not self.id and self.desc is None and not self.seq

    def clear(self):
        self.id, self.desc, self.seq = "", None, ""


class Reader:
    def __init__(self, stream):
        self._f = stream
        self._line = ""

    @classmethod
    def from_file(cls, path):   # fasta.rs:35-39
        return cls(open(os.fspath(path), "r", newline=""))

    @classmethod
    def from_string(cls, text: str):
        return cls(io.StringIO(text))

    def read(self, record: Record):   # fasta.rs:95-123
        record.clear()
        if not self._line:
            self._line = self._f.readline()
            if not self._line:
                return
        if not self._line.startswith(">"):
            raise IOError("Expected > at record start.")
        # splitn(2, char::is_whitespace) on the right-trimmed header: the id ends at the FIRST whitespace
        # character (it is empty when the header starts with one), the description is everything after it
        hdr = self._line[1:].rstrip()
        cut = next((k for k, ch in enumerate(hdr) if ch.isspace()), None)
        record.id = hdr if cut is None else hdr[:cut]
        record.desc = None if cut is None else hdr[cut + 1:]
        parts = []
        while True:
            self._line = self._f.readline()
            if not self._line or self._line.startswith(">"):
                break
            parts.append(self._line.rstrip())
        record.seq = "".join(parts)

    def read_all(self, tile: Tile):   # fasta.rs:125-136
        rec = Record()
        while True:
            self.read(rec)
            if rec.is_empty():
                break
            tile.push(Sequence(rec.seq, id=rec.id))


def read_tile(path) -> Tile:
    t = Tile()
    Reader.from_file(path).read_all(t)
    return t


# ---- native ingest: FASTA text -> the C ABI's batch layout (csrc/fasta_ingest.cpp) ---------------------
def parse_batch(text: bytes, n_threads: int = 0, bits: int = 0):
    """io::fasta::Reader::read_all over `text`, done by the native multi-threaded parser, straight into the layout
    the alignment entry points take.  Returns (ids, residues uint8 array, seq_off uint64[n + 1]); with bits = 2 / 5
    the residues are packed (bg_fasta_parse_packed) and a fourth value, the alphabet (bytes), is returned.  Raises
    IOError ("Expected > at record start.") like the reference."""
    import ctypes as C
    import numpy as np
    from . import native
    L = native.lib()
    f = native.bg_fasta()
    buf = bytes(text)
    rc = L.bg_fasta_parse_packed(buf, len(buf), n_threads, bits, C.byref(f)) if bits else L.bg_fasta_parse(buf, len(buf), n_threads, C.byref(f))
    if rc == native.BG_EINVAL_FASTA:
        raise IOError("Expected > at record start.")
    native.check(rc)
    try:
        n = f.n_records
        seq_off = np.ctypeslib.as_array(C.cast(f.seq_off, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        id_off = np.ctypeslib.as_array(C.cast(f.id_off, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        nres, nid = int(seq_off[n]), int(id_off[n])
        alphabet = bytes(f.alphabet)
        if bits:
            nres = int(L.bg_packed_bytes(nres, bits))     # the packed arena incl. its padding
        residues = np.ctypeslib.as_array(C.cast(f.residues, C.POINTER(C.c_uint8)), shape=(max(nres, 1),))[:nres].copy()
        idbytes = bytes(np.ctypeslib.as_array(C.cast(f.ids, C.POINTER(C.c_uint8)), shape=(max(nid, 1),))[:nid])
        ids = [idbytes[int(id_off[r]):int(id_off[r + 1])].decode("utf-8", "replace") for r in range(n)]
    finally:
        L.bg_fasta_free(C.byref(f))
    return (ids, residues, seq_off, alphabet) if bits else (ids, residues, seq_off)


def read_batch(path, n_threads: int = 0, bits: int = 0):
    """FASTA file -> native.Batch of pairs (record 2p, record 2p + 1) plus the record ids; bits = 2 / 5: packed residues."""
    from . import native
    with open(os.fspath(path), "rb") as fh:
        r = parse_batch(fh.read(), n_threads, bits)
    if bits:
        return native.Batch(r[1], r[2], bits, r[3]), r[0]
    return native.Batch(r[1], r[2]), r[0]
