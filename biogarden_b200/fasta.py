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
        header = self._line[1:].rstrip().split(None, 1)
        # splitn(2, whitespace) on a trimmed header: first field always exists (possibly empty)
        record.id = header[0] if header else ""
        record.desc = header[1] if len(header) > 1 else None
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
