"""Substitution scorers passed to the aligner (reference src/alignment/score.rs).

`blosum62`, `pam250`, `unit` mirror score.rs:38-41, 78-80, 114-116: callables over two residue
BYTES indexing a 26x26 table by `byte - 65`; like the reference they fail for bytes outside
'A'..='Z' (there a panic, here ReferenceUndefined).  Any Python callable `(a: int, b: int) -> int`
is accepted by the aligner as well (the `&dyn Fn(&u8,&u8)->i32` of aligner.rs:85).

The callback never runs on the GPU: `materialise` evaluates it once per (seq1 residue, seq2
residue) combination present in the batch and the dense table goes to the device."""
import numpy as np

from .error import ReferenceUndefined


class TableScorer:
    """A shipped 26x26 scorer."""

    def __init__(self, name):
        self.name = name
        self._table = None

    @property
    def table(self) -> np.ndarray:
        if self._table is None:
            from . import native
            self._table = native.score_table26(self.name)
        return self._table

    def __call__(self, a: int, b: int) -> int:
        ia, ib = int(a) - 65, int(b) - 65
        if not (0 <= ia < 26 and 0 <= ib < 26):
            raise ReferenceUndefined("score::%s indexes out of bounds for bytes (%d, %d) (score.rs:40)" % (self.name, a, b))
        return int(self.table[ia, ib])

    def __repr__(self):
        return "score." + self.name


blosum62 = TableScorer("blosum62")
pam250 = TableScorer("pam250")
unit = TableScorer("unit")


def match_mismatch(match: int, mismatch: int):
    """The "DNA match/mismatch" closure north_star names: |a, b| if a == b { match } else { mismatch }."""
    def f(a, b):
        return match if a == b else mismatch
    f.match_mismatch = (int(match), int(mismatch))
    return f


def materialise(score, hist_a: np.ndarray, hist_b: np.ndarray):
    """(table[int32 n_rows x n_cols], row_code[256], col_code[256]) for the residues present.
    Rows = bytes occurring in any seq1, columns = bytes occurring in any seq2; 0xFF elsewhere."""
    rows = np.nonzero(hist_a)[0]
    cols = np.nonzero(hist_b)[0]
    row_code = np.full(256, 0xFF, np.uint8)
    col_code = np.full(256, 0xFF, np.uint8)
    if len(rows) > 255 or len(cols) > 255:
        raise ValueError("more than 255 distinct residues on one side")
    row_code[rows] = np.arange(len(rows), dtype=np.uint8)
    col_code[cols] = np.arange(len(cols), dtype=np.uint8)
    table = np.zeros((max(1, len(rows)), max(1, len(cols))), np.int32)
    if isinstance(score, TableScorer):
        if len(rows) and len(cols):
            bad = [int(x) for x in list(rows) + list(cols) if not 65 <= x <= 90]
            if bad:
                raise ReferenceUndefined("score::%s panics for byte %d (score.rs:40)" % (score.name, bad[0]))
            table[:len(rows), :len(cols)] = score.table[np.ix_(rows - 65, cols - 65)]
    elif getattr(score, "match_mismatch", None) is not None:
        mt, mm = score.match_mismatch
        if len(rows) and len(cols):
            table[:len(rows), :len(cols)] = np.where(rows[:, None] == cols[None, :], mt, mm)
    else:
        for i, x in enumerate(rows):
            for j, y in enumerate(cols):
                table[i, j] = int(score(int(x), int(y)))
    return table, row_code, col_code
