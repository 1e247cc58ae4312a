"""analysis::stat::p_distance_matrix (reference src/analysis/stat.rs:138-152)."""
import numpy as np

from . import native
from .sequence import Tile
from . import seq as _seq


def p_distance_matrix(matrix: Tile, ctx: native.Context = None) -> np.ndarray:
    """rows x rows float32 matrix of p-distances: mismatches over the zip of two rows, as f32, divided by the
    length of row 0 as f32; zeros on the diagonal.  An empty Tile makes the reference panic (tile.rs:32):
    IndexError here."""
    rows = [bytes(s) for s in matrix]
    if not rows:
        raise IndexError("p_distance_matrix of an empty Tile (the reference panics on data[0])")
    res = np.frombuffer(b"".join(rows) or b"\0", dtype=np.uint8)
    off = np.zeros(len(rows) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in rows])
    return (ctx or _seq._context()).p_distance_matrix(res, off)
