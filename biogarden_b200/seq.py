"""analysis::seq::edit_distance (reference src/analysis/seq.rs:105-130) and hamming_distance (seq.rs:74-83),
single-pair mirrors with the reference's names and error behaviour, and their batched forms."""
from typing import List

from . import native
from .error import InvalidInputSize
from .sequence import Sequence, Tile

_ctx = None


def _context() -> native.Context:
    global _ctx
    if _ctx is None:
        _ctx = native.Context()
    return _ctx


def edit_distance(seq1, seq2) -> int:
    """Levenshtein distance on raw bytes; always succeeds (the reference returns Ok(..) always)."""
    batch = native.Batch.from_sequences([bytes(Sequence(seq1)), bytes(Sequence(seq2))])
    return int(_context().edit_distance_batch(batch)[0])


def edit_distance_batch(pairs: Tile, ctx: native.Context = None) -> List[int]:
    """Distance of (pairs[2p], pairs[2p+1]) for every p; odd Tile length -> InvalidInputSize."""
    if len(pairs) % 2:
        raise InvalidInputSize()
    batch = native.Batch.from_sequences([bytes(s) for s in pairs])
    return [int(x) for x in (ctx or _context()).edit_distance_batch(batch)]


def hamming_distance(seq1, seq2) -> int:
    """Number of positions at which the two sequences differ; Err(InvalidInputSize) -- here the exception --
    when the lengths differ (seq.rs:74-83)."""
    s1, s2 = bytes(Sequence(seq1)), bytes(Sequence(seq2))
    if len(s1) != len(s2):
        raise InvalidInputSize()
    batch = native.Batch.from_sequences([s1, s2])
    return int(_context().hamming_distance_batch(batch)[0])


def hamming_distance_batch(pairs: Tile, ctx: native.Context = None) -> List[int]:
    """Distance of (pairs[2p], pairs[2p+1]) for every p; odd Tile length or a pair of unequal lengths ->
    InvalidInputSize."""
    if len(pairs) % 2:
        raise InvalidInputSize()
    batch = native.Batch.from_sequences([bytes(s) for s in pairs])
    return [int(x) for x in (ctx or _context()).hamming_distance_batch(batch)]
