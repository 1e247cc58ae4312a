"""SequenceAligner: host-side mirror of reference src/alignment/aligner.rs:28-609.

Same five public methods, same argument meaning, same (score, a_align, b_align) result and
the same error behaviour (BioError variants under the reference's conditions), plus the new
batched entry point that takes a Tile of pairs.  Every method is a thin shim over the C ABI
(include/bgalign.h); the DP, end-cell selection and traceback all run in CUDA."""
from typing import List, Tuple

import numpy as np

from . import native, score as score_mod
from .error import InvalidArgumentRange, InvalidInputSize, ReferenceUndefined
from .sequence import Sequence, Tile

MODES = ("global", "local", "semiglobal", "fitting", "overlap")


class SequenceAligner:
    def __init__(self, devices=None):
        """SequenceAligner::new (aligner.rs:44-55).  `devices`: CUDA ordinals to shard batches over."""
        self._ctx = native.Context(devices)

    # ---- the reference's public methods -------------------------------------------------
    def global_alignment(self, seq1, seq2, score, a: int, b: int):
        """aligner.rs:84-121"""
        return self._single("global", seq1, seq2, score, a, b)

    def local_alignment(self, seq1, seq2, score, a: int, b: int):
        """aligner.rs:150-185"""
        return self._single("local", seq1, seq2, score, a, b)

    def fitting_alignment(self, seq1, seq2, score, a: int, b: int):
        """aligner.rs:216-260"""
        return self._single("fitting", seq1, seq2, score, a, b)

    def overlap_alignment(self, seq1, seq2, score, a: int, b: int):
        """aligner.rs:290-321"""
        return self._single("overlap", seq1, seq2, score, a, b)

    def semiglobal_alignment(self, seq1, seq2, score, a: int, b: int):
        """aligner.rs:351-435"""
        return self._single("semiglobal", seq1, seq2, score, a, b)

    # ---- new: batched entry point ------------------------------------------------------
    def align_batch(self, pairs: Tile, mode: str, score, a: int, b: int,
                    with_status: bool = False) -> List[Tuple[int, Sequence, Sequence]]:
        """Aligns pair p = (pairs[2p], pairs[2p+1]) for every p.  Odd Tile length ->
        InvalidInputSize.  Pairs on which the reference itself is undefined are
        returned with the engine's extension; pass with_status=True to get (results, status)."""
        if len(pairs) % 2:
            raise InvalidInputSize()
        batch = native.Batch.from_sequences([bytes(s) for s in pairs])
        res = self.align_batch_raw(batch, mode, score, a, b)
        try:
            out = []
            for p in range(batch.n_pairs):
                x, y = res.strings(p)
                out.append((int(res.score[p]), Sequence(x), Sequence(y)))
            status = res.status.copy()
        finally:
            res.close()
        return (out, status) if with_status else out

    def align_batch_raw(self, batch: native.Batch, mode: str, score, a: int, b: int,
                        score_only: bool = False) -> native.Result:
        """Batch in, native.Result (numpy views over the pinned result arrays) out."""
        params = self.make_params(batch, mode, score, a, b, score_only)
        return self._ctx.align_batch(batch, params)

    def make_params(self, batch: native.Batch, mode: str, score, a: int, b: int, score_only=False) -> native.Params:
        if mode not in MODES:
            raise ValueError("unknown alignment mode %r" % (mode,))
        # reference order of checks: sign of the penalties first (aligner.rs:87-89,153-155,219-221),
        # then the fitting size check (aligner.rs:223-225); the score callback is never called before.
        if mode in ("global", "local", "fitting") and (a > 0 or b > 0):
            raise InvalidArgumentRange()
        if mode == "fitting":
            n, m = batch.lengths()
            if np.any(n < m):
                raise InvalidInputSize()
        ha, hb = batch.histograms()
        table, rc, cc = score_mod.materialise(score, ha, hb)
        return native.Params(mode, a, b, table, rc, cc, score_only)

    @property
    def context(self) -> native.Context:
        return self._ctx

    # ---- internals -----------------------------------------------------------------------
    def _single(self, mode, seq1, seq2, score, a, b):
        s1, s2 = bytes(Sequence(seq1)), bytes(Sequence(seq2))
        batch = native.Batch.from_sequences([s1, s2])
        res = self.align_batch_raw(batch, mode, score, a, b)
        try:
            if res.status[0] != native.ST_OK:
                raise ReferenceUndefined("the reference panics or never returns on this input "
                                         "(mode=%s, len1=%d, len2=%d; SURVEY A.6)" % (mode, len(s1), len(s2)))
            x, y = res.strings(0)
            return int(res.score[0]), Sequence(x), Sequence(y)
        finally:
            res.close()
