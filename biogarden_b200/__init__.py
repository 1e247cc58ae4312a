"""biogarden_b200 -- B200-native batched pairwise alignment behind biogarden's SequenceAligner API.

Host-side mirror of the reference interface for the alignment hot path only
(reference src/alignment/aligner.rs, src/alignment/score.rs, src/analysis/seq.rs::edit_distance).
All arithmetic runs in hand-written CUDA (csrc/) reached through the C ABI of include/bgalign.h;
there is no CPU fallback: importing the compute entry points without the built library fails loudly.
"""
from .error import BioError, InvalidArgumentRange, InvalidInputSize, ReferenceUndefined, EngineError  # noqa: F401
from .sequence import Sequence, Tile  # noqa: F401
from . import fasta  # noqa: F401

__all__ = ["Sequence", "Tile", "fasta", "BioError", "InvalidArgumentRange", "InvalidInputSize",
           "ReferenceUndefined", "EngineError"]


def __getattr__(name):
    # compute-facing modules load the native library on first use
    if name in ("aligner", "score", "seq", "stat", "native", "synth"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
