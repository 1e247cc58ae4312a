"""Error convention at the boundary: mirrors BioError (reference src/error.rs:8-14)."""


class BioError(Exception):
    """Base of the reference's BioError variants that the alignment path can return."""


class InvalidInputSize(BioError):
    """BioError::InvalidInputSize (error.rs:9) -- e.g. fitting_alignment with len1 < len2 (aligner.rs:223-225)."""

    def __str__(self):
        return "Provided inputs have invalid size!"


class InvalidArgumentRange(BioError):
    """BioError::InvalidArgumentRange (error.rs:10) -- positive gap penalties (aligner.rs:87-89,153-155,219-221)."""

    def __str__(self):
        return "The provided has is within an unsupported range!"


class ReferenceUndefined(BioError):
    """Not a reference variant: the reference would panic or loop forever on this input
    (SURVEY Appendix A.6).  Raised by the single-pair drop-in methods; the batched entry
    point reports it per pair in `status` instead."""


class EngineError(RuntimeError):
    """CUDA / allocation failure inside the native library (BG_ECUDA, BG_ENOMEM, ...)."""
