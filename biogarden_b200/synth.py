"""Synthetic workloads of BASELINE.json's configs: thin names over native.synth_pairs."""
from . import native

DNA = b"ACGT"
PROTEIN = b"ACDEFGHIKLMNPQRSTVWY"

# name -> (seed, alphabet, len_lo, len_hi, resize_b, mode, scorer, a, b, full-size pair count)
CONFIGS = {
    # BASELINE config #1 is not synthetic: the reference's own fixture (examples/from_file.rs:20-31), one pair
    "cfg1_from_file": dict(seed=0, alphabet=DNA, lo=9559, hi=8457, resize_b=False, mode="semiglobal",
                           scorer="blosum62", a=-1, b=-2, n_pairs=1, fixture="semiglobal_alignment"),
    "cfg2_dna150_global": dict(seed=2, alphabet=DNA, lo=150, hi=150, resize_b=True, mode="global",
                               scorer="unit", a=-2, b=-1, n_pairs=1_000_000),
    "cfg3_edit_100_300": dict(seed=3, alphabet=DNA, lo=100, hi=300, resize_b=True, mode="edit",
                              scorer=None, a=0, b=0, n_pairs=10_000_000),
    "cfg4_protein_local": dict(seed=4, alphabet=PROTEIN, lo=200, hi=1000, resize_b=False, mode="local",
                               scorer="blosum62", a=-11, b=-1, n_pairs=100_000),
    "cfg5_long_semiglobal": dict(seed=5, alphabet=DNA, lo=50_000, hi=100_000, resize_b=False, mode="semiglobal",
                                 scorer="unit", a=-1, b=-1, n_pairs=1_000),
}


def make(name: str, n_pairs=None, first_pair=0) -> native.Batch:
    c = CONFIGS[name]
    return native.synth_pairs(c["seed"], first_pair, c["n_pairs"] if n_pairs is None else n_pairs, c["alphabet"],
                              c["lo"], c["hi"], c["resize_b"])
