#!/usr/bin/env python3
"""Generate tests/golden/ from the read-only reference checkout.

Run ONCE in the build container (where /root/reference exists); the outputs are
committed because /root/reference does not exist on the GPU box.

What it takes from the reference (data only, no code):
  * the FASTA fixtures the reference's own hot-path tests use
    (tests/integration.rs:62-74, 234-312) -> tests/golden/fasta/{input,output}/
  * the three 26x26 substitution tables (src/alignment/score.rs:5-35, 45-75,
    82-111) parsed into tests/golden/score_tables.json
  * the doctest known answers (src/alignment/aligner.rs:75-82, 141-148,
    206-214, 281-288, 342-349; src/analysis/seq.rs:100-103), transcribed into
    tests/golden/kat.json together with the integration-test parameters.
"""
import json
import os
import re
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")

FASTA = [
    "input/global_alignment.fasta", "output/global_alignment.fasta",
    "input/local_alignment.fasta", "output/local_alignment.fasta",
    "input/semiglobal_alignment.fasta", "output/semiglobal_alignment.fasta",
    "input/fitting_alignment.fasta", "output/fitting_alignment.fasta",
    "input/overlap_alignment.fasta", "output/overlap_alignment.fasta",
    "input/edit_distance.fasta",
    "input/hamming_distance.fasta",
]


def parse_tables():
    src = open(os.path.join(REF, "src/alignment/score.rs")).read()
    blocks = re.findall(r"from_shape_vec\(\(26, 26\), vec!\[(.*?)\]\)", src, re.S)
    assert len(blocks) == 3
    names = ["blosum62", "pam250", "unit"]
    out = {}
    for name, blk in zip(names, blocks):
        blk = re.sub(r"/\*.*?\*/", "", blk, flags=re.S)
        nums = [int(x) for x in re.findall(r"-?\d+", blk)]
        assert len(nums) == 676, (name, len(nums))
        out[name] = [nums[r * 26:(r + 1) * 26] for r in range(26)]
    return out


def main():
    for rel in FASTA:
        dst = os.path.join(OUT, "fasta", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, "tests/data", rel), dst)
    with open(os.path.join(OUT, "score_tables.json"), "w") as f:
        json.dump(parse_tables(), f, separators=(",", ":"))
    kat = {
        "_source": "reference doctests + tests/integration.rs parameters (see tools/make_golden.py)",
        "doctests": [
            {"ref": "src/alignment/aligner.rs:75-82", "mode": "global", "scorer": "blosum62",
             "a": -11, "b": -1, "s1": "PRTEINS", "s2": "PRTWPSEIN",
             "score": 8, "a_align": "PRT---EINS", "b_align": "PRTWPSEIN-"},
            {"ref": "src/alignment/aligner.rs:141-148", "mode": "local", "scorer": "blosum62",
             "a": -11, "b": -1, "s1": "PLEASANTLY", "s2": "MEANLY",
             "score": 12, "a_align": "LEAS", "b_align": "MEAN"},
            {"ref": "src/alignment/aligner.rs:206-214", "mode": "fitting", "scorer": "unit",
             "a": -1, "b": -1,
             "s1": "GCAAACCATAAGCCCTACGTGCCGCCTGTTTAAACTCGCGAACTGAAT"
                   "CTTCTGCTTCACGGTGAAAGTACCACAATGGTATCACACCCCAAGGAAAC",
             "s2": "GCCGTCAGGCTGGTGTCCG",
             "score": 5, "a_align": "GCCCT-A--C-G-TG-CCG", "b_align": "GCCGTCAGGCTGGTGTCCG"},
            {"ref": "src/alignment/aligner.rs:281-288", "mode": "overlap", "scorer": "unit",
             "a": -2, "b": -2, "s1": "CTAAGGGATTCCGGTAATTAGACAG", "s2": "ATAGACCATATGTCAGTGACTGTGTAA",
             "score": 2, "a_align": "ATTAGAC-AG", "b_align": "AT-AGACCAT"},
            {"ref": "src/alignment/aligner.rs:342-349", "mode": "semiglobal", "scorer": "unit",
             "a": -1, "b": -1, "s1": "TAGCACTTGGATTCTCGG", "s2": "CAGCGTGG",
             "score": 4, "a_align": "TAGCA-CTTGGATTCTCGG", "b_align": "---CAGCGTGG--------"},
        ],
        "edit_distance_doctest": {"ref": "src/analysis/seq.rs:100-103",
                                  "s1": "ACTGGATTC", "s2": "ACGT", "distance": 5},
        "integration": [
            {"ref": "tests/integration.rs:234-248", "mode": "global", "scorer": "blosum62",
             "a": -11, "b": -1, "fixture": "global_alignment", "score": 232},
            {"ref": "tests/integration.rs:250-264", "mode": "local", "scorer": "blosum62",
             "a": -11, "b": -1, "fixture": "local_alignment", "score": 20431},
            {"ref": "tests/integration.rs:266-280", "mode": "fitting", "scorer": "unit",
             "a": -1, "b": -1, "fixture": "fitting_alignment", "score": 145},
            {"ref": "tests/integration.rs:282-296", "mode": "overlap", "scorer": "unit",
             "a": -2, "b": -2, "fixture": "overlap_alignment", "score": 698},
            {"ref": "tests/integration.rs:298-312", "mode": "semiglobal", "scorer": "unit",
             "a": -1, "b": -1, "fixture": "semiglobal_alignment", "score": 982},
        ],
        "edit_distance_integration": {"ref": "tests/integration.rs:69-74",
                                      "fixture": "edit_distance", "distance": 299},
        "hamming_distance_doctest": {"ref": "src/analysis/seq.rs:64-73",
                                     "s1": "GAGCCTACTAACGGGAT", "s2": "CATCGTAATGACGGCCT", "distance": 7},
        "hamming_distance_integration": {"ref": "tests/integration.rs:62-67",
                                         "fixture": "hamming_distance", "distance": 477},
    }
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("golden fixtures written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
