"""Native FASTA ingest (csrc/fasta_ingest.cpp) against the Python restatement of the reference
reader (biogarden_b200/fasta.py <- src/io/fasta.rs:95-136).  Host code only: runs without a GPU."""
import os
import random

import numpy as np
import pytest

from biogarden_b200 import fasta


def _via_reader(text: bytes):
    t = fasta.Tile()
    fasta.Reader.from_string(text.decode("ascii")).read_all(t)
    return [(s.id, bytes(s)) for s in t]


def _via_native(text: bytes, threads=0):
    ids, res, off = fasta.parse_batch(text, threads)
    return [(ids[r], bytes(res[int(off[r]):int(off[r + 1])])) for r in range(len(ids))]


CASES = [
    b"",
    b">a\nACGT\n",
    b">a desc here\nACGT\nTTGA\n>b\nGG",                       # description dropped, no trailing newline
    b">a\r\nAC GT \r\n\r\n  TT\t\n>b\tx\r\nGG\r\n",            # CRLF, inner / leading blanks kept, trailing trimmed
    b">a\n\n\n>b\nAC\n",                                       # record with an id and no residues is kept
    b">a\nAC\n>\n>c\nGG\n",                                    # empty record ends read_all: c is never seen
    b">a\nAC\n>   \nTT\n>c\nGG\n",                             # empty id but residues: not an empty record
    b"> lead\nAC\n",                                           # header starting with whitespace: id '' + description
    b">only_header",
    b">a\nAC\n>b",
]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_grammar_cases(idx):
    text = CASES[idx]
    assert _via_native(text) == _via_reader(text)
    assert _via_native(text, 1) == _via_reader(text)


def test_error_on_missing_marker():
    for text in (b"ACGT\n>a\nAC\n", b"\n>a\nAC\n", b" >a\nAC\n"):
        with pytest.raises(IOError):
            _via_native(text)
        with pytest.raises(IOError):
            _via_reader(text)


def test_reference_fixtures(golden_dir):
    d = os.path.join(golden_dir, "fasta", "input")
    for name in sorted(os.listdir(d)):
        text = open(os.path.join(d, name), "rb").read()
        assert _via_native(text) == _via_reader(text), name
        batch, ids = fasta.read_batch(os.path.join(d, name))
        assert batch.n_pairs == len(ids) // 2


def test_random_texts_all_thread_counts():
    """Random record soups (blank lines, CR, trailing blanks, '>' inside lines, empty records) large enough that the
    parser cuts them into several segments; every thread count gives the reader's result."""
    rng = random.Random(3)
    for trial in range(6):
        parts = []
        for r in range(rng.randint(1, 4000)):
            hdr = b">" + bytes(rng.choice(b"abcxyz_01") for _ in range(rng.randint(0 if rng.random() < 0.002 else 1, 12)))
            if rng.random() < 0.3:
                hdr += rng.choice([b" ", b"\t"]) + b"some description"
            parts.append(hdr + rng.choice([b"\n", b"\r\n", b"  \n"]))
            for _ in range(rng.randint(0, 12)):
                line = bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(0, 120)))
                if rng.random() < 0.05:
                    line = line[:3] + b">" + line[3:]          # '>' that is not at a line start
                parts.append(line + rng.choice([b"\n", b"\r\n", b" \t\n"]))
        text = b"".join(parts)
        if trial % 2:
            text = text.rstrip(b"\r\n \t")
        want = _via_reader(text)
        for threads in (1, 2, 3, 7, 0):
            assert _via_native(text * (1 if trial else 8), threads) == (want if trial else _via_reader(text * 8)), (trial, threads)


def test_batch_layout_feeds_the_abi(golden_dir):
    batch, ids = fasta.read_batch(os.path.join(golden_dir, "fasta", "input", "global_alignment.fasta"))
    assert batch.n_pairs == 1 and len(ids) == 2
    n, m = batch.lengths()
    t = fasta.read_tile(os.path.join(golden_dir, "fasta", "input", "global_alignment.fasta"))
    assert (int(n[0]), int(m[0])) == (len(t[0]), len(t[1]))
    assert bytes(batch.residues[:int(n[0])]) == bytes(t[0])
