"""Pin the oracle against every known-answer test the reference holds for the path
(SURVEY 4 / 8c): 5 integration goldens with full strings, 5 doctests, edit_distance x2.
Both oracle forms (literal, lean) must reproduce them bit-exactly."""
import os

import pytest

import _oracle as orc
from biogarden_b200.fasta import read_tile


def _fixture(golden_dir, name):
    inp = read_tile(os.path.join(golden_dir, "fasta", "input", name + ".fasta"))
    outp = os.path.join(golden_dir, "fasta", "output", name + ".fasta")
    out = read_tile(outp) if os.path.exists(outp) else None
    return inp, out


@pytest.mark.parametrize("lean", [False, True])
def test_doctests(kat, lean):
    for t in kat["doctests"]:
        st, sc, a, b = orc.align(t["mode"], t["s1"].encode(), t["s2"].encode(), t["scorer"], t["a"], t["b"], lean=lean)
        assert st == orc.OK, t["ref"]
        assert sc == t["score"], t["ref"]
        assert a == t["a_align"].encode(), t["ref"]
        assert b == t["b_align"].encode(), t["ref"]


@pytest.mark.parametrize("lean", [False, True])
@pytest.mark.parametrize("idx", range(5))
def test_integration_goldens(kat, golden_dir, idx, lean):
    t = kat["integration"][idx]
    inp, out = _fixture(golden_dir, t["fixture"])
    st, sc, a, b = orc.align(t["mode"], bytes(inp[0]), bytes(inp[1]), t["scorer"], t["a"], t["b"], lean=lean)
    assert st == orc.OK
    assert sc == t["score"]
    assert a == bytes(out[0])
    assert b == bytes(out[1])


@pytest.mark.parametrize("lean", [False, True])
def test_edit_distance(kat, golden_dir, lean):
    d = kat["edit_distance_doctest"]
    assert orc.edit_distance(d["s1"].encode(), d["s2"].encode(), lean=lean) == d["distance"]
    t = kat["edit_distance_integration"]
    inp, _ = _fixture(golden_dir, t["fixture"])
    assert orc.edit_distance(bytes(inp[0]), bytes(inp[1]), lean=lean) == t["distance"]


def test_config1_derived_value(golden_dir):
    """BASELINE config #1 (examples/from_file.rs:20-31: blosum62, open -1, enlarge -2) has no
    reference golden.  SURVEY 4 records a survey-time restatement's score (31188) and aligned
    length (11242); both oracle forms reproduce those.  The survey's FNV digests could not be
    reproduced (its hashing recipe is not recorded), so the digests pinned here are this
    repo's own -- literal and lean form agree -- and are labelled derived, not reference."""
    inp, _ = _fixture(golden_dir, "semiglobal_alignment")
    res = [orc.align("semiglobal", bytes(inp[0]), bytes(inp[1]), "blosum62", -1, -2, lean=l) for l in (False, True)]
    assert res[0] == res[1]
    st, sc, a, b = res[0]
    assert st == orc.OK and sc == 31188 and len(a) == 11242

    def fnv(s):
        h = 14695981039346656037
        for c in s:
            h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h
    assert "%016x" % fnv(a) == "48763f4161d42930"
    assert "%016x" % fnv(b) == "600db5c78f5408a0"


def test_hamming_distance(kat, golden_dir):
    """seq.rs:64-73 doctest (7) and tests/integration.rs:62-67 (477); unequal lengths are Err(InvalidInputSize)."""
    d = kat["hamming_distance_doctest"]
    assert orc.hamming_distance(d["s1"].encode(), d["s2"].encode()) == (orc.OK, d["distance"])
    t = kat["hamming_distance_integration"]
    inp, _ = _fixture(golden_dir, t["fixture"])
    assert orc.hamming_distance(bytes(inp[0]), bytes(inp[1])) == (orc.OK, t["distance"])
    assert orc.hamming_distance(b"ACGT", b"ACG")[0] == 2          # ORC_ERR_SIZE
    assert orc.hamming_distance(b"", b"") == (orc.OK, 0)


def test_p_distance_matrix_restatement():
    """stat.rs:138-152 has no reference test; the restatement is checked against the definition written out with
    numpy (f32 division, zip to the shorter row, columns = len(row 0))."""
    import numpy as np
    rows = [b"ACGTACGTAC", b"ACGTTCGTAA", b"TTTTTTTTTT", b"ACGTACG", b""]
    got = orc.p_distance_matrix(rows)
    want = np.zeros((5, 5), np.float32)
    for i, a in enumerate(rows):
        for j, b in enumerate(rows):
            c = sum(1 for x, y in zip(a, b) if x != y) if i != j else 0
            want[i, j] = np.float32(c) / np.float32(len(rows[0]))
    assert got.dtype == np.float32 and np.array_equal(got, want)
    assert got[0, 1] == np.float32(2) / np.float32(10) and got[0, 3] == 0 and got[2, 3] == np.float32(0.6)
