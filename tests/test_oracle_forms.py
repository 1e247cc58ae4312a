"""The two oracle forms must agree: literal (reference layout, Rust run-time checks restated) vs
lean (rolling rows + 4-bit trace, SURVEY A.6 rules restated) on random inputs incl. empty
sequences, tie-heavy alphabets, positive penalties and the 1023/1024/1025 buffer edge."""
import random

import _oracle as orc


def _cls(st):
    return "U" if st in orc.UNDEFINED else st


def test_literal_equals_lean_random():
    rng = random.Random(1)
    modes = list(orc.MODES)
    n_undefined = 0
    for it in range(2500):
        mode = rng.choice(modes)
        alpha = rng.choice([b"ACGT", b"ACDEFGHIKLMNPQRSTVWY", b"AC"])
        if it % 60 == 0:
            n = rng.choice([0, 1, 2, 1022, 1023, 1024, 1025]); m = rng.choice([0, 1, 5, 1023, 1024, 1025, 30])
        else:
            n = rng.randint(0, 24); m = rng.randint(0, 24)
        s1 = bytes(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.6 and n > 0:
            s2 = bytearray()
            for c in s1:
                r = rng.random()
                if r < 0.1:
                    s2.append(rng.choice(alpha))
                elif r < 0.15:
                    continue
                elif r < 0.2:
                    s2.append(c); s2.append(rng.choice(alpha))
                else:
                    s2.append(c)
            s2 = bytes(s2)
        else:
            s2 = bytes(rng.choice(alpha) for _ in range(m))
        scorer = rng.choice(["blosum62", "unit", "pam250"])
        a, b = rng.choice([(-11, -1), (-1, -1), (-2, -1), (-1, -2), (0, 0), (-5, -5), (-3, 0), (1, -1), (-2, -2)])
        lit = orc.align(mode, s1, s2, scorer, a, b)
        lean = orc.align(mode, s1, s2, scorer, a, b, lean=True)
        assert _cls(lit[0]) == _cls(lean[0]), (mode, len(s1), len(s2), scorer, a, b, lit[0], lean[0])
        if lit[0] == orc.OK:
            assert lit == lean, (mode, s1, s2, scorer, a, b)
        n_undefined += _cls(lit[0]) == "U"
    assert n_undefined > 20   # the sweep does reach the reference-undefined domain


def test_edit_distance_forms_agree():
    rng = random.Random(2)
    for _ in range(300):
        s1 = bytes(rng.randrange(256) for _ in range(rng.randint(0, 60)))
        s2 = bytes(rng.randrange(256) for _ in range(rng.randint(0, 60)))
        assert orc.edit_distance(s1, s2) == orc.edit_distance(s1, s2, lean=True)
    assert orc.edit_distance(b"", b"") == 0 and orc.edit_distance(b"", b"ACG") == 3


def test_reused_aligner_equals_fresh_inside_domain():
    """Buffer contents persist across calls (aligner.rs:30-38) but never leak into results."""
    L = orc.lib()
    h = L.orc_aligner_new()
    rng = random.Random(3)
    try:
        for _ in range(200):
            mode = rng.choice(list(orc.MODES))
            s1 = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 40)))
            s2 = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 40)))
            if mode == "fitting" and len(s1) < len(s2):
                s1, s2 = s2, s1
            fresh = orc.align(mode, s1, s2, "unit", -2, -1)
            if fresh[0] != orc.OK:
                continue
            assert orc.align(mode, s1, s2, "unit", -2, -1, aligner=h) == fresh
    finally:
        L.orc_aligner_free(h)
