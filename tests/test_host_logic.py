"""CPU-side tests: the C-ABI library loads and exports every declared symbol, refuses to run
without a device (no CPU fallback), and the host logic (FASTA ingest, synthetic workloads,
callback materialisation, error mapping) behaves like the reference's."""
import os
import re

import numpy as np
import pytest

from biogarden_b200 import native, score
from biogarden_b200.error import EngineError, ReferenceUndefined
from biogarden_b200.fasta import Reader, Record, read_tile
from biogarden_b200.sequence import Sequence, Tile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bgalign.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = native.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(native.SYMBOLS) == declared
    assert L.bg_version() == 2


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_device():
    with pytest.raises(EngineError):
        native.Context()
    from biogarden_b200.aligner import SequenceAligner
    with pytest.raises(EngineError):
        SequenceAligner()


def test_score_tables_match_reference_golden():
    import json
    tabs = json.load(open(os.path.join(ROOT, "tests", "golden", "score_tables.json")))
    for name in ("blosum62", "pam250", "unit"):
        assert np.array_equal(native.score_table26(name), np.array(tabs[name], np.int32)), name
    assert score.blosum62(ord("W"), ord("W")) == 11
    assert score.unit(ord("A"), ord("C")) == -1
    with pytest.raises(ReferenceUndefined):
        score.blosum62(ord("-"), ord("A"))


def test_materialise_respects_argument_order_and_presence():
    calls = []

    def asym(a, b):
        calls.append((a, b))
        return 10 * a - b
    ha = np.zeros(256, np.uint64); hb = np.zeros(256, np.uint64)
    ha[[65, 67]] = 1; hb[[71, 84, 65]] = 1
    table, rc, cc = score.materialise(asym, ha, hb)
    assert table.shape == (2, 3)
    assert set(calls) == {(x, y) for x in (65, 67) for y in (65, 71, 84)}   # never on absent bytes
    assert table[rc[67], cc[84]] == 10 * 67 - 84
    assert rc[71] == 0xFF and cc[67] == 0xFF
    t2, rc2, cc2 = score.materialise(score.match_mismatch(5, -4), ha, hb)
    assert t2[rc2[65], cc2[65]] == 5 and t2[rc2[67], cc2[65]] == -4
    hb[45] = 1
    with pytest.raises(ReferenceUndefined):
        score.materialise(score.unit, ha, hb)


def test_fasta_reader_matches_reference_grammar(tmp_path):
    text = ">id1 some description\nACGT\nAC  \n>id2\n\nGG\n>id3\n"
    r = Reader.from_string(text)
    rec = Record()
    r.read(rec); assert (rec.id, rec.desc, rec.seq) == ("id1", "some description", "ACGTAC")
    r.read(rec); assert (rec.id, rec.desc, rec.seq) == ("id2", None, "GG")
    r.read(rec); assert (rec.id, rec.desc, rec.seq) == ("id3", None, "")
    r.read(rec); assert rec.is_empty()
    with pytest.raises(IOError):
        Reader.from_string("ACGT\n").read(rec)
    p = tmp_path / "x.fasta"
    p.write_text(">a\nAC\n>b\nGT\n")
    t = read_tile(p)
    assert len(t) == 2 and t[0] == Sequence("AC") and t[1].id == "b"
    t2 = read_tile(os.path.join(ROOT, "tests/golden/fasta/input/semiglobal_alignment.fasta"))
    assert (len(t2[0]), len(t2[1])) == (9559, 8457)


def test_sequence_semantics():
    a = Sequence("ACGT", id="x"); b = Sequence(b"ACGT")
    assert a == b and hash(a) == hash(b)      # id ignored (sequence.rs:104-117)
    a.push(ord("A")); assert len(a) == 5 and a.back() == 65 and a.pop() == 65
    a.reverse(); assert bytes(a) == b"TGCA"
    t = Tile([a, b]); assert t.size() == (2, 4) and t[1] == b


def test_synth_is_deterministic_and_shardable():
    full = native.synth_pairs(2, 0, 5000, b"ACGT", 150, 150, True)
    n, m = full.lengths()
    assert np.all(n == 150) and np.all(m == 150)
    part = native.synth_pairs(2, 3000, 1000, b"ACGT", 150, 150, True)
    lo, hi = int(full.seq_off[6000]), int(full.seq_off[8000])
    assert np.array_equal(full.residues[lo:hi], part.residues)
    again = native.synth_pairs(2, 0, 5000, b"ACGT", 150, 150, True)
    assert np.array_equal(full.residues, again.residues)
    v = native.synth_pairs(3, 0, 3000, b"ACGT", 100, 300, True)
    n, m = v.lengths()
    assert n.min() >= 100 and n.max() <= 300 and m.min() >= 100 and m.max() <= 300
    # 90 % of the pairs are mutated copies: mean identity of the aligned prefix is high
    same = np.mean(full.residues[:150] == full.residues[150:300])
    assert 0.0 <= same <= 1.0
    ha, hb = full.histograms()
    assert set(np.nonzero(ha)[0]) == {65, 67, 71, 84} and int(ha.sum()) == 5000 * 150 and int(hb.sum()) == 5000 * 150


def test_batch_layout_and_odd_tile():
    from biogarden_b200.error import InvalidInputSize
    b = native.Batch.from_sequences([b"ACG", b"", b"T", b"GGGG"])
    assert b.n_pairs == 2 and list(b.seq_off) == [0, 3, 3, 4, 8] and b.cells() == 4
    with pytest.raises(InvalidInputSize):
        native.Batch.from_sequences([b"A", b"C", b"G"])


def test_bounded_memory_plan_layout():
    """The launch plan of long pairs under a trace budget (host logic of the bounded-memory traceback, no GPU):
    with room for everything nothing is checkpointed; under a small budget pairs are grouped and cut into row
    blocks of >= 64 rows whose traces fit the budget, every pair's blocks tile its rows exactly once bottom-up, the
    checkpoint regions do not overlap; pairs that fit whole stay on the unbounded path."""
    import ctypes as C
    from biogarden_b200 import native
    L = native.lib()
    L.bg_debug_plan_long.restype = C.c_int
    L.bg_debug_plan_long.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]

    def plan(lens, budget):
        a = np.array(lens, np.uint64).reshape(-1)
        out = np.zeros(5, np.uint64)
        assert L.bg_debug_plan_long(a.ctypes.data, len(a) // 2, budget, out.ctypes.data) == 0
        return [int(x) for x in out]

    pairs = [(5000, 4200), (2100, 9000), (6100, 4500), (300, 5000), (4096, 4097), (1000, 4100), (9000, 17000)]
    cells = lambda ps: sum(n * m for n, m in ps)
    launches, ck, nb, refilled, viol = plan(pairs, 100 << 30)
    assert (ck, refilled, viol) == (0, 0, 0) and launches >= 1
    launches, ck, nb, refilled, viol = plan(pairs, 1 << 20)
    assert viol == 0 and ck >= 3 and nb >= 32 and refilled == cells(pairs) - 300 * 5000
    launches, ck, nb, refilled, viol = plan(pairs, 6 << 20)
    assert viol == 0 and ck == 1 and refilled == cells(pairs) - 1000 * 4100 - 300 * 5000
    # cfg5-like: 40 pairs of 50-100 kbp under 16 GiB
    rng = np.random.default_rng(5)
    big = [(int(n), int(m)) for n, m in rng.integers(50_000, 100_001, size=(40, 2))]
    launches, ck, nb, refilled, viol = plan(big, 16 << 30)
    # the largest pairs (fewer than six of them fit together) are checkpointed, the smallest ones fit six at a time
    assert viol == 0 and ck >= 1 and 0 < refilled < cells(big) and 2 <= nb <= 64
    launches, ck, nb, refilled, viol = plan(big, 1 << 30)
    assert viol == 0 and refilled == cells(big)
    # a budget below one 64-row block of the widest pair is refused
    a = np.array([(9000, 170000)], np.uint64).reshape(-1)
    out = np.zeros(5, np.uint64)
    assert L.bg_debug_plan_long(a.ctypes.data, 1, 1 << 20, out.ctypes.data) == native.BG_ENOMEM


def test_host_expander_matches_definition():
    """bg_expand_ops (host_expand.cpp): 2-bit ops -> aligned strings, both implementations (portable and, when the
    host has it, AVX-512 VBMI2) against the definition written out in Python; lengths around the 16-op word and the
    64-column vector width, all-gap runs, empty alignment.  Pure host code: no GPU involved."""
    import random
    rng = random.Random(3)
    impls = [0] + ([1] if native.expand_kind() == "avx512-vbmi2" else [])
    for ln in list(range(0, 70)) + [127, 128, 129, 191, 192, 193, 1000, 4097]:
        for bias in (0.05, 0.5):
            ops_list = [0 if rng.random() > bias else rng.choice((1, 2)) for _ in range(ln)]
            na = sum(1 for o in ops_list if o != 2); nb = sum(1 for o in ops_list if o != 1)
            s1 = bytes(rng.choice(b"ACGT") for _ in range(na)); s2 = bytes(rng.choice(b"ACGT") for _ in range(nb))
            words = np.zeros((ln + 15) // 16 + 1, np.uint32)
            for q, o in enumerate(ops_list):
                words[q >> 4] |= np.uint32(o << (2 * (q & 15)))
            ia = ib = 0
            wa, wb = bytearray(), bytearray()
            for o in ops_list:
                if o != 2:
                    wa.append(s1[ia]); ia += 1
                else:
                    wa.append(45)
                if o != 1:
                    wb.append(s2[ib]); ib += 1
                else:
                    wb.append(45)
            for impl in impls + [None]:
                got = native.expand_ops(s1, s2, words, ln, impl)
                assert got == (bytes(wa), bytes(wb)), (ln, bias, impl)


def test_packed_residues_roundtrip_and_fasta():
    """bg_pack_residues / bg_unpack_residues (2 bit and 5 bit per residue), bg_residue_histogram on packed batches and
    bg_fasta_parse_packed: packing is lossless at every offset, a fifth letter (2 bit) / a non-letter (5 bit) is an
    error, and the packed FASTA ingest yields the records the byte ingest yields."""
    import ctypes as C
    import random
    from biogarden_b200 import fasta
    rng = random.Random(9)
    L = native.lib()
    for bits, alpha in ((2, b"ACGT"), (2, b"AT"), (5, b"ACDEFGHIKLMNPQRSTVWY"), (5, b"ABCDEFGHIJKLMNOPQRSTUVWXYZ")):
        seqs = [bytes(rng.choice(alpha) for _ in range(rng.choice([0, 1, 2, 3, 4, 5, 7, 8, 9, 31, 150, 1001]))) for _ in range(60)]
        b = native.Batch.from_sequences(seqs)
        pk = b.pack(bits)
        assert pk.packing == bits and np.array_equal(pk.unpacked_residues(), b.residues)
        # any sub-range
        total = int(b.seq_off[-1])
        for _ in range(50):
            first = rng.randrange(0, total + 1); count = rng.randrange(0, total - first + 1)
            out = np.zeros(max(1, count), np.uint8)
            native.check(L.bg_unpack_residues(pk.residues.ctypes.data, bits, pk.alphabet.ctypes.data, first, count, out.ctypes.data))
            assert np.array_equal(out[:count], b.residues[first:first + count])
        ha, hb = b.histograms(); pa, pb = pk.histograms()
        assert np.array_equal(ha, pa) and np.array_equal(hb, pb)
    bad = native.Batch.from_sequences([b"ACGTN", b"ACGT"])
    with pytest.raises(EngineError):
        bad.pack(2)
    with pytest.raises(EngineError):
        native.Batch.from_sequences([b"AC-T", b"ACGT"]).pack(5)
    for name, bits in (("semiglobal_alignment", 2), ("local_alignment", 5), ("edit_distance", 5)):
        path = os.path.join(ROOT, "tests", "golden", "fasta", "input", name + ".fasta")
        plain, ids = fasta.read_batch(path)
        packed, ids2 = fasta.read_batch(path, bits=bits)
        assert ids == ids2 and np.array_equal(plain.seq_off, packed.seq_off)
        assert packed.packing == bits and np.array_equal(packed.unpacked_residues(), plain.residues)
        assert packed.residues.size < plain.residues.size


def test_batch_scan_matches_numpy():
    """The host's one pass over a batch's offsets (scan_batch: validity, length statistics, the cost of every block of
    4096 pairs, the length classes present) against the same quantities computed with numpy -- uniform read sets (runs
    of equal pairs are accounted once), mixed lengths, pairs wider than the K1 classes, and a broken offset."""
    import ctypes as C
    L = native.lib()
    L.bg_debug_scan.restype = C.c_int
    L.bg_debug_scan.argtypes = [C.POINTER(native.bg_batch), C.c_int, C.c_int, C.c_void_p]
    rng = np.random.RandomState(5)

    def mask_bit(m):
        for k, hi in enumerate((64, 96, 128, 160, 192, 256, 384, 512, 640, 768, 1024)):
            if m <= hi:
                return k
        return 11

    cases = {
        "uniform": np.full(2 * 70000, 150, np.uint64),
        "mixed": rng.randint(0, 1200, 2 * 50000).astype(np.uint64),
        "runs": np.repeat(rng.randint(1, 400, 2 * 300).astype(np.uint64).reshape(-1, 2), 177, axis=0).reshape(-1),
        "wide": np.concatenate([rng.randint(50, 300, 2 * 9000), [9000, 5000, 100, 4097]]).astype(np.uint64),
    }
    for name, lens in cases.items():
        off = np.concatenate([[7], 7 + np.cumsum(lens)]).astype(np.uint64)
        batch = native.Batch(np.zeros(1, np.uint8), off)          # the scan never touches the residues
        n, m = lens[0::2].astype(np.float64), lens[1::2].astype(np.float64)
        for with_stats in (0, 1):
            out = np.zeros(8, np.uint64)
            assert L.bg_debug_scan(C.byref(batch.c), with_stats, 1, out.ctypes.data) == 0
            assert out[0] == 1, name
            want_mask = 0
            for v in np.unique(lens[1::2]):
                want_mask |= 1 << mask_bit(int(v))
            assert int(out[1]) == want_mask, (name, hex(int(out[1])), hex(want_mask))
            assert int(out[2]) == int((lens[0::2] + lens[1::2]).max()), name
            assert int(out[3]) == int(lens[1::2].max()), name
            assert int(out[4]) == int((lens[1::2] > 4096).any()), name
            assert int(out[5]) == int((n * m + 64.0).sum()), name
    bad = np.concatenate([[0], np.cumsum(np.full(2 * 40000, 100))]).astype(np.uint64)
    bad[2 * 31234 + 1] = bad[2 * 31234] - np.uint64(1)
    out = np.zeros(8, np.uint64)
    assert L.bg_debug_scan(C.byref(native.Batch(np.zeros(1, np.uint8), bad).c), 1, 1, out.ctypes.data) == 0
    assert out[0] == 0
