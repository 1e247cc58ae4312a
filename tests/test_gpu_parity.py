"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(libbgalign.so) via the host mirror; the oracle is only the checker."""
import os
import random

import numpy as np
import pytest

import _cmp
import _oracle as orc
from biogarden_b200 import native, score as score_mod
from biogarden_b200.aligner import SequenceAligner
from biogarden_b200.fasta import read_tile
from biogarden_b200.sequence import Sequence, Tile
from biogarden_b200 import seq as seqmod
from biogarden_b200 import synth

pytestmark = pytest.mark.gpu

SHAPES = [(8, 8), (8, 12), (8, 16), (8, 19), (8, 24), (16, 10), (16, 16), (32, 5), (32, 8), (32, 12), (32, 16), (32, 20),
          (32, 24), (32, 32)]


@pytest.fixture(scope="module")
def aligner():
    return SequenceAligner()


def _fixture(golden_dir, name):
    inp = read_tile(os.path.join(golden_dir, "fasta", "input", name + ".fasta"))
    outp = os.path.join(golden_dir, "fasta", "output", name + ".fasta")
    return inp, (read_tile(outp) if os.path.exists(outp) else None)


def test_doctests(aligner, kat):
    for t in kat["doctests"]:
        f = getattr(aligner, t["mode"] + "_alignment")
        sc, a, b = f(Sequence(t["s1"]), Sequence(t["s2"]), _cmp.SCORERS[t["scorer"]], t["a"], t["b"])
        assert (sc, bytes(a), bytes(b)) == (t["score"], t["a_align"].encode(), t["b_align"].encode()), t["ref"]


@pytest.mark.parametrize("idx", range(5))
def test_integration_goldens(aligner, kat, golden_dir, idx):
    t = kat["integration"][idx]
    inp, out = _fixture(golden_dir, t["fixture"])
    f = getattr(aligner, t["mode"] + "_alignment")
    sc, a, b = f(inp[0], inp[1], _cmp.SCORERS[t["scorer"]], t["a"], t["b"])
    assert sc == t["score"]
    assert a == out[0], "a_align differs (len %d vs %d)" % (len(a), len(out[0]))
    assert b == out[1]


def test_config1_from_file_example(aligner, golden_dir):
    """BASELINE config #1: examples/from_file.rs:20-31 (blosum62, open -1, enlarge -2)."""
    inp, _ = _fixture(golden_dir, "semiglobal_alignment")
    sc, a, b = aligner.semiglobal_alignment(inp[0], inp[1], score_mod.blosum62, -1, -2)
    st, osc, oa, ob = orc.align("semiglobal", bytes(inp[0]), bytes(inp[1]), "blosum62", -1, -2, lean=True)
    assert st == orc.OK
    assert sc == osc == 31188
    assert bytes(a) == oa and bytes(b) == ob and len(a) == 11242


def test_edit_distance(kat, golden_dir):
    d = kat["edit_distance_doctest"]
    assert seqmod.edit_distance(Sequence(d["s1"]), Sequence(d["s2"])) == d["distance"]
    t = kat["edit_distance_integration"]
    inp, _ = _fixture(golden_dir, t["fixture"])
    assert seqmod.edit_distance(inp[0], inp[1]) == t["distance"]
    assert seqmod.edit_distance_batch(Tile([inp[0], inp[1], inp[1], inp[0], Sequence(""), inp[0]])) == \
        [t["distance"], t["distance"], len(inp[0])]


def _random_batch(rng, n_pairs, alpha, max_len, edge=False):
    seqs = []
    for _ in range(n_pairs):
        n = rng.randint(0, max_len) if not edge else rng.choice([0, 1, 2, 3, max_len])
        s1 = bytes(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.7 and n > 0:
            s2 = bytearray()
            for c in s1:
                r = rng.random()
                if r < 0.08:
                    s2.append(rng.choice(alpha))
                elif r < 0.12:
                    continue
                elif r < 0.16:
                    s2.append(c); s2.append(rng.choice(alpha))
                else:
                    s2.append(c)
            s2 = bytes(s2)
        else:
            m = rng.randint(0, max_len) if not edge else rng.choice([0, 1, 2, max_len])
            s2 = bytes(rng.choice(alpha) for _ in range(m))
        seqs += [s1, s2]
    return native.Batch.from_sequences(seqs)


PARAMS = [
    ("global", "blosum62", -11, -1), ("global", "unit", -2, -1), ("global", "unit", -1, -1),
    ("local", "blosum62", -11, -1), ("local", "unit", -1, -1), ("local", "pam250", -5, -2), ("local", "unit", 0, 0),
    ("semiglobal", "unit", -1, -1), ("semiglobal", "blosum62", -1, -2), ("semiglobal", "unit", 1, -1),
    ("fitting", "unit", -1, -1), ("fitting", "blosum62", -11, -1),
    ("overlap", "unit", -2, -2), ("overlap", "blosum62", -3, -1),
    # -a - b > 127: the packed kernel takes its score pairs from the shared-memory table, not from byte profiles
    ("global", "unit", -100, -60), ("semiglobal", "unit", -90, -50),
]


@pytest.mark.parametrize("shape", SHAPES + [None])
def test_random_vs_oracle_all_modes(aligner, shape):
    """Small random pairs (incl. empty sequences and ties-heavy 2-letter alphabets), every mode,
    every compiled kernel shape: scores, strings and reference-undefined status must match."""
    rng = random.Random(1234 + (shape[0] * 100 + shape[1] if shape else 0))
    ctx = aligner.context
    if shape:
        ctx.set_shape(*shape)
    try:
        problems = []
        for mode, scorer, a, b in PARAMS:
            for alpha, max_len in ((b"ACGT", 70), (b"AC", 40), (b"ACDEFGHIKLMNPQRSTVWY", 150)):
                batch = _random_batch(rng, 96, alpha, max_len)
                if mode == "fitting":   # reference returns Err for len1 < len2; keep the batch inside the Ok domain
                    seqs = []
                    for p in range(batch.n_pairs):
                        s = [bytes(batch.residues[int(batch.seq_off[2 * p + k]):int(batch.seq_off[2 * p + k + 1])]) for k in (0, 1)]
                        seqs += s if len(s[0]) >= len(s[1]) else s[::-1]
                    batch = native.Batch.from_sequences(seqs)
                eng = _cmp.engine_align(aligner, batch, mode, scorer, a, b)
                ora = _cmp.oracle_align(batch, mode, scorer, a, b, lean=False, threads=8)
                problems += _cmp.diff(batch, eng, ora, "%s/%s/%d/%d/%s" % (mode, scorer, a, b, alpha.decode()[:4]))
                eng.close()
        assert not problems, "\n".join(problems[:40])
    finally:
        ctx.set_shape(0, 0)


def test_edge_lengths_and_multiband(aligner):
    """Lengths around the reference's 1024 buffer edge (A.6) and pairs spanning several column
    bands; lean oracle (the literal one agrees with it, tests/test_oracle_forms.py)."""
    rng = random.Random(7)
    seqs = []
    for n, m in [(1023, 1023), (1024, 5), (5, 1024), (1025, 1025), (1030, 600), (300, 2100), (2100, 300), (513, 512),
                 (512, 513), (1, 700), (700, 1), (0, 600), (600, 0), (33, 33), (161, 159)]:
        s1 = bytes(rng.choice(b"ACGT") for _ in range(n))
        s2 = bytearray(s1[:m]) if m <= n else bytearray(s1 + bytes(rng.choice(b"ACGT") for _ in range(m - n)))
        for k in range(0, len(s2), 11):
            s2[k] = rng.choice(b"ACGT")
        seqs += [s1, bytes(s2)]
    batch = native.Batch.from_sequences(seqs)
    problems = []
    for mode, scorer, a, b in [("global", "unit", -2, -1), ("local", "blosum62", -11, -1), ("semiglobal", "unit", -1, -1),
                               ("overlap", "unit", -2, -2)]:
        eng = _cmp.engine_align(aligner, batch, mode, scorer, a, b)
        ora = _cmp.oracle_align(batch, mode, scorer, a, b, lean=True)
        problems += _cmp.diff(batch, eng, ora, "%s/%s" % (mode, scorer))
        eng.close()
    assert not problems, "\n".join(problems)


def test_error_behaviour(aligner):
    from biogarden_b200.error import InvalidArgumentRange, InvalidInputSize
    s1, s2 = Sequence("ACGTACGT"), Sequence("ACGTTT")
    for f in (aligner.global_alignment, aligner.local_alignment, aligner.fitting_alignment):
        with pytest.raises(InvalidArgumentRange):
            f(s1, s2, score_mod.unit, 1, -1)
        with pytest.raises(InvalidArgumentRange):
            f(s1, s2, score_mod.unit, -1, 2)
    with pytest.raises(InvalidInputSize):
        aligner.fitting_alignment(s2, s1, score_mod.unit, -1, -1)
    # semiglobal / overlap accept positive penalties (aligner.rs:290-296,351-357)
    aligner.semiglobal_alignment(s1, s2, score_mod.unit, 1, 1)
    with pytest.raises(InvalidInputSize):
        aligner.align_batch(Tile([s1, s2, s1]), "global", score_mod.unit, -1, -1)
    # a residue the shipped scorer cannot index (score.rs:40) never reaches the device
    from biogarden_b200.error import ReferenceUndefined
    with pytest.raises(ReferenceUndefined):
        aligner.global_alignment(Sequence("AC-T"), s2, score_mod.unit, -1, -1)
    # raw C-ABI call with a code map that lacks a residue -> BG_EINVAL_RESIDUE
    batch = native.Batch.from_sequences([b"ACGT", b"ACGX"])
    rc = np.full(256, 0xFF, np.uint8); rc[[65, 67, 71, 84]] = [0, 1, 2, 3]
    prm = native.Params("global", -1, -1, np.eye(4, dtype=np.int32), rc, rc)
    from biogarden_b200.error import EngineError
    with pytest.raises(EngineError):
        aligner.context.align_batch(batch, prm)


def test_cfg2_sample_vs_oracle(aligner):
    """BASELINE config #2 shape (150 bp DNA, global, +1/-1, a=-2, b=-1): a 20k-pair prefix of the
    seeded stream against the oracle, digest compare of all strings + exact compare of a few."""
    batch = synth.make("cfg2_dna150_global", n_pairs=20000)
    for shape in (None, (32, 5), (16, 10)):
        if shape:
            aligner.context.set_shape(*shape)
        try:
            eng = _cmp.engine_align(aligner, batch, "global", "unit", -2, -1)
        finally:
            aligner.context.set_shape(0, 0)
        ora = _cmp.oracle_align(batch, "global", "unit", -2, -1, lean=True, want_strings=False)
        assert np.all(ora["status"] == orc.OK)
        assert np.array_equal(eng.score, ora["score"]), "shape %r" % (shape,)
        assert np.all(eng.status == 0)
        assert np.array_equal(_cmp.fnv_pairs(eng), ora["hash"]), "shape %r" % (shape,)
        eng.close()


def test_pipeline_mixed_lengths_multi_chunk(aligner):
    """Host-buffer path with several pipeline chunks and several length classes per chunk
    (100-300 bp): every score and every aligned string (digest) against the oracle."""
    batch = synth.make("cfg3_edit_100_300", n_pairs=60000)
    for mode, a, b in (("global", -2, -1), ("semiglobal", -1, -1)):
        eng = _cmp.engine_align(aligner, batch, mode, "unit", a, b)
        ora = _cmp.oracle_align(batch, mode, "unit", a, b, lean=True, want_strings=False)
        ok = ora["status"] == orc.OK
        assert ok.mean() > 0.95
        assert np.array_equal(eng.score[ok], ora["score"][ok]), mode
        assert np.array_equal(eng.status == 0, ok), mode
        assert np.array_equal(_cmp.fnv_pairs(eng)[ok], ora["hash"][ok]), mode
        eng.close()


def test_cfg4_sample_vs_oracle(aligner):
    """Config #4 shape: protein 200-1000 aa, local, blosum62 -11/-1 (integration.rs:258 parameters)."""
    batch = synth.make("cfg4_protein_local", n_pairs=300)
    eng = _cmp.engine_align(aligner, batch, "local", "blosum62", -11, -1)
    ora = _cmp.oracle_align(batch, "local", "blosum62", -11, -1, lean=True)
    problems = _cmp.diff(batch, eng, ora, "cfg4")
    eng.close()
    assert not problems, "\n".join(problems)


def test_cfg3_sample_vs_oracle():
    batch = synth.make("cfg3_edit_100_300", n_pairs=20000)
    ctx = native.Context()
    got = ctx.edit_distance_batch(batch)
    want, _ = orc.edit_distance_batch(batch.residues, batch.seq_off, threads=orc.hw_threads(), lean=True)
    assert np.array_equal(got, want)
    # protein + arbitrary bytes (the reference compares raw bytes)
    rng = random.Random(5)
    seqs = [bytes(rng.randrange(256) for _ in range(rng.randint(0, 400))) for _ in range(400)]
    b2 = native.Batch.from_sequences(seqs)
    want2, _ = orc.edit_distance_batch(b2.residues, b2.seq_off, threads=4, lean=False)
    assert np.array_equal(ctx.edit_distance_batch(b2), want2)
    ctx.close()


def test_cfg5_small_long_pair(aligner):
    """Config #5 shape at a size the lean oracle finishes in seconds: one 6-9 kbp DNA pair, semiglobal."""
    batch = native.synth_pairs(5, 0, 2, b"ACGT", 6000, 9000, False)
    eng = _cmp.engine_align(aligner, batch, "semiglobal", "unit", -1, -1)
    ora = _cmp.oracle_align(batch, "semiglobal", "unit", -1, -1, lean=True)
    problems = _cmp.diff(batch, eng, ora, "cfg5-small")
    eng.close()
    assert not problems, "\n".join(problems)


def _mutated_long_pairs(shapes, seed):
    rng = random.Random(seed)
    seqs = []
    for n, m in shapes:
        s1 = bytes(rng.choice(b"ACGT") for _ in range(n))
        s2 = bytearray()
        for c in s1:
            r = rng.random()
            if r < 0.06:
                s2.append(rng.choice(b"ACGT"))
            elif r < 0.08:
                continue
            elif r < 0.10:
                s2.append(c); s2.append(rng.choice(b"ACGT"))
            else:
                s2.append(c)
        while len(s2) < m:
            s2.append(rng.choice(b"ACGT"))
        seqs += [s1, bytes(s2[:m])]
    return native.Batch.from_sequences(seqs)


def _fitting_domain(batch):
    """fitting_alignment returns Err(InvalidInputSize) for len1 < len2 (aligner.rs:223-225): swap such pairs."""
    seqs = []
    for p in range(batch.n_pairs):
        s = [bytes(batch.residues[int(batch.seq_off[2 * p + k]):int(batch.seq_off[2 * p + k + 1])]) for k in (0, 1)]
        seqs += s if len(s[0]) >= len(s[1]) else s[::-1]
    return native.Batch.from_sequences(seqs)


def test_wavefront_kernel_long_pairs(aligner):
    """K2 (pairs wider than 4096 columns: one pair per thread-block cluster, 1/2/4/8 CTAs per pair)
    mixed with K1 classes in one batch; all modes against the lean oracle."""
    # (9000, 500), (8000, 200), (20000, 60): K1 classes (32,16), (16,16), (8,8) whose pairs are walked by the long-pair
    # walker's generic-geometry instantiation (several pairs per warp: lane_base) because the batch holds long pairs
    batch = _mutated_long_pairs([(5000, 4200), (3000, 9000), (9000, 17000), (7000, 33000), (300, 5000), (6000, 300), (4500, 4097),
                                 (9000, 500), (8000, 200), (8100, 190), (20000, 60)], 99)
    problems = []
    for mode, scorer, a, b in [("semiglobal", "unit", -1, -1), ("local", "blosum62", -11, -1), ("global", "unit", -2, -1),
                               ("overlap", "unit", -2, -2), ("fitting", "unit", -1, -1)]:
        bt = _fitting_domain(batch) if mode == "fitting" else batch
        eng = _cmp.engine_align(aligner, bt, mode, scorer, a, b)
        ora = _cmp.oracle_align(bt, mode, scorer, a, b, lean=True)
        problems += _cmp.diff(bt, eng, ora, "K2 %s/%s" % (mode, scorer))
        eng.close()
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("budget_mb", [1, 6])
def test_bounded_memory_traceback(budget_mb):
    """Pairs whose 4-bit trace exceeds the long-pair trace budget: row checkpoints + block-wise re-fill and a
    walk resumed from block to block (k2_wave<.., CKPT>, k3_walk_diag with WalkState) must give exactly the
    result of the unbounded path, i.e. the oracle's.  1 MiB: all pairs but the smallest are checkpointed, one or
    two per launch group, in blocks of 64 rows and up (up to 141 blocks); 6 MiB: the two small pairs fit whole,
    the five large ones form one group.  All modes:
    the walk starts below / at / above block boundaries (local, fitting column maximum, semiglobal branches)."""
    al = SequenceAligner()
    al.context.set_long_trace_budget(budget_mb << 20)
    batch = _mutated_long_pairs([(5000, 4200), (2100, 9000), (6100, 4500), (300, 5000), (4096, 4097), (1000, 4100), (9000, 17000)], 7)
    problems = []
    for mode, scorer, a, b in [("semiglobal", "unit", -1, -1), ("local", "blosum62", -11, -1), ("global", "unit", -2, -1),
                               ("overlap", "unit", -2, -2), ("semiglobal", "blosum62", -1, -2), ("fitting", "unit", -1, -1)]:
        bt = _fitting_domain(batch) if mode == "fitting" else batch
        eng = _cmp.engine_align(al, bt, mode, scorer, a, b)
        ora = _cmp.oracle_align(bt, mode, scorer, a, b, lean=True)
        problems += _cmp.diff(bt, eng, ora, "bounded %s/%s %d MiB" % (mode, scorer, budget_mb))
        eng.close()
        if mode == "fitting":
            continue        # (swapped pairs: the refilled-cell bookkeeping below is for the original orientation)
        t = al.context.timing()
        n, m = batch.lengths()
        wave_cells = int(np.sum((n * m)[m > 4096]))
        if budget_mb == 1:   # only the trace of (300, 5000) fits whole
            assert t["cells_refilled"] == wave_cells - 300 * 5000, (t["cells_refilled"], wave_cells)
        else:     # the five largest pairs are checkpointed, (1000, 4100) and (300, 5000) fit whole
            assert t["cells_refilled"] == wave_cells - 1000 * 4100 - 300 * 5000, (t["cells_refilled"], wave_cells)
    # the device-resident entry points go through the same plan
    params = al.make_params(batch, "semiglobal", score_mod.unit, -1, -1)
    db = al.context.upload(batch)
    dres = al.context.align_device(db, params)
    res = al.context.download(dres)
    problems += _cmp.diff(batch, res, _cmp.oracle_align(batch, "semiglobal", "unit", -1, -1, lean=True), "bounded device path")
    res.close()
    al.context.free_result(dres)
    al.context.free_batch(db)
    assert not problems, "\n".join(problems)


def test_full_size_properties_cfg2(aligner):
    """Config #2 at full size (1M pairs): size-independent properties + a random sample vs the oracle.
      * global alignment strings, with '-' removed, are exactly the inputs; equal lengths; no column
        of two gaps;
      * score-only pass == scores of the traceback pass;
      * chunking (trace budget) does not change a single output byte."""
    n_pairs = 1_000_000
    batch = synth.make("cfg2_dna150_global", n_pairs=n_pairs)
    eng = _cmp.engine_align(aligner, batch, "global", "unit", -2, -1)
    assert eng.n_pairs == n_pairs and np.all(eng.status == 0)
    off = eng.off.astype(np.int64)
    la = off[1::2] - off[0:-1:2]; lb = off[2::2] - off[1::2]
    assert np.array_equal(la, lb)
    arena = eng.arena
    # gap-free projection of all a_align strings equals all seq1 concatenated (same for b / seq2)
    is_a = np.zeros(len(arena), bool)
    starts = off[0:-1:2]; ends = off[1::2]
    marks = np.zeros(len(arena) + 1, np.int32)
    np.add.at(marks, starts, 1); np.add.at(marks, ends, -1)
    is_a = np.cumsum(marks[:-1]) > 0
    a_chars = arena[is_a]; b_chars = arena[~is_a]
    assert not np.any((a_chars == 45) & (b_chars == 45))
    so = batch.seq_off.astype(np.int64)
    in_a = np.zeros(len(batch.residues) + 1, np.int32)
    np.add.at(in_a, so[0:-1:2], 1); np.add.at(in_a, so[1::2], -1)
    in_a = np.cumsum(in_a[:-1]) > 0
    assert np.array_equal(a_chars[a_chars != 45], batch.residues[in_a])
    assert np.array_equal(b_chars[b_chars != 45], batch.residues[~in_a])
    score_full = eng.score.copy()
    digest_full = (int(np.sum(arena.astype(np.uint64) * (np.arange(len(arena), dtype=np.uint64) % 1000003))), len(arena))
    # sample vs oracle
    rng = np.random.RandomState(11)
    idx = np.sort(rng.choice(n_pairs, 3000, replace=False))
    seqs = []
    for p in idx:
        seqs += [bytes(batch.residues[so[2 * p]:so[2 * p + 1]]), bytes(batch.residues[so[2 * p + 1]:so[2 * p + 2]])]
    sub = native.Batch.from_sequences(seqs)
    ora = _cmp.oracle_align(sub, "global", "unit", -2, -1, lean=True)
    for q, p in enumerate(idx):
        assert int(ora["score"][q]) == int(eng.score[p])
        assert orc.batch_strings(ora, sub.seq_off, q) == eng.strings(int(p))
    eng.close()
    so_res = _cmp.engine_align(aligner, batch, "global", "unit", -2, -1, score_only=True)
    assert np.array_equal(so_res.score, score_full)
    so_res.close()
    aligner.context.set_trace_budget(1 << 30)
    try:
        eng2 = _cmp.engine_align(aligner, batch, "global", "unit", -2, -1)
    finally:
        aligner.context.set_trace_budget(8 << 30)
    a2 = eng2.arena
    assert np.array_equal(eng2.score, score_full)
    assert (int(np.sum(a2.astype(np.uint64) * (np.arange(len(a2), dtype=np.uint64) % 1000003))), len(a2)) == digest_full
    eng2.close()


@pytest.mark.parametrize("logical", [False, True])
def test_in_process_multi_gpu_sharding_is_result_invariant(logical):
    """bg_create with several devices: every device pulls chunks from one shared queue (long pairs: dealt by size,
    largest first); the output must be byte-identical to the single-device result whatever device aligned what.
    logical=True runs 4 logical shards on ONE GPU (device list [0, 0, 0, 0]), so a 1-GPU box exercises the same code;
    logical=False uses every GPU of the box (and is the same single-device comparison on a 1-GPU box)."""
    import torch
    n_dev = torch.cuda.device_count()
    devs = [0, 0, 0, 0] if logical else list(range(min(n_dev, 8)))
    one = SequenceAligner([0])
    many = SequenceAligner(devs)
    batch = synth.make("cfg3_edit_100_300", n_pairs=150000)    # several chunks per device, six length classes
    for mode, a, b in (("global", -2, -1), ("semiglobal", -1, -1)):
        r1 = _cmp.engine_align(one, batch, mode, "unit", a, b)
        r2 = _cmp.engine_align(many, batch, mode, "unit", a, b)
        assert np.array_equal(r1.score, r2.score) and np.array_equal(r1.status, r2.status)
        assert np.array_equal(r1.off, r2.off) and np.array_equal(r1.arena, r2.arena)
        r1.close(); r2.close()
    e1 = one.context.edit_distance_batch(batch); e2 = many.context.edit_distance_batch(batch)
    assert np.array_equal(e1, e2)
    # long pairs are dealt to the devices by size (K2 + K1 classes mixed), results scattered back to caller order
    lb = _mutated_long_pairs([(5000, 4200), (300, 5000), (3000, 9000), (200, 150), (9000, 17000), (4500, 4097), (6000, 300),
                              (7000, 8000), (150, 150), (4100, 4100)], 5)
    for mode, a, b in (("semiglobal", -1, -1), ("local", -2, -1)):
        r1 = _cmp.engine_align(one, lb, mode, "unit", a, b)
        r2 = _cmp.engine_align(many, lb, mode, "unit", a, b)
        assert np.array_equal(r1.score, r2.score) and np.array_equal(r1.status, r2.status)
        assert np.array_equal(r1.off, r2.off) and np.array_equal(r1.arena, r2.arena)
        r1.close(); r2.close()


def test_compact_results_ops_and_host_expansion(aligner):
    """bg_align_batch_ops: the same alignments as 2-bit ops + bg_expand_ops on the host (both implementations) must
    reproduce the oracle's strings, in every mode, for short (one thread per pair walker), packed (K1h) and long
    (warp-per-pair walker, K2) pairs; and bg_align_batch -- which expands the same ops into its arena -- must agree."""
    rng = random.Random(77)
    impls = [None, 0] + ([1] if native.expand_kind() == "avx512-vbmi2" else [])
    short = _random_batch(rng, 400, b"ACGT", 200)
    prot = _random_batch(rng, 200, b"ACDEFGHIKLMNPQRSTVWY", 300)
    longb = _mutated_long_pairs([(5000, 4200), (300, 5000), (20000, 60), (150, 150), (0, 10), (10, 0), (4500, 4097)], 3)
    problems = []
    for batch, label in ((short, "dna"), (prot, "protein"), (longb, "long")):
        for mode, scorer, a, b in (("global", "unit", -2, -1), ("local", "blosum62", -11, -1), ("semiglobal", "unit", -1, -1),
                                   ("overlap", "unit", -2, -2), ("fitting", "unit", -1, -1)):
            bt = _fitting_domain(batch) if mode == "fitting" else batch
            params = aligner.make_params(bt, mode, _cmp.SCORERS[scorer], a, b)
            ops = aligner.context.align_batch_ops(bt, params)
            full = aligner.context.align_batch(bt, params)
            ora = _cmp.oracle_align(bt, mode, scorer, a, b, lean=True)
            problems += _cmp.diff(bt, full, ora, "strings %s %s" % (label, mode))
            assert np.array_equal(ops.score, full.score) and np.array_equal(ops.status, full.status)
            for p in range(bt.n_pairs):
                want = full.strings(p)
                assert int(ops.len[p]) == len(want[0])
                for impl in impls:
                    assert ops.strings(p, impl) == want, (label, mode, p, impl)
            ops.close(); full.close()
    assert not problems, "\n".join(problems)


def test_cfg5_real_size_pairs_vs_oracle():
    """BASELINE config #5 at its real size: two pairs of the cfg5 stream (50-100 kbp each side, semiglobal, unit
    +1/-1, open -1, extend -1; integration.rs:306 parameters) against the lean oracle, every aligned character --
    once with the whole traces in memory (several GB per pair, 100-200 bands, ~13 CTAs per pair) and once with the
    long-pair trace budget forced down to 1 GiB so that both pairs take the bounded-memory path (row checkpoints,
    block-wise re-fill, resumed walks).  This is the regime of 64-bit trace offsets and multi-CTA group merges."""
    batch = native.synth_pairs(5, 0, 2, b"ACGT", 50000, 100000, False)
    n, m = batch.lengths()
    assert n.min() >= 50000 and m.min() >= 40000
    ora = _cmp.oracle_align(batch, "semiglobal", "unit", -1, -1, lean=True, threads=2)
    assert np.all(ora["status"] == orc.OK)
    al = SequenceAligner()
    for budget in (0, 1 << 30):
        al.context.set_long_trace_budget(budget)
        eng = _cmp.engine_align(al, batch, "semiglobal", "unit", -1, -1)
        problems = _cmp.diff(batch, eng, ora, "cfg5 real size, budget %d" % budget)
        t = al.context.timing()
        if budget:
            assert t["cells_refilled"] == t["cells"], t
        else:
            assert t["cells_refilled"] == 0, t
        eng.close()
        assert not problems, "\n".join(problems)
    al.context.set_long_trace_budget(0)


def test_cfg4_full_size_sample_and_chunk_invariance(aligner):
    """Config #4 at full size (100 000 protein pairs, 200-1000 aa, local, blosum62 -11/-1): 5 000 random pairs
    against the lean oracle (scores + every string), and the whole output byte-identical under a different chunking
    (trace budget 1 GiB instead of 8: several launches per length class)."""
    n_pairs = 100_000
    batch = synth.make("cfg4_protein_local", n_pairs=n_pairs)
    eng = _cmp.engine_align(aligner, batch, "local", "blosum62", -11, -1)
    assert np.all(eng.status == 0)
    so = batch.seq_off.astype(np.int64)
    rng = np.random.RandomState(4)
    idx = np.sort(rng.choice(n_pairs, 5000, replace=False))
    seqs = []
    for p in idx:
        seqs += [bytes(batch.residues[so[2 * p]:so[2 * p + 1]]), bytes(batch.residues[so[2 * p + 1]:so[2 * p + 2]])]
    sub = native.Batch.from_sequences(seqs)
    ora = _cmp.oracle_align(sub, "local", "blosum62", -11, -1, lean=True)
    for q, p in enumerate(idx):
        assert int(ora["score"][q]) == int(eng.score[p]), p
        assert orc.batch_strings(ora, sub.seq_off, q) == eng.strings(int(p)), p
    score_full = eng.score.copy(); off_full = eng.off.copy(); arena_full = eng.arena.copy()
    eng.close()
    aligner.context.set_trace_budget(1 << 30)
    try:
        eng2 = _cmp.engine_align(aligner, batch, "local", "blosum62", -11, -1)
    finally:
        aligner.context.set_trace_budget(8 << 30)
    assert np.array_equal(eng2.score, score_full) and np.array_equal(eng2.off, off_full) and np.array_equal(eng2.arena, arena_full)
    eng2.close()


def test_cfg3_full_size_sample_and_path_invariance():
    """Config #3 at one GPU's share (1 250 000 pairs, 100-300 bp): 5 000 random pairs against the oracle, and the
    host-buffer pipeline (24 chunks) byte-identical to the device-resident pass (one launch per length class)."""
    n_pairs = 1_250_000
    batch = synth.make("cfg3_edit_100_300", n_pairs=n_pairs)
    ctx = native.Context()
    got = ctx.edit_distance_batch(batch)
    so = batch.seq_off.astype(np.int64)
    rng = np.random.RandomState(3)
    idx = np.sort(rng.choice(n_pairs, 5000, replace=False))
    seqs = []
    for p in idx:
        seqs += [bytes(batch.residues[so[2 * p]:so[2 * p + 1]]), bytes(batch.residues[so[2 * p + 1]:so[2 * p + 2]])]
    sub = native.Batch.from_sequences(seqs)
    want, _ = orc.edit_distance_batch(sub.residues, sub.seq_off, threads=orc.hw_threads(), lean=True)
    assert np.array_equal(got[idx], want)
    db = ctx.upload(batch, 0, prepare="edit")
    r = ctx.edit_distance_device(db)
    assert np.array_equal(ctx.download_u64(r, n_pairs), got)
    ctx.free_result(r); ctx.free_batch(db)
    ctx.close()


def test_edit_distance_bit_parallel_and_fallback():
    """K4b (Myers bit-vectors, <= 4 byte values) against the oracle on mixed lengths (also > 320 columns,
    which stay on the systolic kernel), through both the host-buffer and the device-resident entry points;
    plus the fallback: a fifth symbol hidden in the middle of a large batch (the alphabet is sampled from
    the ends) must be caught on the device and the batch redone with the general kernel."""
    rng = random.Random(31)
    seqs = []
    for _ in range(3000):
        n = rng.choice([0, 1, 31, 32, 33, 64, 100, 128, 129, 200, 256, 257, 300, 320, 321, 400, 700])
        m = rng.choice([0, 1, 31, 32, 33, 64, 100, 128, 129, 200, 256, 257, 300, 320, 321, 400, 700])
        s1 = bytes(rng.choice(b"ACGT") for _ in range(n))
        s2 = bytearray(s1[:m]) + bytes(rng.choice(b"ACGT") for _ in range(max(0, m - n)))
        for k in range(0, len(s2), 7):
            s2[k] = rng.choice(b"ACGT")
        seqs += [s1, bytes(s2)]
    batch = native.Batch.from_sequences(seqs)
    want, _ = orc.edit_distance_batch(batch.residues, batch.seq_off, threads=orc.hw_threads(), lean=True)
    ctx = native.Context()
    assert np.array_equal(ctx.edit_distance_batch(batch), want)
    db = ctx.upload(batch, 0, prepare="edit")
    r = ctx.edit_distance_device(db)
    assert np.array_equal(ctx.download_u64(r, batch.n_pairs), want)
    ctx.free_result(r); ctx.free_batch(db)
    # every len2 <= 320: the chunk's launch slots are built on the device (k0_eplan.cuh); same results as with the
    # host planner, for raw and for 2-bit packed residues, with empty sequences and len1 up to 700
    seqs2 = []
    for k in range(0, len(seqs), 2):
        a_, b_ = seqs[k], seqs[k + 1]
        seqs2 += [a_, b_] if len(b_) <= 320 else ([b_, a_] if len(a_) <= 320 else [a_[:300], b_[:320]])
    batch2 = native.Batch.from_sequences(seqs2 * 6)     # 18 000 pairs: several pipeline chunks
    assert int(np.diff(batch2.seq_off.astype(np.int64))[1::2].max()) <= 320
    want3, _ = orc.edit_distance_batch(batch2.residues, batch2.seq_off, threads=orc.hw_threads(), lean=True)
    got_dev = ctx.edit_distance_batch(batch2)
    assert np.array_equal(got_dev, want3)
    assert np.array_equal(ctx.edit_distance_batch(batch2.pack(2)), want3)
    ctx.set_host_plan(True)
    try:
        assert np.array_equal(ctx.edit_distance_batch(batch2), want3)
    finally:
        ctx.set_host_plan(False)
    # the device-planned path does not read the offsets on the host: a single long pair (len2 > 320) or a broken offset
    # hidden between the pairs the host samples must be caught on the device and go through the general path
    seqs3 = list(seqs2 * 6)
    seqs3[2 * 7], seqs3[2 * 7 + 1] = seqs3[2 * 7] + b"ACGT" * 100, (seqs3[2 * 7 + 1] + b"TTGACA" * 80)
    batch3 = native.Batch.from_sequences(seqs3)
    want4, _ = orc.edit_distance_batch(batch3.residues, batch3.seq_off, threads=orc.hw_threads(), lean=True)
    assert np.array_equal(ctx.edit_distance_batch(batch3), want4)
    bad_off = batch2.seq_off.copy()
    bad_off[2 * 11 + 1] = bad_off[2 * 11] - 1 if bad_off[2 * 11] > 0 else bad_off[2 * 11 + 2] + 1
    from biogarden_b200.error import EngineError
    with pytest.raises(EngineError):
        ctx.edit_distance_batch(native.Batch(batch2.residues, bad_off))
    # fallback
    big = synth.make("cfg3_edit_100_300", n_pairs=30000)
    res = big.residues.copy()
    mid = int(big.seq_off[30001])
    res[mid + 5] = ord("N")
    big2 = native.Batch(res, big.seq_off)
    want2, _ = orc.edit_distance_batch(big2.residues, big2.seq_off, threads=orc.hw_threads(), lean=True)
    assert np.array_equal(ctx.edit_distance_batch(big2), want2)
    ctx.close()


def test_hamming_distance_goldens_and_random(kat, golden_dir):
    """K5 hamming (seq.rs:74-83) through the C ABI: the reference's doctest and integration golden, then random
    batches against the oracle -- every alignment of the two sequences relative to 16 / 4 bytes, lengths around
    the vector width, empty pairs, pairs longer than one 64 KiB piece -- and the error for unequal lengths."""
    from biogarden_b200.error import InvalidInputSize
    d = kat["hamming_distance_doctest"]
    assert seqmod.hamming_distance(Sequence(d["s1"]), Sequence(d["s2"])) == d["distance"]
    t = kat["hamming_distance_integration"]
    inp, _ = _fixture(golden_dir, t["fixture"])
    assert seqmod.hamming_distance(inp[0], inp[1]) == t["distance"]
    with pytest.raises(InvalidInputSize):
        seqmod.hamming_distance(Sequence("ACGT"), Sequence("ACG"))
    with pytest.raises(InvalidInputSize):
        seqmod.hamming_distance_batch(Tile([Sequence("AC"), Sequence("AC"), Sequence("A"), Sequence("AC")]))
    rng = random.Random(5)
    seqs = []
    for ln in list(range(0, 70)) + [127, 128, 129, 4095, 65535, 65536, 65537, 200001]:
        s1 = bytes(rng.choice(b"ACGT") for _ in range(ln))
        s2 = bytes(c if rng.random() < 0.8 else rng.choice(b"ACGT") for c in s1)
        seqs += [s1, s2]
    batch = native.Batch.from_sequences(seqs)
    ctx = native.Context()
    got = ctx.hamming_distance_batch(batch)
    want = [orc.hamming_distance(seqs[2 * p], seqs[2 * p + 1])[1] for p in range(len(seqs) // 2)]
    assert got.tolist() == want
    ctx.close()


def test_p_distance_matrix_vs_oracle():
    """K5 p_distance_matrix (stat.rs:138-152): bit-exact f32 against the oracle, ragged rows (zip to the shorter
    one), an empty row, one row only, a 300 x 2000 DNA matrix; empty Tile -> the reference panics."""
    from biogarden_b200 import stat
    rng = random.Random(11)
    cases = [[b"ACGTACGTAC", b"ACGTTCGTAA", b"TTTTTTTTTT", b"ACGTACG", b""], [b"ACGT"],
             [bytes(rng.choice(b"ACGT") for _ in range(2000)) for _ in range(300)],
             [bytes(rng.choice(b"ACDEFGHIKL") for _ in range(rng.randint(1, 700))) for _ in range(41)]]
    for rows in cases:
        got = stat.p_distance_matrix(Tile([Sequence(r) for r in rows]))
        want = orc.p_distance_matrix(rows)
        assert got.dtype == np.float32 and got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    with pytest.raises(IndexError):
        stat.p_distance_matrix(Tile([]))


def test_user_closure_scorer(aligner):
    """A scorer that is not one of the shipped tables (any `Fn(&u8,&u8)->i32`): match +5 /
    mismatch -4 through the table materialisation, every mode, against the oracle driven by the same 256x256 table;
    DNA (packed kernel for the non-local modes) and a 20-letter alphabet (int32 kernel, dense table in shared memory)."""
    rng = random.Random(21)
    problems = []
    for alpha, max_len in ((b"ACGT", 300), (b"ACDEFGHIKLMNPQRSTVWY", 260)):
        batch = _random_batch(rng, 300, alpha, max_len)
        for mode, a, b in (("global", -10, -1), ("local", -6, -2), ("semiglobal", -3, -3), ("overlap", -5, -1), ("fitting", -10, -1)):
            bt = _fitting_domain(batch) if mode == "fitting" else batch
            eng = _cmp.engine_align(aligner, bt, mode, None, a, b, match_mismatch=(5, -4))
            ora = _cmp.oracle_align(bt, mode, None, a, b, match_mismatch=(5, -4))
            problems += _cmp.diff(bt, eng, ora, "closure %s %s" % (mode, alpha[:4]))
            eng.close()
    assert not problems, "\n".join(problems)


def test_device_planner_matches_host_planner(aligner):
    """k0_plan.cuh (launch descriptors built on the GPU from the sequence offsets) against build_plan (the host
    planner): the two descriptor arrays must agree field by field -- slot order (stable sort by class, len1
    descending), K1h holes, step counts, trace / output / band-scratch offsets -- for uniform chunks (no sort), mixed
    lengths (holes), protein classes, multi-band K1 pairs, pairs long enough for the warp walker's shape set, and
    forced shapes; and the alignments must not depend on which planner ran."""
    rng = random.Random(12)
    batches = {
        "cfg2": synth.make("cfg2_dna150_global", n_pairs=20000),
        "cfg2_odd": synth.make("cfg2_dna150_global", n_pairs=20001),     # uniform chunk (k_plan_uniform), a hole behind the last pair
        "uniform_multiband": native.synth_pairs(9, 0, 37, b"ACGT", 1500, 1500, True),  # uniform, two bands per pair (band scratch offsets)
        "uniform_protein": native.synth_pairs(4, 0, 999, synth.PROTEIN, 300, 300, True),
        "cfg3": synth.make("cfg3_edit_100_300", n_pairs=70000),
        "cfg4": synth.make("cfg4_protein_local", n_pairs=3000),
        "ragged": _random_batch(rng, 5000, b"ACGT", 260),
        "multiband": _mutated_long_pairs([(300, 1500), (1200, 1100), (2500, 4000), (100, 4096), (700, 700), (20000, 90), (17000, 300), (50, 50),
                                          (1030, 1030), (1030, 1030), (1031, 1030)] * 3, 8),
    }
    ctx = aligner.context
    for name, batch in batches.items():
        for mode, scorer, a, b in (("global", "unit", -2, -1), ("local", "blosum62", -11, -1), ("semiglobal", "unit", -1, -1)):
            if name in ("cfg4", "uniform_protein") and scorer == "unit":
                continue
            for shape in (None, (8, 19), (32, 8)):
                if shape and name not in ("ragged", "cfg3"):
                    continue
                if shape:
                    ctx.set_shape(*shape)
                try:
                    params = aligner.make_params(batch, mode, _cmp.SCORERS[scorer], a, b)
                    elig, compared, differing, surplus = ctx.plan_compare(batch, params)
                    assert elig == 1, (name, mode, shape)
                    assert compared >= batch.n_pairs and differing == 0 and surplus == 0, (name, mode, shape, compared, differing, surplus)
                    ctx.set_host_plan(True)
                    r_host = ctx.align_batch(batch, params)
                    ctx.set_host_plan(False)
                    r_dev = ctx.align_batch(batch, params)
                    assert np.array_equal(r_host.score, r_dev.score) and np.array_equal(r_host.status, r_dev.status), (name, mode, shape)
                    assert np.array_equal(r_host.off, r_dev.off) and np.array_equal(r_host.arena, r_dev.arena), (name, mode, shape)
                    r_host.close(); r_dev.close()
                finally:
                    ctx.set_host_plan(False)
                    ctx.set_shape(0, 0)


def test_packed_batches_give_identical_results(aligner):
    """Packed residues at the ABI (bg_batch::packing = 2 / 5 bit, what bg_fasta_parse_packed emits): every entry
    point must return exactly what it returns for the byte batch -- alignment through the chunk pipeline (chunk
    boundaries fall at arbitrary 2-bit offsets), compact ops results, long pairs, the device-resident path, edit
    distance (bit-parallel and general kernel), hamming -- with the strings expanded from host-unpacked residues."""
    rng = random.Random(41)
    dna = synth.make("cfg3_edit_100_300", n_pairs=70000)
    dna_p = dna.pack(2)
    ctx = aligner.context
    for mode, a, b in (("global", -2, -1), ("semiglobal", -1, -1), ("local", -3, -1)):
        params = aligner.make_params(dna, mode, score_mod.unit, a, b)
        params_p = aligner.make_params(dna_p, mode, score_mod.unit, a, b)
        assert np.array_equal(params.table, params_p.table)
        r0 = ctx.align_batch(dna, params); r1 = ctx.align_batch(dna_p, params_p)
        assert np.array_equal(r0.score, r1.score) and np.array_equal(r0.status, r1.status)
        assert np.array_equal(r0.off, r1.off) and np.array_equal(r0.arena, r1.arena), mode
        r0.close(); r1.close()
    prot = _random_batch(rng, 500, b"ACDEFGHIKLMNPQRSTVWY", 400)
    prot_p = prot.pack(5)
    r0 = _cmp.engine_align(aligner, prot, "local", "blosum62", -11, -1); r1 = _cmp.engine_align(aligner, prot_p, "local", "blosum62", -11, -1)
    assert np.array_equal(r0.score, r1.score) and np.array_equal(r0.off, r1.off) and np.array_equal(r0.arena, r1.arena)
    r0.close(); r1.close()
    longb = _mutated_long_pairs([(5000, 4200), (301, 5003), (20001, 61), (149, 150), (4500, 4097)], 13)
    long_p = longb.pack(2)
    r0 = _cmp.engine_align(aligner, longb, "semiglobal", "unit", -1, -1); r1 = _cmp.engine_align(aligner, long_p, "semiglobal", "unit", -1, -1)
    assert np.array_equal(r0.score, r1.score) and np.array_equal(r0.off, r1.off) and np.array_equal(r0.arena, r1.arena)
    # device-resident path and compact results
    params_p = aligner.make_params(long_p, "semiglobal", score_mod.unit, -1, -1)
    db = ctx.upload(long_p); dres = ctx.align_device(db, params_p); r2 = ctx.download(dres)
    assert np.array_equal(r0.score, r2.score) and np.array_equal(r0.arena, r2.arena)
    r2.close(); ctx.free_result(dres); ctx.free_batch(db)
    r0.close(); r1.close()
    small = native.Batch(dna.residues[:int(dna.seq_off[2000])], dna.seq_off[:2001])
    small_p = small.pack(2)
    e0 = ctx.edit_distance_batch(dna); e1 = ctx.edit_distance_batch(dna_p)
    assert np.array_equal(e0, e1)
    assert np.array_equal(ctx.edit_distance_batch(prot), ctx.edit_distance_batch(prot_p))
    eq = native.Batch.from_sequences([s for p in range(200) for s in (lambda x: (x, bytes(c if rng.random() < 0.9 else rng.choice(b"ACGT") for c in x)))(bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 300))))])
    assert np.array_equal(ctx.hamming_distance_batch(eq), ctx.hamming_distance_batch(eq.pack(2)))
    assert small_p.packing == 2


def test_fine_wavefront_kernel_vs_oracle_and_k2():
    """K2f (k2f_fine.cuh: one column per lane, shared-memory hand-over between warps, direction codes transposed into
    K2's trace layout) against the lean oracle in every mode, forced onto launches of many pairs (pairs run one after
    the other inside the launch), with several warps-per-CTA settings so that both the shared-memory and the
    global-memory hand-over carry the boundary columns; and byte-identical to K2 (fine kernel switched off)."""
    batch = _mutated_long_pairs([(5000, 4200), (3000, 9000), (300, 5000), (4500, 4097), (9000, 17000), (33, 4100), (4100, 4099), (1, 5000), (0, 4200)], 77)
    al = SequenceAligner()
    problems = []
    for mode, scorer, a, b in [("semiglobal", "unit", -1, -1), ("local", "blosum62", -11, -1), ("global", "unit", -2, -1),
                               ("overlap", "unit", -2, -2), ("fitting", "unit", -1, -1), ("semiglobal", "blosum62", -1, -2)]:
        bt = _fitting_domain(batch) if mode == "fitting" else batch
        al.context.set_fine_pairs(0)
        k2 = _cmp.engine_align(al, bt, mode, scorer, a, b)
        al.context.set_fine_pairs(64)
        fine = _cmp.engine_align(al, bt, mode, scorer, a, b)
        ora = _cmp.oracle_align(bt, mode, scorer, a, b, lean=True)
        problems += _cmp.diff(bt, fine, ora, "K2f %s/%s" % (mode, scorer))
        assert np.array_equal(k2.score, fine.score) and np.array_equal(k2.status, fine.status), mode
        assert np.array_equal(k2.off, fine.off) and np.array_equal(k2.arena, fine.arena), mode
        k2.close(); fine.close()
    assert not problems, "\\n".join(problems)
    al.context.set_fine_pairs(0)
