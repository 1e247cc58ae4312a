"""Mid-length DNA pairs (the north star's 150-500 bp range), global affine + traceback, device-resident:
GCUPS of the whole path and of the fill, with the packed kernel (default) and without (BG_NO_HALF=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biogarden_b200 import native, score
from biogarden_b200.aligner import SequenceAligner
al = SequenceAligner([0]); ctx = al.context
for ln, n in ((150, 400000), (250, 200000), (350, 120000), (500, 60000), (700, 30000), (1000, 16000)):
    batch = native.synth_pairs(2, 0, n, b"ACGT", ln, ln, True)
    params = al.make_params(batch, "global", score.unit, -2, -1)
    db = ctx.upload(batch, 0, prepare="align"); ctx.sync()
    for i in range(3):
        r = ctx.align_device(db, params); ctx.sync(); tm = ctx.timing(); ctx.free_result(r)
    ctx.free_batch(db)
    cells = batch.cells()
    tot = tm["fill_ms"] + tm["walk_ms"] + tm["compact_ms"]
    print("len %4d x %7d pairs: fill %7.3f ms = %6.0f GCUPS, walk %6.3f, compact %6.3f -> %6.0f GCUPS; packed cells %d%%" % (
        ln, n, tm["fill_ms"], cells / tm["fill_ms"] / 1e6, tm["walk_ms"], tm["compact_ms"], cells / tot / 1e6,
        100 * tm["cells_packed16"] // max(1, tm["cells"])), flush=True)
