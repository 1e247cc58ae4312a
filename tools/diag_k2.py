import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biogarden_b200 import native, score
from biogarden_b200.aligner import SequenceAligner
al = SequenceAligner([0]); ctx = al.context
def run(label, npairs, length, shape=None, reps=2):
    batch = native.synth_pairs(5, 0, npairs, b"ACGT", length, length, True)
    params = al.make_params(batch, "semiglobal", score.unit, -1, -1)
    if shape: ctx.set_shape(*shape)
    db = ctx.upload(batch, 0, prepare="align"); ctx.sync()
    for i in range(reps):
        r = ctx.align_device(db, params); ctx.sync(); tm = ctx.timing(); ctx.free_result(r)
    ctx.free_batch(db); ctx.set_shape(0, 0)
    cells = batch.cells()
    print("%-40s pairs %5d len %6d: fill %8.2f ms (%7.1f GCUPS)  walk %7.2f ms  launches %d" % (label, npairs, length, tm["fill_ms"], cells / tm["fill_ms"] / 1e6, tm["walk_ms"], tm["launches"]), flush=True)
run("K2 uniform 2 rounds (Q=4)", 37, 65536)
run("K2 uniform 1 round  (Q=4)", 37, 32768)
run("K2 uniform 1.5 rounds (Q=4)", 37, 49152)
run("K2 74 pairs 1 round", 74, 32768)
run("K2 few pairs Q=8", 16, 65536)
run("K1 multiband (32,16) 4 bands", 4736, 2048, shape=(32, 16))
run("K1 multiband (32,16) 16 bands", 2368, 8192, shape=(32, 16))
run("K1 single band (32,16)", 18944, 512, shape=(32, 16))
