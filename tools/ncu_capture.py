"""One device-resident pass of a named workload at a given pair count -- the command ncu captures kernels from.

  python tools/ncu_capture.py cfg2|cfg3|cfg4|cfg4u|cfg5 N_PAIRS [REPS]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biogarden_b200 import score, synth
from biogarden_b200.aligner import SequenceAligner

WL = {"cfg2": ("cfg2_dna150_global", "global", score.unit, -2, -1),
      "cfg4": ("cfg4_protein_local", "local", score.blosum62, -11, -1),
      "cfg5": ("cfg5_long_semiglobal", "semiglobal", score.unit, -1, -1)}
name, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
if name == "cfg3":        # edit distance, score only (bit-parallel kernel, one launch per 32-column-block class)
    batch = synth.make("cfg3_edit_100_300", n_pairs=n)
    from biogarden_b200 import native
    ctx = native.Context([0])
    db = ctx.upload(batch, 0, prepare="edit"); ctx.sync()
    for i in range(reps):
        r = ctx.edit_distance_device(db); ctx.sync(); tm = ctx.timing(); ctx.free_result(r)
    print("cfg3 %d pairs, %d cells: fill %.3f ms (%d launches)" % (n, batch.cells(), tm["fill_ms"], tm["fill_launches"]))
    sys.exit(0)
if name == "cfg1":        # the reference's example pair: one K2 launch of 17 one-warp CTAs + the long-pair walker
    import bench
    batch = bench.make_batch("cfg1_from_file", n)
    al = SequenceAligner([0]); ctx = al.context
    params = al.make_params(batch, "semiglobal", score.blosum62, -1, -2)
    db = ctx.upload(batch, 0, prepare="align"); ctx.sync()
    for i in range(reps):
        r = ctx.align_device(db, params); ctx.sync(); tm = ctx.timing(); ctx.free_result(r)
    print("cfg1 %d pairs, %d cells: fill %.3f ms, walk %.3f ms" % (n, batch.cells(), tm["fill_ms"], tm["walk_ms"]))
    sys.exit(0)
if name == "cfg3e2e":     # edit distance through the host-buffer entry point: K0e planner + K4b per chunk
    from biogarden_b200 import native
    batch = synth.make("cfg3_edit_100_300", n_pairs=n).pack(2)
    ctx = native.Context([0])
    for i in range(reps):
        out = ctx.edit_distance_batch(batch); tm = ctx.timing()
    print("cfg3e2e %d pairs, %d cells, %d launches" % (n, tm["cells"], tm["launches"]))
    sys.exit(0)
if name == "cfg4u":      # one length class of config #4: uniform 600 aa protein pairs -> a single k1_fill<32,20,local> launch
    from biogarden_b200 import native
    mode, sc, a, b = "local", score.blosum62, -11, -1
    batch = native.synth_pairs(4, 0, n, synth.PROTEIN, 600, 600, True)
else:
    cfg, mode, sc, a, b = WL[name]
    batch = synth.make(cfg, n_pairs=n)
al = SequenceAligner([0]); ctx = al.context
params = al.make_params(batch, mode, sc, a, b)
db = ctx.upload(batch, 0, prepare="align"); ctx.sync()
for i in range(reps):
    r = ctx.align_device(db, params); ctx.sync(); tm = ctx.timing(); ctx.free_result(r)
print("%s %d pairs, %d cells: fill %.3f ms (%d launches), walk %.3f ms, compact %.3f ms" %
      (name, n, batch.cells(), tm["fill_ms"], tm["fill_launches"], tm["walk_ms"], tm["compact_ms"]))
