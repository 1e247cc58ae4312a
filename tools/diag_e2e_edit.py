"""cfg3 through bg_edit_distance_batch (pinned host buffers): wall time per call, raw and 2-bit packed residues.
BG_PROFILE_HOST=1 prints the pipeline's timeline."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from biogarden_b200 import native, synth
raw = synth.make("cfg3_edit_100_300", n_pairs=1250000)
b3 = bench.pinned_batch(raw)
b3p = bench.pinned_batch(raw.pack(2))
ctx = native.Context([0])
for name, b in (("raw", b3), ("packed", b3p)):
    for i in range(4):
        t0 = time.perf_counter(); ctx.edit_distance_batch(b); t1 = time.perf_counter()
        print("%s edit_distance_batch %.2f ms" % (name, 1e3 * (t1 - t0)), flush=True)
