import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from biogarden_b200 import native, synth
b3 = bench.pinned_batch(synth.make("cfg3_edit_100_300", n_pairs=1250000))
ctx = native.Context([0])
for i in range(4):
    t0 = time.perf_counter(); ctx.edit_distance_batch(b3); t1 = time.perf_counter()
    print("edit_distance_batch %.2f ms" % (1e3 * (t1 - t0)), flush=True)
