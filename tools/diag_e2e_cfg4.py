import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from biogarden_b200 import native, score, synth
from biogarden_b200.aligner import SequenceAligner
batch = bench.pinned_batch(synth.make("cfg4_protein_local", n_pairs=100000))
al = SequenceAligner([0]); ctx = al.context
params = al.make_params(batch, "local", score.blosum62, -11, -1)
for i in range(4):
    t0 = time.perf_counter(); r = ctx.align_batch(batch, params); t1 = time.perf_counter(); r.close()
    print("align_batch %.2f ms" % (1e3 * (t1 - t0)), flush=True)
