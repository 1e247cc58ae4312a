"""One host-buffer call (bg_align_batch on a packed cfg2 batch) -- the command ncu captures the pipeline's small
kernels from: k_unpack2 (packed residues -> bytes), k_plan_* (device-side launch planner), k_ops_counts / k_pack_ops /
k_ops_sample (compact results).   python tools/ncu_capture_e2e.py N_PAIRS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biogarden_b200 import score, synth
from biogarden_b200.aligner import SequenceAligner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
batch = synth.make("cfg3_edit_100_300", n_pairs=n).pack(2)      # mixed lengths: the planner sorts, K1h classes have holes
al = SequenceAligner([0])
params = al.make_params(batch, "global", score.unit, -2, -1)
for _ in range(2):
    r = al.context.align_batch(batch, params); t = al.context.timing(); r.close()
print("e2e capture: %d pairs, %d cells, h2d %d B, d2h %d B, %d launches" % (n, t["cells"], t["h2d_bytes"], t["d2h_bytes"], t["launches"]))
