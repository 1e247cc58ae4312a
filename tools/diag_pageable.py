"""Host-to-host time of bg_align_batch on cfg2 with the caller's arena pageable, registered (bg_pin_host) and
allocated pinned."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from biogarden_b200 import native, score, synth
from biogarden_b200.aligner import SequenceAligner
raw = synth.make("cfg2_dna150_global", n_pairs=1000000)
al = SequenceAligner([0]); ctx = al.context
params = al.make_params(raw, "global", score.unit, -2, -1)
def run(label, batch):
    best = 1e9
    for i in range(5):
        t0 = time.perf_counter(); r = ctx.align_batch(batch, params); t1 = time.perf_counter(); r.close()
        best = min(best, 1e3 * (t1 - t0))
    print("%-34s %.2f ms" % (label, best), flush=True)
run("pageable numpy arena", raw)
L = native.lib()
assert L.bg_pin_host(raw.residues.ctypes.data, raw.residues.nbytes) == 0
assert L.bg_pin_host(raw.seq_off.ctypes.data, raw.seq_off.nbytes) == 0
run("same arena after bg_pin_host", raw)
L.bg_unpin_host(raw.residues.ctypes.data); L.bg_unpin_host(raw.seq_off.ctypes.data)
run("cudaHostAlloc'd copy (bench)", bench.pinned_batch(raw))
