"""Host-side timeline of the host-buffer path: a few bg_align_batch calls (BG_PROFILE_HOST=1 prints the timeline).
usage: python tools/diag_e2e.py [workload] [pairs] [host_plan 0/1] [n_devices] [ops|strings] [real|logical]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from biogarden_b200 import score, synth
from biogarden_b200.aligner import SequenceAligner
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg_name, full = bench.WORKLOADS[wl]
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else full
host_plan = len(sys.argv) > 3 and sys.argv[3] == "1"
ndev = int(sys.argv[4]) if len(sys.argv) > 4 else 1
kind = sys.argv[5] if len(sys.argv) > 5 else "strings"
real = len(sys.argv) > 6 and sys.argv[6] == "real"
cfg = synth.CONFIGS[cfg_name]
batch = bench.pinned_batch(bench.make_batch(cfg_name, pairs))
al = SequenceAligner(list(range(ndev)) if real else [0] * ndev)
al.context.set_host_plan(host_plan)
params = al.make_params(batch, cfg["mode"], getattr(score, cfg["scorer"]), cfg["a"], cfg["b"])
fn = al.context.align_batch_ops if kind == "ops" else al.context.align_batch
ts = []
for i in range(8):
    t0 = time.perf_counter(); r = fn(batch, params); ts.append(time.perf_counter() - t0); r.close()
print("calls (ms):", " ".join("%.2f" % (1e3 * x) for x in ts), "host_plan", host_plan, "ndev", ndev, kind, "real" if real else "logical", file=sys.stderr)
