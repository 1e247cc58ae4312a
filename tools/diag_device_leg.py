import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biogarden_b200 import native, score, synth
from biogarden_b200.aligner import SequenceAligner
batch = synth.make("cfg2_dna150_global", n_pairs=1000000)
al = SequenceAligner([0]); ctx = al.context
params = al.make_params(batch, "global", score.unit, -2, -1)
db = ctx.upload(batch, 0, prepare="align"); ctx.sync()
for keep in (False, True):
    held = []
    for i in range(6):
        t0 = time.perf_counter()
        r = ctx.align_device(db, params)
        t1 = time.perf_counter()
        ctx.sync()
        t2 = time.perf_counter()
        tm = ctx.timing()
        if keep: held.append(r)
        else: ctx.free_result(r)
        t3 = time.perf_counter()
        print("keep=%s step %d: call %.2f ms, sync %.2f ms, free %.2f ms; phases fill %.2f walk %.2f compact %.2f total %.2f" % (keep, i, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), tm["fill_ms"], tm["walk_ms"], tm["compact_ms"], tm["total_ms"]))
    for r in held: ctx.free_result(r)
