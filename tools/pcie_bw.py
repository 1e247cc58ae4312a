"""Measured host<->device copy bandwidth of the box (pinned memory), one direction and both at once.
The e2e leg of bench.py cannot be faster than bytes / these rates."""
import json, time, torch
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * reps / dt / 1e9
for _ in range(2): run(True, True, 2)
out = {"h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "both_each_gbs": run(True, True)}
# small-copy sizes like one pipeline chunk
for mb in (4, 16, 48):
    k = mb << 20
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): h2[:k].copy_(d2[:k], non_blocking=True)
    torch.cuda.synchronize(); out["d2h_%dMB_gbs" % mb] = k * 20 / (time.perf_counter() - t0) / 1e9
print(json.dumps(out))
