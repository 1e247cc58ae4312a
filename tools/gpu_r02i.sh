O=gpurun_out/r02i; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "wavefront or bounded or cfg5 or sharding" 2>&1 | tail -n 4
python bench.py --workload cfg5 --steps 3 --warmup 1 --no-cpu-baseline > $O/bench_cfg5.log 2>&1
BG_NO_WAVE_OVERLAP=1 python bench.py --workload cfg5 --steps 3 --warmup 1 --no-cpu-baseline > $O/bench_cfg5_noov.log 2>&1
python - <<'PY'
import json
for f in ("bench_cfg5", "bench_cfg5_noov"):
    d = json.loads(open("gpurun_out/r02i/%s.log" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 1), d["phases_ms_last_step"], "frac", d["roofline"]["frac"], d["roofline"]["launches_per_step"])
PY
