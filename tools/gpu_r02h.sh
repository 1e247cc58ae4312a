O=gpurun_out/r02h; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "random_vs_oracle or edge or wavefront or bounded or cfg2_sample or cfg4_sample or golden or doctest or closure" 2>&1 | tail -n 4
(time python bench.py --steps 10 --warmup 3) > $O/bench.log 2> $O/bench.err; tail -n 3 $O/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02h/bench.log").read().strip().splitlines()[-1])
print("cfg2", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["phases_ms_last_step"], "frac", d["roofline"]["frac"])
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"], 1), "e2e", round(v["e2e"]["value"], 1), v["phases_ms_last_step"], "frac", v["roofline"]["frac"])
PY
