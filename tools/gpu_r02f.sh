O=gpurun_out/r02f; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "packed or compact or planner" 2>&1 | tail -n 12
(time python bench.py --steps 10 --warmup 3) > $O/bench.log 2> $O/bench.err; tail -n 3 $O/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f/bench.log").read().strip().splitlines()[-1])
print("cfg2", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], "packed", d["e2e_packed"], "frac", d["roofline"]["frac"])
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"], 1), "e2e", round(v["e2e"]["value"], 1), "e2e ms", round(v["e2e"]["ms_per_step"], 2), "packed", v["e2e_packed"] and (round(v["e2e_packed"]["value"], 1), round(v["e2e_packed"]["ms_per_step"], 2), v["e2e_packed"]["h2d_bytes_per_step"]))
PY
