O=gpurun_out/r02e; mkdir -p $O
(time python bench.py --steps 10 --warmup 3) > $O/bench.log 2> $O/bench.err; tail -n 3 $O/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02e/bench.log").read().strip().splitlines()[-1])
print("cfg2", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["clocks"])
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"], 1), "e2e", round(v["e2e"]["value"], 1), "ms", round(v["ms_per_step"], 2), "e2e ms", round(v["e2e"]["ms_per_step"], 2), "frac", v["roofline"]["frac"], v["phases_ms_last_step"], v["clocks"]["sm_mhz"], v["clocks"]["samples"])
PY
