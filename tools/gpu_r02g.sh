O=gpurun_out/r02g; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "packed or compact or sharding or cfg2" 2>&1 | tail -n 4
(time python bench.py --steps 10 --warmup 3 --no-configs) > $O/bench.log 2> $O/bench.err; tail -n 3 $O/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02g/bench.log").read().strip().splitlines()[-1])
print("cfg2", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], "strings", d["e2e_strings"], "packed", d["e2e_packed"]["ms_per_step"])
PY
BG_PROFILE_HOST=1 python tools/diag_e2e.py cfg2 1000000 0 4 > $O/e2e_profile_4logical.log 2>&1
python tools/diag_e2e.py cfg2 1000000 0 4 2>&1 | tail -n 1
python tools/diag_e2e.py cfg2 4000000 0 4 2>&1 | tail -n 1
