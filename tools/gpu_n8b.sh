O=gpurun_out/n8b; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-configs > $O/bench_n8.log 2> $O/bench_n8.err
python - $O <<'PY'
import json, sys
d = json.loads(open(sys.argv[1] + "/bench_n8.log").read().strip().splitlines()[-1])
print("cfg2 N=8 value", round(d["value"]), "e2e(ops)", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 2), "strings", round(d["e2e_strings"]["ms_per_step"], 2), "packed", round(d["e2e_packed"]["ms_per_step"], 2))
PY
python tools/diag_e2e.py cfg2 8000000 0 8 ops real 2>&1 | tail -n 1
python tools/diag_e2e.py cfg2 8000000 0 8 strings real 2>&1 | tail -n 1
BG_PROFILE_HOST=1 python tools/diag_e2e.py cfg2 8000000 0 8 ops real > $O/prof_ops.log 2>&1
BG_HOST_THREADS=8 python tools/diag_e2e.py cfg2 8000000 0 8 ops real 2>&1 | tail -n 1
