O=gpurun_out/n8_r02b; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.log 2> $O/bench_n8.err
tail -n 3 $O/bench_n8.err | cut -c1-300
python - $O <<'PY'
import json, sys
d = json.loads(open(sys.argv[1] + "/bench_n8.log").read().strip().splitlines()[-1])
print("cfg2 N=8 value", round(d["value"]), "e2e(ops)", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 2), "strings", round(d["e2e_strings"]["ms_per_step"], 2), "packed", round(d["e2e_packed"]["ms_per_step"], 2))
print("inprocess", d.get("e2e_inprocess"))
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"], 1), "e2e", round(v["e2e"]["value"], 1), "e2e ms", round(v["e2e"]["ms_per_step"], 2), "packed", v.get("e2e_packed") and (round(v["e2e_packed"]["value"], 1), round(v["e2e_packed"]["ms_per_step"], 2)))
PY
