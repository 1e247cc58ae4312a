O=gpurun_out/r03f; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "random_vs_oracle or cfg2 or closure or golden or doctest or planner or edge or pipeline or compact" > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
for tg in 2 1 0; do
BG_K1H_TG=$tg python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tg $tg cfg2', round(d['value']), d['phases_ms_last_step'], 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
python tools/diag_midlen.py
