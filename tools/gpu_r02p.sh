python -m pytest tests/test_gpu_parity.py -x -q -k "random_vs_oracle or cfg2 or closure or pipeline or planner or packed or golden or doctest" 2>&1 | tail -n 3
python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', round(d['value']), d['phases_ms_last_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e_strings']['ms_per_step'], 'trace bytes', d['roofline']['hbm'])"
