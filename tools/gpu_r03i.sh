O=gpurun_out/r03i; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "wavefront or config1 or cfg1 or bounded or cfg5_small or long or goldens or multiband" > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
run() { echo "== $*"; env "$@" python bench.py --workload cfg1 --steps 10 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1', round(d['value'],1), d['phases_ms_last_step'], 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
run A=1
run BG_K2_NARROW=0
run BG_K2_NARROW=1 BG_K2_WPC=2
