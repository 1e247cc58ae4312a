O=gpurun_out/r02t; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "wavefront or cfg1 or bounded or cfg5_small or long" > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
run() { echo "== $*"; env "$@" python bench.py --workload cfg1 --steps 10 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1', round(d['value'],1), d['phases_ms_last_step'], 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
(run A=1
for p in 16 8 4 2 1; do run BG_K2_WPC=$p; done) > $O/sweep.log 2>&1
cat $O/sweep.log
python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline > $O/bench_cfg2.log 2>&1; tail -n 1 $O/bench_cfg2.log | cut -c 1-600
