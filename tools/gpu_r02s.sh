O=gpurun_out/r02s; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "random_vs_oracle or cfg2 or closure or golden or doctest or planner or edge" > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
run() { echo "== $EXTRA $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', round(d['value']), d['phases_ms_last_step'], 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
(run A=1
for p in 0x00 0x01 0x04 0x05 0x11 0x15 0x45 0x51 0x54 0x55 0x50 0x44 0x14 0x41; do run BG_HBP_PIPES=$p; done) > $O/sweep.log 2>&1
cat $O/sweep.log
