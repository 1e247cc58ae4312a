python -m pytest tests/test_gpu_parity.py -x -q -k "random_vs_oracle or cfg2 or closure or golden or doctest or planner or edge" 2>&1 | tail -n 3
run() { echo "== $EXTRA $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', round(d['value']), d['phases_ms_last_step'], 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
run A=1
for s in 8,19 8,20 16,10; do
for p in 0x15 0x55 0x57 0x5F 0x7F 0xFF 0xF5 0xD5; do EXTRA="--shape $s" run BG_HBP_PIPES=$p; done
done
python tools/diag_midlen.py
BG_NO_HALF_PROF=1 python tools/diag_midlen.py
