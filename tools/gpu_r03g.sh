run() { echo "== $EXTRA $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', round(d['value']), d['phases_ms_last_step'], 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
python -m pytest tests/test_gpu_parity.py -x -q -k "cfg2_sample" 2>&1 | tail -n 1
run A=1
for p in 0x00 0x05 0x15 0x45 0x55; do run BG_HBP_PIPES=$p; done
