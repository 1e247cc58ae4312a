O=gpurun_out/r02j; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fine_wavefront or config1 or cfg5_small" 2>&1 | tail -n 12
for w in 4 8 16 32; do BG_FINE_WARPS=$w timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fine_wavefront" 2>&1 | tail -n 1; done
for w in 0 4 8 16; do BG_FINE_WARPS=$w timeout 120 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('W=$w cfg1', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['phases_ms_last_step'], 'e2e ms', round(d['e2e']['ms_per_step'],2), 'frac', d['roofline']['frac'])"; done
BG_FINE_PAIRS=0 timeout 120 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('K2 cfg1', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['phases_ms_last_step'])"
