O=gpurun_out/r02k; mkdir -p $O
for ch in 3 4 5 6; do BG_PIPE_CHUNKS=$ch python tools/diag_e2e.py cfg2 1000000 0 1 ops 2>&1 | tail -n 1; done
python tools/diag_e2e.py cfg2 1000000 0 1 strings 2>&1 | tail -n 1
python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --shape 8,19 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('8,19', round(d['value']), d['phases_ms_last_step'])"
python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --shape 16,16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('16,16', round(d['value']), d['phases_ms_last_step'])"
python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('auto', round(d['value']), d['phases_ms_last_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e_strings']['ms_per_step'])"
for mb in 1024 2048 4096; do BG_TRACE_BUDGET_MB=$mb python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('budget $mb cfg2', round(d['value']), round(d['ms_per_step'],2), d['phases_ms_last_step'], d['roofline']['launches_per_step'])"; done
for mb in 2048 4096; do BG_TRACE_BUDGET_MB=$mb python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('budget $mb cfg4', round(d['value']), round(d['ms_per_step'],2), d['phases_ms_last_step'], d['roofline']['launches_per_step'])"; done
