O=gpurun_out/r02u; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "edit or cfg3 or packed or sharding" > $O/pytest.log 2>&1; tail -n 5 $O/pytest.log
run() { echo "== $*"; env "$@" python bench.py --workload cfg3 --steps 5 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3', round(d['value'],1), d['phases_ms_last_step'], 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']), 'packed', round(d['e2e_packed']['ms_per_step'],2), round(d['e2e_packed']['value']))"; }
(run A=1
for p in 4 6 8 16 24; do run BG_EDIT_CHUNKS_DEV=$p; done
run BG_HOST_PLAN=1) > $O/sweep.log 2>&1
cat $O/sweep.log
