# Round-2 validation on one B200: the whole GPU suite, smoke, the default bench line (+ configs), the reference arm.
O=gpurun_out/final_r02; mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/pytest.log 2>&1
python __graft_entry__.py smoke > $O/smoke.log 2>&1
python bench.py > $O/bench.log 2> $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err
tail -n 6 $O/pytest.log; tail -n 1 $O/smoke.log
tail -n 1 $O/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cfg2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'strings', round(d['e2e_strings']['ms_per_step'],2), 'packed', round(d['e2e_packed']['ms_per_step'],2), d['phases_ms_last_step'], 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], d['clocks'])
print('cpu', d['cpu_baseline'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), 'ms', round(v['ms_per_step'],2), 'e2e', round(v['e2e']['value'],1), round(v['e2e']['ms_per_step'],2), 'packed', (round(v['e2e_packed']['ms_per_step'],2) if v.get('e2e_packed') else None), 'frac', round(v['roofline']['frac'],3) if v.get('roofline') else None)"
tail -n 1 $O/bench_ref.log | cut -c 1-400
python tools/diag_midlen.py > $O/midlen.log 2>&1; BG_NO_HALF_PROF=1 python tools/diag_midlen.py > $O/midlen_table.log 2>&1
cat $O/midlen.log
