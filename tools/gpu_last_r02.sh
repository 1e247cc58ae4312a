O=gpurun_out/last_r02; mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/pytest.log 2>&1
python __graft_entry__.py smoke > $O/smoke.log 2>&1
python bench.py > $O/bench.log 2> $O/bench.err
tail -n 4 $O/pytest.log; tail -n 1 $O/smoke.log
tail -n 1 $O/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cfg2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'strings', round(d['e2e_strings']['ms_per_step'],2), 'packed', round(d['e2e_packed']['ms_per_step'],2), d['phases_ms_last_step'], 'frac', round(d['roofline']['frac'],3), d['clocks'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), 'ms', round(v['ms_per_step'],2), 'e2e', round(v['e2e']['value'],1), round(v['e2e']['ms_per_step'],2), 'packed', (round(v['e2e_packed']['ms_per_step'],2) if v.get('e2e_packed') else None))"
