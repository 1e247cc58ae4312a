O=gpurun_out/r01e; mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/pytest.log 2>&1
python __graft_entry__.py smoke > $O/smoke.log 2>&1
python bench.py > $O/bench_cfg2.log 2>&1
python bench.py --workload cfg4 --steps 5 --warmup 3 > $O/bench_cfg4.log 2>&1
python bench.py --workload cfg3 --steps 5 --warmup 3 > $O/bench_cfg3.log 2>&1
python bench.py --workload cfg5 --steps 3 --warmup 1 > $O/bench_cfg5.log 2>&1
python bench.py --workload cfg1 --steps 10 --warmup 3 > $O/bench_cfg1.log 2>&1
head -3 $O/pytest.log; tail -n 1 $O/smoke.log
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do tail -n 1 $O/bench_$c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$c', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['phases_ms_last_step'], 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])"; done
