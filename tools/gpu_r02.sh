# GPU pass of round 2: parity tests first (fail fast), then smoke + the default bench line.
O=gpurun_out/${1:-r02b}; mkdir -p $O
(time python -m pytest tests -m gpu -x -q ${PYTEST_ARGS}) > $O/pytest.log 2>&1
tail -n 25 $O/pytest.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.log 2>&1; tail -c 1500 $O/bench_cfg2.log
