O=gpurun_out/r02c; mkdir -p $O
(time python -m pytest tests/test_gpu_parity.py -x -q -k "planner or compact or sharding or cfg2 or pipeline or random_vs_oracle") > $O/pytest.log 2>&1
tail -n 15 $O/pytest.log
python tools/diag_e2e.py cfg2 1000000 0 > $O/e2e_devplan.log 2>&1
python tools/diag_e2e.py cfg2 1000000 1 > $O/e2e_hostplan.log 2>&1
python tools/diag_e2e.py cfg2 1000000 0 4 > $O/e2e_devplan_4logical.log 2>&1
BG_PROFILE_HOST=1 python tools/diag_e2e.py cfg2 1000000 0 > $O/e2e_profile.log 2>&1
tail -n 1 $O/e2e_devplan.log $O/e2e_hostplan.log $O/e2e_devplan_4logical.log
python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.log 2>&1; tail -c 1200 $O/bench_cfg2.log
