python tools/diag_e2e.py cfg4 100000 0 1 ops 2>&1 | tail -n 1
python tools/diag_e2e.py cfg4 100000 0 1 strings 2>&1 | tail -n 1
python tools/diag_e2e.py cfg2 1000000 0 1 ops 2>&1 | tail -n 1
python -m pytest tests/test_gpu_parity.py -x -q -k "cfg4 or pipeline or planner or packed or sharding" 2>&1 | tail -n 3
