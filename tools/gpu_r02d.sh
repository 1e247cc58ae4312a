O=gpurun_out/r02d; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "planner or compact or sharding or cfg2 or cfg4 or wavefront or bounded or edge" 2>&1 | tail -n 5
BG_PROFILE_HOST=1 python tools/diag_e2e.py cfg2 1000000 0 > $O/e2e_profile.log 2>&1
python tools/diag_e2e.py cfg2 1000000 0 2>&1 | tail -n 1
python tools/diag_e2e.py cfg4 100000 0 2>&1 | tail -n 1
