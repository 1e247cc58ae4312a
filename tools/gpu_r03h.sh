O=gpurun_out/r03h; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "planner or cfg2 or pipeline or compact or packed" > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
for v in 0 1; do
if [ $v = 1 ]; then export BG_NO_UNIFORM_PLAN=1; fi
echo "== no_uniform=$v"; python tools/diag_e2e.py cfg2 1000000 0 1 ops 2>&1 | tail -n 1
python tools/diag_e2e.py cfg2 1000000 0 1 strings 2>&1 | tail -n 1
done
