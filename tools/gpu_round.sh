# Full GPU validation pass of a round: parity tests, smoke, benches of every config, ncu launch list and captures.
O=gpurun_out/r01b; mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/pytest.log 2>&1
python __graft_entry__.py smoke > $O/smoke.log 2>&1
python bench.py > $O/bench_cfg2.log 2>&1
python bench.py --workload cfg4 --steps 5 --warmup 3 > $O/bench_cfg4.log 2>&1
python bench.py --workload cfg3 --steps 5 --warmup 3 > $O/bench_cfg3.log 2>&1
python bench.py --workload cfg5 --steps 3 --warmup 1 > $O/bench_cfg5.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.log 2>&1
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1
# full captures of the dominant kernels
python tools/ncu_capture.py cfg2 200000 > $O/cap_cfg2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1h_fill -s 1 -c 1 -f -o $O/k1h_fill_r01b python tools/ncu_capture.py cfg2 200000 > $O/cap_cfg2.log 2>&1
python tools/ncu_capture.py cfg4 30000 > $O/cap_cfg4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_fill -s 8 -c 8 -f -o $O/k1_fill_cfg4_r01b python tools/ncu_capture.py cfg4 30000 > $O/cap_cfg4.log 2>&1
python tools/ncu_capture.py cfg5 40 > $O/cap_cfg5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k2_wave|k3_walk_diag" -s 2 -c 2 -f -o $O/k2_wave_r01b python tools/ncu_capture.py cfg5 40 > $O/cap_cfg5.log 2>&1
tail -n 3 $O/pytest.log $O/smoke.log $O/cap_*plain.log
ls -la $O
