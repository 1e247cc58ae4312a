# Full GPU validation pass of a round: parity tests, smoke, benches of every config, ncu launch list and captures.
# The .ncu-rep files are turned into raw CSV pages on the box and deleted (gpurun_out may carry 64 MiB back).
O=gpurun_out/r01d; mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/pytest.log 2>&1
python __graft_entry__.py smoke > $O/smoke.log 2>&1
python bench.py > $O/bench_cfg2.log 2>&1
python bench.py --workload cfg4 --steps 5 --warmup 3 > $O/bench_cfg4.log 2>&1
python bench.py --workload cfg3 --steps 5 --warmup 3 > $O/bench_cfg3.log 2>&1
python bench.py --workload cfg5 --steps 3 --warmup 1 > $O/bench_cfg5.log 2>&1
python bench.py --workload cfg1 --steps 10 --warmup 3 > $O/bench_cfg1.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.log 2>&1
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1
cap() {  # name, workload, pairs, kernel regex, skip, count
  python tools/ncu_capture.py $2 $3 > $O/cap_$1_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$4" -s $5 -c $6 -f -o $O/$1 python tools/ncu_capture.py $2 $3 > $O/cap_$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2> /dev/null
  rm -f $O/$1.ncu-rep
}
cap k1h_fill_cfg2 cfg2 200000 "k1h_fill|k3_walk|k_gather" 3 3
cap k1_fill_local_cfg4u cfg4u 20000 "k1_fill" 1 1
cap k2_wave_cfg5 cfg5 32 "k2_wave|k3_walk_skew" 2 2
tail -n 3 $O/pytest.log $O/smoke.log $O/cap_*plain.log
ls -la $O
