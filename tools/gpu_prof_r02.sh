# ncu evidence of round 2: launch list of the default bench command + full captures of the changed kernels.
O=gpurun_out/prof_r02; mkdir -p $O
python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > $O/ncu_launches.log 2>&1
cap() {  # name, script, args, kernel regex, skip, count
  python $2 $3 > $O/cap_$1_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$4" -s $5 -c $6 -f -o $O/$1 python $2 $3 > $O/cap_$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2> /dev/null
  rm -f $O/$1.ncu-rep
}
cap k1h_fill_cfg2 tools/ncu_capture.py "cfg2 200000" "k1h_fill|k3_walk|k_gather" 3 3
cap k1_fill_local_cfg4u tools/ncu_capture.py "cfg4u 20000" "k1_fill" 1 1
cap k2_wave_cfg5 tools/ncu_capture.py "cfg5 32" "k2_wave|k3_walk_skew" 2 2
cap e2e_small tools/ncu_capture_e2e.py "100000" "k_unpack|k_plan|k_ops|k_pack" 0 40
tail -n 2 $O/cap_*plain.log
ls -la $O
