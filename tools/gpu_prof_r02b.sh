# ncu evidence after the second half of round 2 (K1h byte profiles + (8,19), branch-free walker, K2 one-warp CTAs, K0e):
# launch list of the default bench command + full captures of the changed kernels.  Raw pages are exported as CSV here
# (the .ncu-rep files are too large to travel back).
O=gpurun_out/prof_r02b; mkdir -p $O
python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > $O/ncu_launches.log 2>&1
cap() {  # name, script, args, kernel regex, skip, count
  python $2 $3 > $O/cap_$1_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$4" -s $5 -c $6 -f -o $O/$1 python $2 $3 > $O/cap_$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2> /dev/null
  rm -f $O/$1.ncu-rep
}
cap k1h_fill_cfg2 tools/ncu_capture.py "cfg2 200000" "k1h_fill|k3_walk|k_gather" 3 3
cap k2_wave_cfg1 tools/ncu_capture.py "cfg1 1" "k2_wave|k3_walk_skew" 2 2
cap k4_myers_cfg3 tools/ncu_capture.py "cfg3 300000" "k4_myers" 3 3
cap k0e_cfg3e2e tools/ncu_capture.py "cfg3e2e 100000" "k_eplan|k_unpack" 0 12
tail -n 2 $O/cap_*plain.log
ls -la $O
