# compute-sanitizer memcheck over the small parity tests (every kernel family once): out-of-bounds / misaligned accesses
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests -m gpu -x -q -k "doctests or integration_goldens or hamming or p_distance or cfg5_small or edge_lengths or test_edit_distance or error_behaviour" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
tail -n 12 gpurun_out/sanitize_memcheck.log
python tools/bench_k5.py 1000000 1000 2>&1 | tail -1
python tools/bench_k5.py 4000000 150 2>&1 | tail -1
