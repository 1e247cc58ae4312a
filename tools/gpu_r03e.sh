for rd in 0 1; do for ch in 3 4 6; do
echo "== ramp_down $rd chunks $ch"; BG_RAMP_DOWN=$rd BG_PIPE_CHUNKS=$ch python tools/diag_e2e.py cfg2 1000000 0 1 ops 2>&1 | tail -n 1
done; done
