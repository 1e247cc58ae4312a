python -m pytest tests -m gpu -x -q -k "pipeline or cfg2_sample or cfg3 or cfg4 or multi or edit" 2>&1 | tail -3
for w in cfg2 cfg3 cfg4; do python bench.py --workload $w --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"; done
BG_PROFILE_HOST=1 python tools/diag_e2e.py 2>&1 | grep -v "gpu ws\|plan of\|issued\|results on" | tail -8
