for c in 6 12 24; do echo "edit chunks=$c"; BG_EDIT_CHUNKS=$c python bench.py --workload cfg3 --steps 6 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"; done
python -m pytest tests -m gpu -x -q -k "edit" 2>&1 | tail -3
