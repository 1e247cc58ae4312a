python -m pytest tests -m gpu -x -q -k "bounded or wavefront or cfg5 or config1 or edge or goldens" 2>&1 | tail -8
python bench.py --workload cfg5 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['phases_ms_last_step'], d['roofline']['launches_per_step'])"
python bench.py --workload cfg1 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases_ms_last_step'])"
