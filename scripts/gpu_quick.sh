# Development helper (GPU box): selected tests + a short bench.  usage: gpu_quick.sh "<pytest -k expr>"
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 -k "$1" 2>&1 | tail -12
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["config"]["phase_ms_per_step"], "e2e", d["e2e"]["value"], "value", d["value"], "roofline", d["roofline"]["frac"])'
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "$show"
