# Development helper (GPU box): ring solve parity tests (with a hang guard) + bench + ring phase profile
timeout 300 python -m pytest tests -m gpu -x -q -k "updates or ssub or golden or ring" 2>&1 | tail -3
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["config"]["phase_ms_per_step"])'
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "$show"
CNMFE_RING_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu 2>&1 >/dev/null | grep "ring profile" | tail -2
