# Development helper (GPU box): full GPU tests + bench + host-side section profile
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["config"]["phase_ms_per_step"], d["config"].get("call_wall_ms_per_step"), "e2e", d["e2e"]["value"], "value", d["value"])'
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "$show"
CNMFE_HOST_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu 2>&1 >/dev/null | grep "cnmfe host" | tail -19
