show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], json.dumps(d["config"]["phase_ms_per_step"]), "e2e", d["e2e"]["value"], "ok", (d["config"]["full_size_checks"] or {}).get("ok"), d["roofline"]["kernel"][:30], d["roofline"]["frac"])'
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
echo "== default"
timeout 280 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "$show"
echo "== no proj cache"
CNMFE_NO_PROJ_CACHE=1 timeout 280 python bench.py --steps 3 --warmup 3 --no-cpu --no-oracle-checks 2>/dev/null | python -c "$show"
