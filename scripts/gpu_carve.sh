show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("ring_solve ms", d["config"]["phase_ms_per_step"]["ring_solve"])'
for c in default 25 40 50 60 75; do
  echo "== carveout $c"
  if [ $c = default ]; then timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu 2>/dev/null | python -c "$show"
  else CNMFE_RING_CARVEOUT=$c timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu 2>/dev/null | python -c "$show"; fi
done
