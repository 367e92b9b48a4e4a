show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], json.dumps(d["config"]["phase_ms_per_step"]), "e2e", d["e2e"]["value"], "ok", (d["config"]["full_size_checks"] or {}).get("ok"))'
timeout 900 python -m pytest tests/test_gpu_oasis.py tests/test_gpu_updates.py tests/test_golden.py tests/test_golden_c1.py -q -m gpu -x 2>&1 | tail -4
echo "== profile"
CNMFE_HALS_PROFILE=1 timeout 280 python bench.py --steps 1 --warmup 1 --no-cpu --no-oracle-checks 2>gpurun_out/r2k_prof.err | python -c "$show"
grep "cnmfe hals" gpurun_out/r2k_prof.err | tail -3 | cut -c1-700
echo "== default"
timeout 280 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "$show"
