# quick A/B: OASIS/update tests + bench + HALS profile
timeout 600 python -m pytest tests/test_gpu_oasis.py tests/test_gpu_updates.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>gpurun_out/r2x_bench.err | tail -1 > gpurun_out/r2x_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2x_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'])
print(json.dumps(d['config']['phase_ms_per_step']))
print(d['config']['full_size_checks']['ok'], json.dumps(d['config']['full_size_checks'].get('oracle', {}).get('deconvTemporal_vs_oracle')))
PY
CNMFE_HALS_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-oracle-checks 2>&1 | grep "hals profile" | tail -3 | cut -c1-900
