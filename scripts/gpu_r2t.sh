# OASIS latency work (two-level pow table, closed-form cumsum(h.^2), flat two-pass rss_g): full GPU suite + bench + HALS profile
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2t_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>gpurun_out/r2t_bench.err | tail -1 > gpurun_out/r2t_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2t_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
print(json.dumps(d['config']['phase_ms_per_step']))
print(d['config']['full_size_checks']['ok'], json.dumps(d['config']['full_size_checks'].get('oracle', {}).get('deconvTemporal_vs_oracle')))
PY
CNMFE_HALS_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-oracle-checks 2>&1 | grep "hals profile" | tail -3 | cut -c1-1200
