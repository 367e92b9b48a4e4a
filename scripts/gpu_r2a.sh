# Round 2, first evidence run: GPU tests, bench with the oracle spot checks of the full-size results.
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2a_tests.log
timeout 900 python bench.py --no-cpu 2>gpurun_out/r2a_bench.err | tail -1 > gpurun_out/r2a_bench_n1.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2a_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])
print(json.dumps(d['config']['phase_ms_per_step']))
print(json.dumps(d['config']['full_size_checks'], indent=1))
PY
tail -5 gpurun_out/r2a_bench.err
