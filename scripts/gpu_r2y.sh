# default bench line of the final commit (CPU baseline + oracle checks)
timeout 900 python bench.py 2>gpurun_out/r2_bench.err | tail -1 > gpurun_out/r2_bench_n1.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
print(json.dumps(d['config']['phase_ms_per_step']))
print(d['config']['full_size_checks']['ok'], d['config']['full_size_checks']['oracle']['ok'], json.dumps(d['cpu_baseline']))
PY
