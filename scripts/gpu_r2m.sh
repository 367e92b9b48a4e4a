timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu 2>gpurun_out/r2m_c4.err | tail -1 > gpurun_out/r2m_bench_c4.json
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2m_bench_c4.json'))
    print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel'][:40], d['roofline']['frac'], d['roofline'].get('als_iterations'))
    print(json.dumps(d['config']['phase_ms_per_step']), d['config']['full_size_checks'])
except Exception as e:
    print('FAILED', e)
PY
grep -v "Warning\|A_t = " gpurun_out/r2m_c4.err | tail -5
