# Round 2 final single-GPU evidence: full GPU suite, smoke(), default bench (CPU baseline + oracle checks at full size),
# the reference arm, and the ncu launch list of the same bench command.
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2_tests_n1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
timeout 900 python bench.py 2>gpurun_out/r2_bench.err | tail -1 > gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --impl reference 2>gpurun_out/r2_ref.err | tail -1 > gpurun_out/r2_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cnmfe -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-oracle-checks > gpurun_out/r2_ncu_bench.log 2>&1
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
print(json.dumps(d['config']['phase_ms_per_step']))
print(json.dumps(d['config']['full_size_checks'], indent=1)[:1500])
print(json.dumps(d['cpu_baseline']))
print(open('gpurun_out/r2_bench_reference.json').read()[:600])
PY
tail -3 gpurun_out/r2_bench.err
