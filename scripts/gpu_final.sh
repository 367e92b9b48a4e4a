# Round-end evidence run (GPU box): full GPU test suite, smoke, the default bench line (with CPU baseline).
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r1e_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>gpurun_out/r1e_bench.err | tail -1 > gpurun_out/r1e_bench_n1.json
python -c "import json; d=json.load(open('gpurun_out/r1e_bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'], d['gpu_launches'], d['clocks'], d['config']['phase_ms_per_step'])"
