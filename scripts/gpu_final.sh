# Round-end evidence run (GPU box): full GPU test suite, smoke, the default bench line (with CPU baseline), the reference
# arm, and a reduced OASIS-stress line.  Outputs under gpurun_out/.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r1d_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>gpurun_out/r1d_bench.err | tail -1 > gpurun_out/r1d_bench_n1.json
python -c "import json; d=json.load(open('gpurun_out/r1d_bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1d_bench_reference.json
cut -c1-300 gpurun_out/r1d_bench_reference.json
timeout 600 python bench.py --workload oasis --oasis-traces 1000 --steps 2 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r1d_bench_oasis.json
cat gpurun_out/r1d_bench_oasis.json
