timeout 900 python -m pytest tests/test_golden_c1.py tests/test_gpu_oasis.py -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --workload oasis --steps 3 --warmup 3 2>gpurun_out/r2j_oasis.err | tail -1 > gpurun_out/r2j_bench_oasis.json
python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_oasis.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['config']['parity'], d['gpu_launches'], d['clocks'])"
tail -3 gpurun_out/r2j_oasis.err
