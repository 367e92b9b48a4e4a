# final multi-GPU evidence at N GPUs: (N=2) GPU test suite incl. the multi-rank parity test; default bench and C3 at N
N=${1:-2}
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_updates.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2o_tests_n$N.log; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-cpu 2>gpurun_out/r2o_bench_n$N.err | tail -1 > gpurun_out/r2o_bench_n$N.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --workload c3 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r2o_c3_n$N.err | tail -1 > gpurun_out/r2o_c3_n$N.json
python - <<PY
import json
for f in ('gpurun_out/r2o_bench_n$N.json', 'gpurun_out/r2o_c3_n$N.json'):
    try:
        d = json.load(open(f))
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['gpu_launches'])
        print(json.dumps(d['config']['phase_ms_per_step']), d['config'].get('call_wall_ms_per_step'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
grep -v "Warning\|A_t = \|^W1017\|\*\*\*\|OMP_NUM" gpurun_out/r2o_bench_n$N.err | tail -3
