# N-GPU evidence: default weak-scaling bench + C3, N given
N=${1:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --no-cpu 2>gpurun_out/r2d_bench_n$N.err | tail -1 > gpurun_out/r2d_bench_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --workload c3 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r2d_c3_n$N.err | tail -1 > gpurun_out/r2d_c3_n$N.json
python - <<PY
import json
for f in ('gpurun_out/r2d_bench_n$N.json', 'gpurun_out/r2d_c3_n$N.json'):
    try:
        d = json.load(open(f))
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'])
        print(json.dumps(d['config']['phase_ms_per_step']), d['config'].get('call_wall_ms_per_step'), d['config']['patches'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 3 gpurun_out/r2d_bench_n$N.err gpurun_out/r2d_c3_n$N.err
