# Round 2: multi-GPU parity + sharded deconv + lazy ring mirror + tensor gram for kf > 1 (run with gpurun --gpus 2)
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/r2b_tests.log
timeout 600 python bench.py --no-cpu --no-oracle-checks 2>gpurun_out/r2b_bench1.err | tail -1 > gpurun_out/r2b_bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu 2>gpurun_out/r2b_bench2.err | tail -1 > gpurun_out/r2b_bench_n2.json
python - <<'PY'
import json
for f in ('gpurun_out/r2b_bench_n1.json', 'gpurun_out/r2b_bench_n2.json'):
    try:
        d = json.load(open(f))
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], 'e2e', d['e2e'], d['gpu_launches'])
        print(json.dumps(d['config']['phase_ms_per_step']), d['config'].get('call_wall_ms_per_step'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -3 gpurun_out/r2b_bench1.err gpurun_out/r2b_bench2.err
