# C3 (configs[2]) on N GPUs: bash scripts/gpu_r2c.sh N
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c3 --steps 2 --warmup 3 --no-cpu 2>gpurun_out/r2c_c3_n$N.err | tail -1 > gpurun_out/r2c_c3_n$N.json
python - <<PY
import json
d = json.load(open('gpurun_out/r2c_c3_n$N.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel'], d['roofline']['frac'])
print(json.dumps(d['config']['phase_ms_per_step']), d['config'].get('call_wall_ms_per_step'), d['config']['patches'])
PY
tail -n 5 gpurun_out/r2c_c3_n$N.err
