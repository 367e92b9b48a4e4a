# weak-scaling bench line of the final commit at N GPUs (with per-rank diagnostics)
N=${1:-4}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --no-cpu 2>gpurun_out/r2z_bench_n$N.err | tail -1 > gpurun_out/r2z_bench_n$N.json
python - <<PY
import json
d = json.load(open('gpurun_out/r2z_bench_n$N.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'])
print(json.dumps(d['config']['phase_ms_per_step']))
for r in d['config'].get('per_rank') or []:
    print(r['rank'], [round(x,1) for x in r['call_wall_ms']], [round(x,1) for x in r['phase_ms']])
PY
