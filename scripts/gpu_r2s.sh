# Multi-GPU diagnostic: per-rank wall / phase times and the host profile of the e2e path at N ranks
N=${1:-4}
CNMFE_HOST_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/r2s_n$N.out 2> gpurun_out/r2s_n$N.err
tail -1 gpurun_out/r2s_n$N.out > gpurun_out/r2s_bench_n$N.json
python - <<PY
import json
d = json.load(open('gpurun_out/r2s_bench_n$N.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'])
print(json.dumps(d['config']['phase_ms_per_step']))
for r in d['config'].get('per_rank') or []:
    print(r['rank'], [round(x,1) for x in r['call_wall_ms']], [round(x,1) for x in r['phase_ms']], round(r['device_ms'],1), round(r['wall_ms'],1))
PY
grep -h "e2e host profile\|cnmfe host\|\[host" gpurun_out/r2s_n$N.err | head -40
