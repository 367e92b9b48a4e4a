#!/bin/bash
# Development helper (GPU box): bench A/B variants of the library; prints the per-phase ms of each.
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["config"]["phase_ms_per_step"], "e2e", d["e2e"]["value"], "value", d["value"])'
echo "== base + profiles"
CNMFE_HALS_PROFILE=1 CNMFE_RING_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu 2>gpurun_out/ab_prof.err | python -c "$show"
grep "cnmfe" gpurun_out/ab_prof.err | tail -4
for v in "$@"; do
  echo "== variant $v"
  CNMFE_B200_LIB=$PWD/cnmf_e_b200/libcnmfe_b200_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "$show"
done
echo "== base"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "$show"
