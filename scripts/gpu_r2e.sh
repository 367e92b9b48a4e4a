# MMA ring solver validation + A/B, and CTA-size variants of the per-trace kernels (1 GPU)
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], json.dumps(d["config"]["phase_ms_per_step"]), "e2e", d["e2e"]["value"], "ok", d["config"]["full_size_checks"].get("ok"))'
timeout 600 python -m pytest tests/test_gpu_updates.py tests/test_golden.py -q -x 2>&1 | tail -4
echo "== mma + profile"
CNMFE_RING_PROFILE=1 timeout 280 python bench.py --steps 1 --warmup 1 --no-cpu --no-oracle-checks 2>gpurun_out/r2e_prof.err | python -c "$show"
grep "cnmfe ring" gpurun_out/r2e_prof.err | tail -3
echo "== mma"
timeout 280 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "$show"
echo "== simt"
CNMFE_RING_SOLVER=simt timeout 280 python bench.py --steps 3 --warmup 3 --no-cpu --no-oracle-checks 2>/dev/null | python -c "$show"
for v in b512 b1024; do
  echo "== variant $v"
  CNMFE_B200_LIB=$PWD/cnmf_e_b200/libcnmfe_b200_$v.so timeout 280 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "$show"
done
