show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], json.dumps(d["config"]["phase_ms_per_step"]))'
for v in a b c; do
  echo "== variant $v"
  CNMFE_RING_PROFILE=1 CNMFE_B200_LIB=$PWD/cnmf_e_b200/libcnmfe_b200_$v.so timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-oracle-checks 2>gpurun_out/r2h_$v.err | python -c "$show"
  grep "cnmfe ring" gpurun_out/r2h_$v.err | tail -1 | cut -c1-400
  CNMFE_B200_LIB=$PWD/cnmf_e_b200/libcnmfe_b200_$v.so timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-oracle-checks 2>/dev/null | python -c "$show"
done
