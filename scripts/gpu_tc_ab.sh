# Development helper (GPU box): A/B of the tensor-core second-moment kernel variants.
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("gram ms", d["config"]["phase_ms_per_step"]["gram"], "step", d["ms_per_step"])'
for cfg in "2 1 1" "2 1 0" "2 0 1" "2 0 0" "1 1 1" "4 1 1"; do
  set -- $cfg
  echo "== cluster=$1 order=$2 epi=$3"
  CNMFE_TC_CLUSTER=$1 CNMFE_TC_ORDER=$2 CNMFE_TC_EPI=$3 timeout 180 python -m pytest tests/test_gpu_ring_tc.py -m gpu -x -q 2>&1 | tail -2
  CNMFE_TC_CLUSTER=$1 CNMFE_TC_ORDER=$2 CNMFE_TC_EPI=$3 timeout 180 python bench.py --steps 2 --warmup 2 --no-cpu 2>/dev/null | python -c "$show"
done
