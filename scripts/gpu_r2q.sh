# CTA-pair (cta_group::2) second-moment kernel: parity against the exact SIMT kernel, then A/B against the one-CTA kernel.
export CNMFE_TC_MODE=pair
timeout 240 python -m pytest tests/test_gpu_ring_tc.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2q_tests.log
if grep -q "passed" gpurun_out/r2q_tests.log && ! grep -q "failed" gpurun_out/r2q_tests.log; then
  for m in pair single; do
    CNMFE_TC_MODE=$m timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-oracle-checks 2>/dev/null | tail -1 > gpurun_out/r2q_$m.json
    python - <<PY
import json
d = json.load(open('gpurun_out/r2q_$m.json'))
print('$m', d['ms_per_step'], json.dumps(d['config']['phase_ms_per_step']), d['clocks'])
PY
  done
fi
