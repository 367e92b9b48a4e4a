# Is the CTA-pair second-moment kernel bound by the chip-level L2 throughput?  (time vs. SMs occupied); + AR2 thresholded/optimize_b test
timeout 240 python -m pytest tests/test_gpu_oasis.py -x -q -k "thresholded_ar2" 2>&1 | tail -3
for n in 124 108; do
  CNMFE_TC_MODE=pair CNMFE_TC_SMS=$n timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-oracle-checks 2>/dev/null | tail -1 > gpurun_out/r2r_pair_$n.json
  python - <<PY
import json
d = json.load(open('gpurun_out/r2r_pair_$n.json'))
print('pair TC_SMS', $n, d['ms_per_step'], json.dumps(d['config']['phase_ms_per_step']))
PY
done
