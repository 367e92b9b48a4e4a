# ncu --set full capture of one hals_temporal_kernel launch (source-level stall samples)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hals_temporal --launch-skip 1 --launch-count 1 \
  -o gpurun_out/r2_hals_full -f python bench.py --steps 1 --warmup 1 --no-cpu --no-oracle-checks > gpurun_out/r2v_ncu.log 2>&1
tail -3 gpurun_out/r2v_ncu.log
ls -la gpurun_out/r2_hals_full.ncu-rep
