# ncu launch list of the bench command (our kernels only), final code
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:ring_|proj_|hals_|deconv_|spatial_|temporal_|small_gram|gather_|transpose_|row_sum|add_small' -c 400 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-oracle-checks > gpurun_out/r2_ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches.csv
