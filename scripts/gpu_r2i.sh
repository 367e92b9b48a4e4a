# ncu: launch list of one bench step + full capture of the MMA ring solver and of the hals sweeps kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:cnmfe -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-oracle-checks > gpurun_out/r2_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ring_solve_mma -s 1 -c 1 -o gpurun_out/r2_solve_mma python bench.py --steps 1 --warmup 1 --no-cpu --no-oracle-checks > gpurun_out/r2_solve_ncu.log 2>&1
ls -la gpurun_out/r2_*
