CNMFE_HALS_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-oracle-checks 2>&1 | grep "hals profile" | tail -4 | cut -c1-1500
