timeout 900 python -m pytest tests/test_gpu_rss.py tests/test_gpu_oasis.py -q -m gpu 2>&1 | tail -12
