python tools/lines_dev.py
timeout 900 python -m pytest tests/test_sheppard_gpu.py tests/test_hanser_big_gpu.py tests/test_api_gpu.py tests/test_pr_gpu.py tests/test_random_sweep_gpu.py tests/test_full_size_properties_gpu.py -x -q -m gpu 2>&1 | tail -6
