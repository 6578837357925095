timeout 900 python -m pytest tests/test_api_gpu.py tests/test_hash_driver.py -x -q 2>&1 | tail -15
