timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | grep -E "^E|^tests.*Error|^FAILED|assert" | head -20
