mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -x -q -m gpu -k "dropin or reference_main or reference_time or seam" 2>&1 | tail -15 > gpurun_out/r2_t10.log
cat gpurun_out/r2_t10.log
