(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batched_long_context" -s 2>&1 | tail -12) > gpurun_out/r2_t61.log 2>&1
