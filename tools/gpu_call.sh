(timeout 300 python tools/debug_batch.py 2>&1 | tail -40) > gpurun_out/r2_t25_debug.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch" 2>&1 | tail -25) > gpurun_out/r2_t25_batch_tests.log 2>&1
