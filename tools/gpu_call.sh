(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05 or dense or prefill or encoder or clone" 2>&1 | tail -5) > gpurun_out/r2_t36_gemm_tests.log 2>&1
(timeout 300 python tools/gemm_check.py 2>&1 | tail -12) > gpurun_out/r2_t36_gemm_check.log 2>&1
