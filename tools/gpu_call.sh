(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mixed_prompt_depths or batch" 2>&1 | tail -12) > gpurun_out/r2_t69.log 2>&1
