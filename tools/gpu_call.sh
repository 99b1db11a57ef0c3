(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch" 2>&1 | tail -5) > gpurun_out/r2_t58.log 2>&1
timeout 900 python tools/batch_determinism.py --reps 40 --tokens 400 2>&1 | tail -3 >> gpurun_out/r2_t58.log
