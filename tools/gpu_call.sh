(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch" 2>&1 | tail -3) > gpurun_out/r2_t72.log 2>&1
timeout 600 python tools/batch_determinism.py --reps 8 2>&1 | tail -1 >> gpurun_out/r2_t72.log
timeout 300 python tools/batch_bench.py --utts 8 --reps 2 2>&1 | tail -1 >> gpurun_out/r2_t72.log
