timeout 900 python tools/batch_determinism.py --reps 16 2>&1 | tail -8 >> gpurun_out/r2_t52.log
timeout 900 python tools/batch_determinism.py --reps 6 --utts 5 --tokens 500 2>&1 | tail -4 >> gpurun_out/r2_t52.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch" 2>&1 | tail -3) >> gpurun_out/r2_t52.log 2>&1
timeout 300 python tools/batch_bench.py --utts 1 2 3 4 8 --reps 3 --profile 2>&1 | tail -10 >> gpurun_out/r2_t52.log
