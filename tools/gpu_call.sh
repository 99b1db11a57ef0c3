(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch" 2>&1 | tail -5) > gpurun_out/r2_t31_batch_tests.log 2>&1
(timeout 600 python tools/batch_bench.py --utts 1 2 4 8 --profile --reps 2 2>&1 | tail -20) > gpurun_out/r2_t31_bb_prof.log 2>&1
(DIA_BATCH_NO_MULTICAST=1 timeout 600 python tools/batch_bench.py --utts 8 --reps 2 2>&1 | tail -3) > gpurun_out/r2_t31_bb_nomc.log 2>&1
