timeout 300 python tools/batch_bench.py --utts 8 --reps 2 --profile 2>&1 | tail -2 > gpurun_out/r2_t62.log
