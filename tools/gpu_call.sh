set -x
(time timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2_t65_all.log 2>&1
(timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_t65_bench.json 2> gpurun_out/r2_t65_bench.err)
(timeout 600 python tools/batch_bench.py --utts 1 2 3 4 8 --reps 3 --profile 2>&1 | tail -10) > gpurun_out/r2_t65_bb.log 2>&1
