set -x
(time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2_t12_all.log 2>&1; tail -30 gpurun_out/r2_t12_all.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_t12_bench.json 2> gpurun_out/r2_t12_bench.err; tail -c 3000 gpurun_out/r2_t12_bench.json; tail -5 gpurun_out/r2_t12_bench.err
