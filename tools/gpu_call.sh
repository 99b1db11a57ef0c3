set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "batched_decode_step or generate_batch_equals" 2>&1 | tail -40 > gpurun_out/r2_t4_batch_tiny.log; tail -25 gpurun_out/r2_t4_batch_tiny.log
timeout 300 python tools/batch_bench.py --tiny --steps 16 --slot 100 > gpurun_out/r2_t4_bb_tiny.log 2>&1; tail -6 gpurun_out/r2_t4_bb_tiny.log
timeout 600 python tools/batch_bench.py > gpurun_out/r2_t4_bb_full.log 2>&1; tail -6 gpurun_out/r2_t4_bb_full.log
timeout 600 bash tools/ab_tree.sh > gpurun_out/r2_t4_ab.log 2>&1; tail -4 gpurun_out/r2_t4_ab.log
timeout 1500 python -m pytest tests -m gpu -x -q -k "full and not batch" 2>&1 | tail -30 > gpurun_out/r2_t4_full.log; tail -12 gpurun_out/r2_t4_full.log
