(timeout 300 python tools/debug_batch.py --utts 1 3 --steps 2 2>&1 | tail -8) > gpurun_out/r2_t71.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch" 2>&1 | tail -8) >> gpurun_out/r2_t71.log 2>&1
for i in 1 2; do
  echo -n "base: " >> gpurun_out/r2_t71.log; DIA_B200_LIB=$PWD/tools/ab/base.so timeout 300 python tools/batch_bench.py --utts 8 --reps 3 2>&1 | tail -1 >> gpurun_out/r2_t71.log
  echo -n "new : " >> gpurun_out/r2_t71.log; timeout 300 python tools/batch_bench.py --utts 8 --reps 3 2>&1 | tail -1 >> gpurun_out/r2_t71.log
done
timeout 300 python tools/batch_bench.py --utts 8 --reps 2 --profile 2>&1 | tail -2 >> gpurun_out/r2_t71.log
