for i in 1 2; do
  echo -n "base: " >> gpurun_out/r2_t67_ab.log; DIA_B200_LIB=$PWD/tools/ab/base.so timeout 300 python tools/batch_bench.py --utts 8 --reps 3 2>&1 | tail -1 >> gpurun_out/r2_t67_ab.log
  echo -n "new : " >> gpurun_out/r2_t67_ab.log; timeout 300 python tools/batch_bench.py --utts 8 --reps 3 2>&1 | tail -1 >> gpurun_out/r2_t67_ab.log
done
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch" 2>&1 | tail -3) >> gpurun_out/r2_t67_ab.log 2>&1
timeout 600 python tools/batch_determinism.py --reps 10 2>&1 | tail -2 >> gpurun_out/r2_t67_ab.log
