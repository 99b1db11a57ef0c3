for i in 1 2; do
  echo -n "base: " >> gpurun_out/r2_t35_ab.log; DIA_B200_LIB=$PWD/tools/ab/base.so timeout 300 python tools/batch_bench.py --utts 8 --reps 3 2>&1 | tail -1 >> gpurun_out/r2_t35_ab.log
  echo -n "new : " >> gpurun_out/r2_t35_ab.log; timeout 300 python tools/batch_bench.py --utts 8 --reps 3 2>&1 | tail -1 >> gpurun_out/r2_t35_ab.log
done
