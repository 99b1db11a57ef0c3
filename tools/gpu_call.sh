set -x
for v in bf0 bf2; do echo "== $v"; DIA_B200_LIB=$PWD/tools/ab/$v.so timeout 300 python tools/batch_bench.py --utts 1 8 --profile --reps 2 2>&1 | tail -4; done > gpurun_out/r2_t16_fence.log 2>&1; cat gpurun_out/r2_t16_fence.log
