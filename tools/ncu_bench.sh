#!/bin/bash
# Profiling recipe of this repo (run under gpurun on one B200); outputs land in gpurun_out/ (TAG defaults to r2).
#  1. launch list of our kernels for the bench command (shares, not absolutes: ncu serialises and runs cold)
#  2. one `--set full` capture of a 64-step launch of the persistent step kernel (DRAM traffic, stall reasons)
#  3. the same for a 16-step launch of the batched kernel with 8 utterances, and for the tcgen05 GEMM at M=2048, N=16384, K=2048
# Every profiled command first runs to completion WITHOUT ncu.
TAG=${TAG:-r2}
set -x
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch-utterances 8 > gpurun_out/plain_bench_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dia -c 900 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch-utterances 8 > gpurun_out/ncu_launches_$TAG.log 2>&1
python tools/stress.py --reps 2 --steps 64 > gpurun_out/plain_stress_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dia_step_kernel -s 1 -c 1 -f -o gpurun_out/step_${TAG}_full \
    python tools/stress.py --reps 2 --steps 64 > gpurun_out/ncu_full_$TAG.log 2>&1
python tools/batch_bench.py --utts 8 --steps 16 --reps 2 > gpurun_out/plain_batch_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dia_batch_step_kernel -s 1 -c 1 -f -o gpurun_out/batch_${TAG}_full \
    python tools/batch_bench.py --utts 8 --steps 16 --reps 2 > gpurun_out/ncu_batch_$TAG.log 2>&1
python tools/gemm_check.py > gpurun_out/plain_gemm_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dia_gemm_tcgen05 -s 24 -c 1 -f -o gpurun_out/gemm_${TAG}_full \
    python tools/gemm_check.py > gpurun_out/ncu_gemm_$TAG.log 2>&1
for f in gpurun_out/ncu_full_$TAG.log gpurun_out/ncu_batch_$TAG.log gpurun_out/ncu_gemm_$TAG.log; do tail -n 2 $f; done
