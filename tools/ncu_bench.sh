#!/bin/bash
# Profiling recipe of this repo (run under gpurun on one B200); outputs land in gpurun_out/.
#  1. launch list of our kernels for the bench command (shares, not absolutes: ncu serialises and runs cold)
#  2. one `--set full` capture of a 64-step launch of the persistent step kernel (DRAM traffic, stall reasons)
set -x
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dia -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/stress.py --reps 2 --steps 64 > gpurun_out/plain_stress.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dia_step_kernel -s 1 -c 1 -o gpurun_out/step_r1_full \
    python tools/stress.py --reps 2 --steps 64 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
