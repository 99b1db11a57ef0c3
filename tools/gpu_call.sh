set -x
(time timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2_t70_all.log 2>&1
python tools/batch_bench.py --utts 8 --steps 16 --reps 2 > gpurun_out/plain_batch_r2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dia_batch_step_kernel -s 1 -c 1 -f -o gpurun_out/batch_r2b_full \
    python tools/batch_bench.py --utts 8 --steps 16 --reps 2 > gpurun_out/ncu_batch_r2b.log 2>&1
tail -n 2 gpurun_out/ncu_batch_r2b.log
