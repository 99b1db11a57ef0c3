timeout 1500 python tools/batch_determinism.py --reps 200 --tokens 260 2>&1 | tail -3 > gpurun_out/r2_t63.log
timeout 900 python tools/batch_determinism.py --reps 30 --tokens 300 --utts 3 2>&1 | tail -2 >> gpurun_out/r2_t63.log
