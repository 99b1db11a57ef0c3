set -x
(time timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/r2_t21_all.log 2>&1; tail -14 gpurun_out/r2_t21_all.log
