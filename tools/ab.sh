#!/bin/bash
# A/B timing on ONE box: tools/ab/base.so (a previous build) against the in-tree library, alternating.
for i in 1 2 3; do
  echo -n "base: "; DIA_B200_LIB=$PWD/tools/ab/base.so python tools/stress.py --reps ${REPS:-15} --steps 64 --slot ${SLOT:-1500} | tail -1
  echo -n "new : "; python tools/stress.py --reps ${REPS:-15} --steps 64 --slot ${SLOT:-1500} | tail -1
done
