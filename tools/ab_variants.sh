#!/bin/bash
# A/B/C... timing on ONE box: every build-time variant of the in-tree sources in tools/ab/*.so (same ABI, same Python),
# alternating (LIBS overrides the list), plus the r1 tree as the fixed baseline.
for i in 1 2; do
  echo -n "r1_tree: "; (cd tools/ab/r1_tree && python tools/stress.py --reps ${REPS:-12} --steps 64 --slot ${SLOT:-1500} | tail -1)
  for v in ${LIBS:-$(ls tools/ab/*.so | grep -v r1.so)}; do echo -n "$(basename $v .so): "; DIA_B200_LIB=$PWD/$v python tools/stress.py --reps ${REPS:-12} --steps 64 --slot ${SLOT:-1500} | tail -1; done
done
