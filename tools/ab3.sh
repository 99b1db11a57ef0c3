#!/bin/bash
# A/B/C timing on ONE box: every library in tools/ab/*.so, alternating (LIBS overrides the list).
for i in 1 2; do
  for v in ${LIBS:-$(ls tools/ab/*.so)}; do echo -n "$(basename $v .so): "; DIA_B200_LIB=$PWD/$v python tools/stress.py --reps ${REPS:-12} --steps 64 --slot ${SLOT:-1500} | tail -1; done
done
