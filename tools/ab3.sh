#!/bin/bash
for i in 1 2; do
  for v in base regs; do echo -n "$v: "; DIA_B200_LIB=$PWD/tools/ab/$v.so python tools/stress.py --reps 12 --steps 64 | tail -1; done
  echo -n "new : "; python tools/stress.py --reps 12 --steps 64 | tail -1
done
