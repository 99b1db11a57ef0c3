#!/bin/bash
# A/B timing on ONE box: a previous source tree + its library (tools/ab/<name>_tree, e.g. made with
# `git archive <rev> dia_tts_prune_b200 tools/stress.py include oracle | tar -x -C tools/ab/<name>_tree`) against the
# in-tree build, alternating.  BASE=<name> picks the tree (default r1).
base=tools/ab/${BASE:-r1}_tree
for i in 1 2 3; do
  echo -n "base: "; (cd $base && python tools/stress.py --reps ${REPS:-15} --steps 64 --slot ${SLOT:-1500} | tail -1)
  echo -n "new : "; python tools/stress.py --reps ${REPS:-15} --steps 64 --slot ${SLOT:-1500} | tail -1
done
