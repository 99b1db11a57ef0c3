#!/bin/bash
# Per-kernel counts of the SASS opcodes that prove which hardware paths a kernel uses (B200_PROFILING.md):
#   UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = 1-D bulk copy (TMA engine),
#   HMMA = legacy mma.sync, HMMA.SP = mma.sp (2:4), SYNCS = mbarrier ops, FFMA = fp32 FMA
so=${1:-dia_tts_prune_b200/csrc/libdia_b200.so}
cuobjdump -sass "$so" | awk '
  /Function :/ { fn=$3 }
  { for (i=1;i<=NF;i++) {
      if ($i ~ /^UTCHMMA/) c[fn,"UTCHMMA"]++
      else if ($i ~ /^LDTM/) c[fn,"LDTM"]++
      else if ($i ~ /^UTMALDG/) c[fn,"UTMALDG"]++
      else if ($i ~ /^UBLKCP/) c[fn,"UBLKCP"]++
      else if ($i ~ /^HMMA\.SP/) c[fn,"HMMA.SP"]++
      else if ($i ~ /^HMMA/) c[fn,"HMMA"]++
      else if ($i ~ /^SYNCS/) c[fn,"SYNCS"]++
      else if ($i ~ /^FFMA/) c[fn,"FFMA"]++
      else if ($i ~ /^UTCBAR/) c[fn,"UTCBAR"]++
  } fns[fn]=1 }
  END { printf "%-64s %8s %6s %8s %7s %6s %8s %6s %7s %6s\n","kernel","UTCHMMA","LDTM","UTMALDG","UBLKCP","HMMA","HMMA.SP","SYNCS","UTCBAR","FFMA";
        for (f in fns) if (f != "") printf "%-64s %8d %6d %8d %7d %6d %8d %6d %7d %6d\n", substr(f,1,64), c[f,"UTCHMMA"], c[f,"LDTM"], c[f,"UTMALDG"], c[f,"UBLKCP"], c[f,"HMMA"], c[f,"HMMA.SP"], c[f,"SYNCS"], c[f,"UTCBAR"], c[f,"FFMA"] }' | sort
