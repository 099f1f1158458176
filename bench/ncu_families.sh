#!/bin/bash
# ncu --set full captures, one kernel family per run (bench/profile_kernels.py).  Only the raw CSV page of each report is kept
# (gpurun brings back at most 64 MiB): <outdir>/ncu_<family>_raw.csv
# Usage (under gpurun): bash bench/ncu_families.sh <outdir> <family>:<kernel regex>:<max launches> ...
out=$1; shift
mkdir -p $out
for spec in "$@"; do
  fam=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; cnt=${rest##*:}
  timeout 600 ncu --set full --clock-control none -k regex:"$rx" -c $cnt -f -o /tmp/ncu_$fam python bench/profile_kernels.py $fam > $out/ncu_$fam.log 2>&1
  ncu -i /tmp/ncu_$fam.ncu-rep --page raw --csv > $out/ncu_${fam}_raw.csv 2>/dev/null
  rm -f /tmp/ncu_$fam.ncu-rep
done
