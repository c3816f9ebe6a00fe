#!/bin/bash
# ncu --set full capture of one launch of the fused kernel per named r2_sweep case, exported to CSV on the GPU box
# (the .ncu-rep files are too big to bring several back): tools/gpu_ncu_cases.sh <tag> case...
mkdir -p gpurun_out
TAG=$1; shift
for c in "$@"; do
  rep=/tmp/prof_${TAG}_$c
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pmu_fused -s 6 -c 1 -o $rep -f python tools/r2_sweep.py --cases $c --steps 16 > gpurun_out/ncu_${TAG}_$c.log 2>&1
  echo "$c rc=$?"
  ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_${c}_raw.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_${TAG}_${c}_source.csv.gz
done
