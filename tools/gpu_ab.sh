#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
for v in main NOPITCH ALWAYS_MASK GENERIC_COL main; do
  if [ $v = main ]; then L=synchropmu_b200/lib/libpmu_estimator.so; else L=synchropmu_b200/lib_$v/libpmu_estimator.so; fi
  echo "== $v"; timeout 300 python tools/tune.py --quick --tag $v --lib $L 2>&1 | grep -E "fit_fixed|\"n\": 24576|rror" | tee -a gpurun_out/ab.log
done
