#!/bin/bash
# A/B of estimator build variants (directories synchropmu_b200/lib_<tag>)
mkdir -p gpurun_out
CASES=${CASES:-c3,c3_1f,c3_trig,c3_f32,c5,c5_trig,mc,mc_trig}
for t in "$@"; do
  lib=synchropmu_b200/lib_$t/libpmu_estimator.so; [ "$t" = main ] && lib=synchropmu_b200/lib/libpmu_estimator.so
  echo "== $t"; timeout 600 python tools/r2_sweep.py --tag $t --lib $lib --cases $CASES --env "${ENVS:-}" > gpurun_out/r2b_$t.jsonl 2> gpurun_out/r2b_$t.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2b_$t.jsonl; tail -2 gpurun_out/r2b_$t.err
done
