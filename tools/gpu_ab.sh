#!/bin/bash
# A/B of two library builds with tools/tune.py: gpu_ab.sh <dirA> <dirB>  (directories under synchropmu_b200/)
set -u
A=${1:-lib_base}; B=${2:-lib}
mkdir -p gpurun_out
for rep in 1 2; do
for d in $A $B; do
  python tools/tune.py --lib synchropmu_b200/$d/libpmu_estimator.so --tag $d > gpurun_out/ab_${d}_$rep.jsonl 2> gpurun_out/ab_${d}_$rep.err
  python tools/tune.py --lib synchropmu_b200/$d/libpmu_estimator.so --tag $d --cfg cfg5_600 --quick > gpurun_out/ab5_${d}_$rep.jsonl 2>> gpurun_out/ab_${d}_$rep.err
done
done
