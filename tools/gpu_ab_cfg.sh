#!/bin/bash
# A/B of two library builds on several configurations: gpu_ab_cfg.sh <dirA> <dirB> cfg...
A=$1; B=$2; shift 2
mkdir -p gpurun_out
for cfg in "$@"; do for d in $A $B; do
  python tools/tune.py --lib synchropmu_b200/$d/libpmu_estimator.so --tag $d --cfg $cfg --quick > gpurun_out/abc_${d}_$cfg.jsonl 2>> gpurun_out/abc.err
done; done
