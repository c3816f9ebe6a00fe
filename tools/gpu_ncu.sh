#!/bin/bash
# ncu evidence for the fused kernel: launch list of the bench's timed region + one full capture
mkdir -p gpurun_out
TAG=${1:-r1b}
SHORT="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $SHORT > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pmu_fused -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmu_fused -s 4 -c 2 -o gpurun_out/prof_$TAG -f $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full_$TAG.log
