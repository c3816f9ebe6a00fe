#!/bin/bash
# parity tests, bench, then ncu launch list + one full capture of the fused kernel
mkdir -p gpurun_out
echo "== pytest" ; timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 600 python bench.py --steps 2000 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
SHORT="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
echo "== ncu"
timeout 300 $SHORT > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmu_fused -s 3 -c 2 -o gpurun_out/prof_r1 -f $SHORT > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.log
