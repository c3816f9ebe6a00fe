#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== tune" ; timeout 600 python tools/tune.py --tag v4 > gpurun_out/tune_v4.log 2>&1; grep -vE "post_slots" gpurun_out/tune_v4.log
echo "== tune384" ; timeout 600 python tools/tune.py --tag v4_384 --lib synchropmu_b200/lib384/libpmu_estimator.so > gpurun_out/tune_v4_384.log 2>&1; grep -E "fit|frames_per|24576" gpurun_out/tune_v4_384.log | grep -vE "post_slots|stages="
