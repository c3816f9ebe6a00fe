#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== tune 256 noinline" ; timeout 600 python tools/tune.py --tag v9_256 > gpurun_out/tune_v9_256.log 2>&1; grep -vE "post_slots|stages=" gpurun_out/tune_v9_256.log
echo "== tune 384 noinline" ; timeout 600 python tools/tune.py --tag v9_384 --lib synchropmu_b200/lib384/libpmu_estimator.so > gpurun_out/tune_v9_384.log 2>&1; grep -vE "post_slots|stages=" gpurun_out/tune_v9_384.log
