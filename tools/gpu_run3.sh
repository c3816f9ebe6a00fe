#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (remaining)" ; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3 or config5 or pointers or checkpoint" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu2.log
echo "== tune 256" ; timeout 600 python tools/tune.py --tag t256 > gpurun_out/tune_256.log 2>&1; cat gpurun_out/tune_256.log
echo "== tune 384" ; timeout 600 python tools/tune.py --tag t384 --lib synchropmu_b200/lib384/libpmu_estimator.so > gpurun_out/tune_384.log 2>&1; cat gpurun_out/tune_384.log
