#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== tune" ; timeout 600 python tools/tune.py --tag v2 > gpurun_out/tune_v2.log 2>&1; cat gpurun_out/tune_v2.log
echo "== bench" ; timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
