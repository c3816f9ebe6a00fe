#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2e_pytest.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "rc=$?"; tail -3 gpurun_out/r2e_bench.err; cat gpurun_out/r2e_bench.json
