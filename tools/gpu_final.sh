#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, reference arm, our arm
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ref_final.log 2>gpurun_out/bench_ref_final.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref_final.log
echo "== ours"; timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "rc=$?"; cat gpurun_out/bench_final.log; tail -3 gpurun_out/bench_final.err
