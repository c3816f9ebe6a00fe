#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== tune" ; timeout 600 python tools/tune.py --tag v10 > gpurun_out/tune_v10.log 2>&1; grep -vE "post_slots" gpurun_out/tune_v10.log
for ARGS in "--workload config3" "--workload config3 --dtype f32" "--workload config5"; do
echo "== bench $ARGS"; timeout 600 python bench.py $ARGS --no-cpu-baseline --no-e2e > gpurun_out/b.log 2> gpurun_out/b.err; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['single_frame_launch']['value'], d['clocks'])"
done
