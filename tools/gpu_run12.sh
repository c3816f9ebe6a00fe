#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q -k "stream" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for ARGS in "--workload config5 --dtype f32" "--workload config5"; do
echo "== bench $ARGS"; timeout 600 python bench.py $ARGS --no-cpu-baseline > gpurun_out/b.log 2> gpurun_out/b.err; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['e2e']['value'], d['e2e_stream'])"
done
