#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== tune" ; timeout 600 python tools/tune.py --tag v5 > gpurun_out/tune_v5.log 2>&1; grep -vE "post_slots|stages=|consumers=" gpurun_out/tune_v5.log
for WL in config3 config5 mclass; do
echo "== bench $WL f64"; timeout 600 python bench.py --workload $WL --no-cpu-baseline > gpurun_out/bench_$WL.log 2> gpurun_out/bench_$WL.err; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_$WL.log').read());print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['single_frame_launch'], d['e2e']['value'], d['clocks'])"
done
echo "== bench config3 f32"; timeout 600 python bench.py --dtype f32 --no-cpu-baseline > gpurun_out/bench_config3_f32.log 2> gpurun_out/bench_config3_f32.err; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_config3_f32.log').read());print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['single_frame_launch'], d['e2e']['value'])"
echo "== bench config3 interharmonic"; timeout 600 python bench.py --interharmonic --no-cpu-baseline --no-e2e > gpurun_out/bench_config3_trig.log 2> gpurun_out/bench_config3_trig.err; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_config3_trig.log').read());print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])"
