#!/bin/bash
# one compute-sanitizer tool per call (B200_PROFILING.md): memcheck over a few small parity cases
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_frame and 7 or stream_push and default_F3 or odd_window or three_phase and config.ini and f64" > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_frame and 7 or stream_push and default_F3 or odd_window or three_phase and config.ini and f64" > gpurun_out/sanitize_run.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/sanitize_run.log; tail -5 gpurun_out/memcheck.log
