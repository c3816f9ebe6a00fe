#!/bin/bash
# the parity suite under starved kernel resources (2 ring stages, 2 raw buffers; then 3 consumer warps):
# exercises every wait path (stage_seq, raw_owner, grp_done) much harder than the default plan
mkdir -p gpurun_out
echo "== starved scratch"; PMU_STAGES=2 PMU_RAW_BUFFERS=2 timeout 900 python -m pytest tests -x -q -m gpu -k "not latency" > gpurun_out/pytest_stress1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_stress1.log
echo "== 3 consumer warps, 3 stages"; PMU_STAGES=3 PMU_CONSUMER_WARPS=3 PMU_RAW_BUFFERS=3 timeout 900 python -m pytest tests -x -q -m gpu -k "not latency" > gpurun_out/pytest_stress2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_stress2.log
echo "== per-element loader everywhere"; PMU_FORCE_ELEM_LOAD=1 timeout 900 python -m pytest tests -x -q -m gpu -k "not latency" > gpurun_out/pytest_stress3.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_stress3.log
echo "== row bulk copies instead of tensor-map slabs, batched estimator for small launches too, no programmatic dependent launch"; PMU_NO_TENSOR_MAP=1 PMU_WARP_PER_WINDOW_MAX=0 PMU_ZERO_COPY_MAX=0 PMU_NO_PDL=1 timeout 900 python -m pytest tests -x -q -m gpu -k "not latency and not warp_per_window" > gpurun_out/pytest_stress4.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_stress4.log
echo "== no window packing (one window per warp pass everywhere)"; PMU_PACK=1 timeout 900 python -m pytest tests -x -q -m gpu -k "not latency" > gpurun_out/pytest_stress5.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_stress5.log
