#!/bin/bash
# round 2, first GPU pass: parity of the all-warps-post kernel, then A/B against the round-1 build
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2a_pytest.log
echo "== new"; timeout 600 python tools/r2_sweep.py --tag new > gpurun_out/r2a_new.jsonl 2> gpurun_out/r2a_new.err; echo "rc=$?"; cat gpurun_out/r2a_new.jsonl
echo "== base"; timeout 600 python tools/r2_sweep.py --tag base --lib synchropmu_b200/lib_base/libpmu_estimator.so > gpurun_out/r2a_base.jsonl 2> gpurun_out/r2a_base.err; echo "rc=$?"; cat gpurun_out/r2a_base.jsonl
echo "== knobs"; timeout 600 python tools/r2_sweep.py --tag knobs --cases c3,c3_trig,mc_trig,c5_trig --env "PMU_RAW_BUFFERS=4;PMU_RAW_BUFFERS=8;PMU_TRADE_STAGE=0;PMU_RAW_BUFFERS=3" > gpurun_out/r2a_knobs.jsonl 2> gpurun_out/r2a_knobs.err; echo "rc=$?"; cat gpurun_out/r2a_knobs.jsonl
