#!/bin/bash
# what the driver does at round end + the evidence captures, in one gpurun call
mkdir -p gpurun_out
echo "== traffic (ncu dram bytes tied to the kernel source hash)"; timeout 900 python tools/update_traffic.py > gpurun_out/final_traffic.log 2>&1; echo "rc=$?"; cp profiles/traffic_per_launch.json gpurun_out/traffic_per_launch.json
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/final_pytest_gpu.log; grep -a "latency\]\|f32 sweep" gpurun_out/final_pytest_gpu.log | sort -u > gpurun_out/final_latency.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/final_bench_ref.json 2>gpurun_out/final_bench_ref.err; echo "rc=$?"; cut -c1-200 gpurun_out/final_bench_ref.json
echo "== ours (driver-like: --steps 20 --warmup 5)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?"; tail -3 gpurun_out/final_bench.err
echo "== ours (defaults)"; timeout 900 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "rc=$?"
echo "== sweeps"; timeout 900 python tools/r2_sweep.py --tag final > gpurun_out/final_sweeps.jsonl 2>/dev/null; echo "rc=$?"
SHORT="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-workloads --no-sustained"
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final_launches_ncu.csv $SHORT > /dev/null 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmu_fused -s 4 -c 1 -o /tmp/prof_final -f $SHORT > /dev/null 2>&1; echo "rc=$?"; ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/final_ncu_full_raw.csv 2>/dev/null
