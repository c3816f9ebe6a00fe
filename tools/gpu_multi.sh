#!/bin/bash
# multi-GPU bench exactly as the driver launches it, plus the reference arm
N=${1:-2}
mkdir -p gpurun_out
echo "== reference arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.log; tail -3 gpurun_out/bench_ref.err
echo "== ours N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cat gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2> gpurun_out/bench_ref_n$N.err; echo "rc=$?"; cat gpurun_out/bench_ref_n$N.log
