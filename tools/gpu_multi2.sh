#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
echo "== ipc test"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "frames_land or sequence or c37118" 2>&1 | tail -3
echo "== ours N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1]);print(d['value'], d['roofline']['frac'], d.get('frame_gather_ms'), d.get('frames_p2p_to_rank0'), d['e2e']['value'], d['e2e_stream']['value'])"; tail -5 gpurun_out/bench_n$N.err
