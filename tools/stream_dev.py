#!/usr/bin/env python
"""tools/stream_dev.py — device-resident streaming front end: rings and new samples in HBM, F frames per push,
CUDA-event timed (development aid; the same leg as bench.py's `stream_device`).

    python tools/stream_dev.py [cfg] [F] [pushes] [i16]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from synchropmu_b200 import BatchEstimator, configs, signals  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "config.ini"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pushes = int(sys.argv[3]) if len(sys.argv) > 3 else 40
trig = len(sys.argv) > 4 and sys.argv[4] == "trig"
cfg = configs.NAMED[cfgname]
n_inst, n_ch = (21846, 6) if cfgname == "cfg5_600" else (4096, 6)
N, hop = configs.win_len(cfg), configs.hop(cfg)
dev = torch.device("cuda")
prm = signals.three_phase_params(n_inst)
pt = {k: torch.as_tensor(v, device=dev) for k, v in prm.items()}
harm = [(0.1, 75.0, 0.0)] if trig else []
n_res = 2 * F + 1
wins = []
for t in range(n_res):
    wins.append(signals.steady(cfg, t, pt["amp"], pt["freq"], pt["phase"], xp=torch, ramp=pt["ramp"], harmonics=harm,
                               like=pt["amp"]).reshape(n_inst, n_ch, N))
blocks = [torch.stack([wins[k * F + f + 1][:, :, N - hop:] for f in range(F)]).contiguous() for k in range(2)]
hist = wins[0][:, :, :N - hop].contiguous()
del wins
out = torch.empty((F, n_inst, n_ch, 4), dtype=torch.float64, device=dev)
est = BatchEstimator(cfg, n_inst, n_ch)
s = torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(s):
    est.set_stream(s.cuda_stream)
    est.stream_begin(hop=hop, max_frames_per_push=F, history=hist)
    for i in range(4):
        est.stream_push(blocks[i % 2], None, out=out, asynchronous=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(pushes):
        est.stream_push(blocks[i % 2], None, out=out, asynchronous=True)
    e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / pushes
n = n_inst * n_ch * F
print(json.dumps(dict(cfg=cfgname, F=F, trig=trig, ms_per_push=round(ms, 4), est_per_s=round(n / ms * 1e3),
                      finite=bool(torch.isfinite(out).all().item()))))
