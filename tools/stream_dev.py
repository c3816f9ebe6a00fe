#!/usr/bin/env python
"""tools/stream_dev.py — device-resident streaming front end: rings and new samples in HBM, F frames per push,
CUDA-event timed (development aid; the same leg as bench.py's `stream_device`).

    python tools/stream_dev.py [cfg] [F] [pushes] [trig|-] [inplace|ingest|both]

The pushes come from ONE record that is continuous in time (cycling a couple of blocks would put a phase jump into every
window that straddles a seam, and those windows would run the interference loop).  `ingest`: the record is handed over as
device blocks and scattered into the rings by the ingest kernel.  `inplace` (default `both`: ingest first, then in place):
the rings were sized for the whole record, every sample is in them already (PMUB_PUSH_IN_PLACE) - the estimator launch
alone."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from synchropmu_b200 import BatchEstimator, configs, signals  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "config.ini"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pushes = int(sys.argv[3]) if len(sys.argv) > 3 else 16
trig = len(sys.argv) > 4 and sys.argv[4] == "trig"
mode = sys.argv[5] if len(sys.argv) > 5 else "both"
cfg = configs.NAMED[cfgname]
n_inst, n_ch = (21846, 6) if cfgname == "cfg5_600" else (4096, 6)
N, hop = configs.win_len(cfg), configs.hop(cfg)
dev = torch.device("cuda")
prm = signals.three_phase_params(n_inst)
pt = {k: torch.as_tensor(v, device=dev) for k, v in prm.items()}
harm = [(0.1, 75.0, 0.0)] if trig else []
warm = 3
K = warm + pushes


def span(first, n):
    return signals.steady(cfg, 0, pt["amp"], pt["freq"], pt["phase"], xp=torch, ramp=pt["ramp"], harmonics=harm,
                          like=pt["amp"], span=(first, n)).reshape(n_inst, n_ch, n)


hist = span(0, N - hop).contiguous()
rec = torch.empty((K * F, n_inst, n_ch, hop), dtype=torch.float64, device=dev)
for j in range(K * F):
    rec[j] = span(N - hop + j * hop, hop)
out = torch.empty((F, n_inst, n_ch, 4), dtype=torch.float64, device=dev)
est = BatchEstimator(cfg, n_inst, n_ch)
s = torch.cuda.Stream()
torch.cuda.synchronize()


def timed(fn):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    est.reset()      # ROCOF memory only: a window across the replay's seam may have left a NaN in a stream's delay line
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(warm, K):
        fn(i)
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / pushes


with torch.cuda.stream(s):
    est.set_stream(s.cuda_stream)
    est.stream_begin(hop=hop, max_frames_per_push=K * F, history=hist)
    n = n_inst * n_ch * F
    res = dict(cfg=cfgname, F=F, trig=trig, pushes=pushes)
    ms = timed(lambda i: est.stream_push(rec[i * F:(i + 1) * F], None, out=out, asynchronous=True))
    res.update(ingest_ms_per_push=round(ms, 4), ingest_est_per_s=round(n / ms * 1e3))
    if mode in ("inplace", "both"):
        # the write position is back at 0: the record is replayed (the N/hop - 1 windows across its seam fall into the warm-up)
        ms = timed(lambda i: est.stream_advance(F, out))
        res.update(ms_per_push=round(ms, 4), est_per_s=round(n / ms * 1e3))
    res["finite"] = bool(torch.isfinite(out).all().item())
    import os
    if os.environ.get("SUSTAIN_PUSHES"):
        # replay the rings over and over (N/hop - 1 of every K*F frames straddle the record's seam): clocks under the power cap
        n_s = int(os.environ["SUSTAIN_PUSHES"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for i in range(n_s):
            est.stream_advance(F, out)
        e1.record(s)
        torch.cuda.synchronize()
        res.update(sustained_ms_per_push=round(e0.elapsed_time(e1) / n_s, 4), sustained_est_per_s=round(n * n_s / e0.elapsed_time(e1) * 1e3))
    est.set_debug(True)
    est.stream_advance(1, out[:1])
    torch.cuda.synchronize()
    res["triggered_fraction"] = float(est.get_debug()[1].mean())
print(json.dumps(res))
