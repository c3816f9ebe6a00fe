#!/usr/bin/env python
"""tools/tune.py — kernel-time sweeps on one GPU (development aid, not part of the product).

Times pmub_estimate_async with CUDA events for several batch sizes and launch-plan knobs
(PMU_STAGES, PMU_CONSUMER_WARPS, PMU_POST_SLOTS, PMU_RAW_BUFFERS are read at creation), and fits
T(n) = a + b n to separate the fixed per-launch cost (ramp + tail) from the streaming rate."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from synchropmu_b200 import api, configs, signals  # noqa: E402


def time_case(cfg, n_inst, n_ch, lib, steps=200, frames=3, dtype=np.float64, harm=(), fpl=1):
    dev = torch.device("cuda")
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    prm = signals.three_phase_params(n_inst)
    pt = {k: torch.as_tensor(v, device=dev) for k, v in prm.items()}
    N = configs.win_len(cfg)
    xs, ms = [], []
    frames = max(frames, fpl)
    for t in range(frames):
        x = signals.steady(cfg, t, pt["amp"], pt["freq"], pt["phase"], xp=torch, ramp=pt["ramp"], harmonics=list(harm), like=pt["amp"])
        xs.append(x.to(tdt).reshape(n_inst, n_ch, N).contiguous())
        ms.append(torch.full((n_inst,), signals.mid_fracsec(cfg, t), dtype=tdt, device=dev))
    out = torch.empty((max(1, fpl), n_inst, n_ch, 4), dtype=tdt, device=dev)
    if fpl > 1:
        xs = [torch.stack(xs[:fpl]).contiguous()]
        ms = [torch.stack(ms[:fpl]).contiguous()]
        frames = 1
        steps = max(10, steps // fpl)
    est = api.BatchEstimator(cfg, n_inst, n_ch, dtype=dtype, lib=lib)
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        est.set_stream(s.cuda_stream)
        if os.environ.get("SWEEP_OVERLAP") == "1":       # the inputs are all resident before the first launch
            est.set_overlap(True)
        def go(i):
            if fpl > 1:
                est.estimate_frames_async(fpl, xs[0], ms[0], out)
            else:
                est.estimate_async(xs[i % frames], ms[i % frames], out)
        for i in range(5):
            go(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for i in range(steps):
            go(i)
        e1.record(s)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / steps * 1e3 / max(1, fpl)
    info = est.info
    est.close()
    del xs, ms, out
    torch.cuda.empty_cache()
    return us, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--cfg", default="config.ini")
    ap.add_argument("--tag", default="")
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    lib = api.load_library(Path(a.lib)) if a.lib else api.load_library()
    cfg = configs.NAMED[a.cfg]
    n_ch = 6
    bpe = configs.bytes_per_estimate(cfg, 8, n_ch)
    res = []

    def run(label, n_inst, cfg_=cfg, env=None, **kw):
        old = {}
        for k, v in (env or {}).items():
            old[k] = os.environ.get(k)
            os.environ[k] = str(v)
        try:
            us, info = time_case(cfg_, n_inst, n_ch, lib, **kw)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        n = n_inst * n_ch
        gbs = configs.bytes_per_estimate(cfg_, 8, n_ch) * n / us / 1e3
        r = dict(tag=a.tag, label=label, n=n, us=round(us, 2), est_per_s=round(n / us * 1e6), gbs=round(gbs, 1),
                 stages=info.n_stages, block=info.block, grid=info.grid)
        res.append(r)
        print(json.dumps(r), flush=True)

    for n_inst in (1024, 2048, 4096, 8192, 16384):
        run("size", n_inst)
    ns = np.array([r["n"] for r in res]); ts = np.array([r["us"] for r in res])
    b, a0 = np.polyfit(ns, ts, 1)
    print(json.dumps(dict(tag=a.tag, fit_fixed_us=round(a0, 2), fit_ns_per_window=round(b * 1e3, 3),
                          streaming_gbs=round(bpe / (b * 1e3), 1))), flush=True)
    if a.quick:
        return
    for fpl in (2, 4, 8, 16):
        run(f"frames_per_launch={fpl}", 4096, fpl=fpl)
    run("light_post(P=1,iter=0)", 4096, cfg_=dict(cfg, P=1, iter_eipdft=0))
    run("triggered(10% 75Hz)", 4096, harm=[(0.1, 75.0, 0.0)])
    for st in (3, 5, 7):
        run(f"stages={st}", 4096, env=dict(PMU_STAGES=st))
    for cw in (3, 5):
        run(f"consumers={cw}", 4096, env=dict(PMU_CONSUMER_WARPS=cw))
    run("post_slots=2", 4096, env=dict(PMU_POST_SLOTS=2))
    run("post_slots=7,raw=6", 4096, env=dict(PMU_POST_SLOTS=7, PMU_RAW_BUFFERS=6))


if __name__ == "__main__":
    main()
