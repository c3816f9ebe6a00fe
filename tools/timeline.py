#!/usr/bin/env python
"""tools/timeline.py — where does a single-frame launch spend its fixed time?  (profiling aid)
Uses pmub_debug_timeline: per CTA / warp globaltimer stamps of the last launch."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from synchropmu_b200 import api, configs, signals  # noqa: E402

cfg = configs.DEFAULT_INI
n_inst, n_ch = 4096, 6
dev = torch.device("cuda")
prm = signals.three_phase_params(n_inst)
pt = {k: torch.as_tensor(v, device=dev) for k, v in prm.items()}
x = signals.steady(cfg, 0, pt["amp"], pt["freq"], pt["phase"], xp=torch, ramp=pt["ramp"], like=pt["amp"]).reshape(n_inst, n_ch, -1).contiguous()
m = torch.zeros(n_inst, dtype=torch.float64, device=dev)
out = torch.empty((n_inst, n_ch, 4), dtype=torch.float64, device=dev)
est = api.BatchEstimator(cfg, n_inst, n_ch)
lib = est.lib
lib.pmub_debug_timeline.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
nc, nw = C.c_int(), C.c_int()
lib.pmub_debug_timeline(est._h, 1, None, C.byref(nc), C.byref(nw))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    est.set_stream(s.cuda_stream)
    for _ in range(5):
        est.estimate_async(x, m, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    est.estimate_async(x, m, out)
    e1.record(s)
torch.cuda.synchronize()
tl = np.zeros((nc.value, nw.value, 4), dtype=np.uint64)
lib.pmub_debug_timeline(est._h, 1, tl.ctypes.data, C.byref(nc), C.byref(nw))
t = tl[:, 1:, :].astype(np.int64)          # consumer warps
t0 = t[..., 0].min()
rel = (t - t0) / 1e3
print(f"event-timed launch: {e0.elapsed_time(e1) * 1e3:.1f} us")
for i, name in enumerate(["setup done", "first window ready", "tickets exhausted", "warp exit"]):
    v = rel[..., i]
    print(f"{name:20s} min {v.min():7.1f}  median {np.median(v):7.1f}  max {v.max():7.1f} us (since the first warp's setup)")
cta_end = rel[..., 3].max(axis=1)
cta_dft_end = rel[..., 2].max(axis=1)
print(f"per-CTA: DFT stream ends  median {np.median(cta_dft_end):6.1f} max {cta_dft_end.max():6.1f};  CTA ends median {np.median(cta_end):6.1f} max {cta_end.max():6.1f}")
print(f"tail after the last ticket (per CTA): median {np.median(cta_end - cta_dft_end):5.1f} max {(cta_end - cta_dft_end).max():5.1f} us")
