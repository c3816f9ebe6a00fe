#!/usr/bin/env python
"""tools/single_call.py — one instance x one channel, reference test/test.c parameters (Q = 22, interharmonic present):
the latency-bound drop-in path.  Prints microseconds per synchronous call; run under ncu for a per-instruction profile."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from synchropmu_b200 import BatchEstimator, configs  # noqa: E402

cfg = configs.NAMED[sys.argv[1] if len(sys.argv) > 1 else "test_c_q22"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
N = configs.win_len(cfg)
t = np.arange(N) / cfg["fs"]
x = (2 * np.cos(2 * np.pi * 51 * t) + 0.2 * np.cos(2 * np.pi * 75 * t)).reshape(1, 1, N)
mid = np.zeros(1)
with BatchEstimator(cfg, 1, 1) as est:
    out = np.zeros((1, 1, 4))
    for _ in range(5):
        est.estimate(x, mid, out=out)
    t0 = time.perf_counter()
    for _ in range(n):
        est.estimate(x, mid, out=out)
    dt = time.perf_counter() - t0
print(f"{dt / n * 1e6:.1f} us per call, frame {out.ravel()}")
