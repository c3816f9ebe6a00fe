#!/usr/bin/env python
"""tools/wide_bench.py — throughput of the general (n_bins > 23) kernel next to the fused one (development aid)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from synchropmu_b200 import api, configs  # noqa: E402
import tune  # noqa: E402

lib = api.load_library()
for over in (dict(n_bins=23), dict(n_bins=24), dict(n_bins=40), dict(n_bins=100), dict(n_cycles=30, n_bins=40)):
    cfg = dict(configs.DEFAULT_INI, **over)
    us, info = tune.time_case(cfg, 1024, 6, lib, steps=20)
    n = 1024 * 6
    print(json.dumps(dict(over=over, win_len=configs.win_len(cfg), kernel="general" if info.dft_radix == 0 else "fused",
                          us_per_frame=round(us, 1), est_per_s=round(n / us * 1e6))), flush=True)
