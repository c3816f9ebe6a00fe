#!/usr/bin/env python
"""tools/sweep_smem.py — how to split shared memory between ring stages, raw-bin buffers and estimator slots
(development aid): times 8-frame launches for PMU_POST_SLOTS x PMU_RAW_BUFFERS on two configurations."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from synchropmu_b200 import api, configs  # noqa: E402
import tune  # noqa: E402

lib = api.load_library()
for cfgname in ("m_class_config.ini",):
    cfg = configs.NAMED[cfgname]
    for post in (1, 2):
        for raw in (2,):
            os.environ["PMU_POST_SLOTS"] = str(post)
            os.environ["PMU_RAW_BUFFERS"] = str(raw)
            n_inst = 4096 if cfgname != "cfg5_600" else 16384
            for fpl in (1, 8):
                us, info = tune.time_case(cfg, n_inst, 6, lib, fpl=fpl, steps=160)
                print(json.dumps(dict(pitch96=os.environ.get("PMU_PITCH96"), cfg=cfgname, post=post, raw=raw, fpl=fpl, stages=info.n_stages, us_per_frame=round(us, 2))), flush=True)
