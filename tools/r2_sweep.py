#!/usr/bin/env python
"""tools/r2_sweep.py — device-timed throughput of the named workloads, untriggered and with every stream triggering
the interference iterations, optionally under launch-plan knobs (development aid; prints one JSON line per case).

    python tools/r2_sweep.py [--lib path/to/libpmu_estimator.so] [--tag T] [--cases a,b,...] [--env K=V,K=V]
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from synchropmu_b200 import api, configs  # noqa: E402
from tune import time_case  # noqa: E402

PEAK = 6462.1
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass

TRIG = [(0.1, 75.0, 0.0)]
CASES = {
    # name: (cfg, dtype, n_inst, frames per launch, harmonics)
    "c3": ("config.ini", np.float64, 4096, 8, ()),
    "c3_1f": ("config.ini", np.float64, 4096, 1, ()),
    "c3_trig": ("config.ini", np.float64, 4096, 8, TRIG),
    "c3_f32": ("config.ini", np.float32, 4096, 8, ()),
    "c3_f32_trig": ("config.ini", np.float32, 4096, 8, TRIG),
    "c5": ("cfg5_600", np.float64, 21846, 8, ()),
    "c5_trig": ("cfg5_600", np.float64, 21846, 8, TRIG),
    "mc": ("m_class_config.ini", np.float64, 4096, 8, ()),
    "mc_trig": ("m_class_config.ini", np.float64, 4096, 8, TRIG),
    "pc": ("p_class_config.ini", np.float64, 8192, 8, ()),
    "q22_trig": ("test_c_q22", np.float64, 4096, 4, TRIG),
    "t2": ("test2_16cyc", np.float64, 1024, 4, ()),
    "mc_1f": ("m_class_config.ini", np.float64, 4096, 1, ()),
    "c3_trig_1f": ("config.ini", np.float64, 4096, 1, TRIG),
    "c3_small_1f": ("config.ini", np.float64, 1024, 1, ()),
    "pc_f32": ("p_class_config.ini", np.float32, 8192, 8, ()),
    "c5_f32": ("cfg5_600", np.float32, 21846, 8, ()),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--tag", default="")
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--env", default="")
    ap.add_argument("--steps", type=int, default=160)
    a = ap.parse_args()
    lib = api.load_library(Path(a.lib)) if a.lib else api.load_library()
    # --env "A=1,B=2;A=3" : one pass over the cases per ';'-separated set (the plan knobs are read at creation)
    for envset in a.env.split(";"):
        run_cases(a, lib, envset)


def run_cases(a, lib, envset):
    old = {}
    for kv in filter(None, envset.split(",")):
        k, v = kv.split("=")
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        _run_cases(a, lib, envset)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run_cases(a, lib, envset):
    for name in a.cases.split(","):
        cfgname, dt, n_inst, fpl, harm = CASES[name]
        cfg = configs.NAMED[cfgname]
        us, info = time_case(cfg, n_inst, 6, lib, steps=a.steps, dtype=dt, harm=harm, fpl=fpl)
        n = n_inst * 6                      # estimates per frame; time_case returns microseconds per frame
        us_launch = us * max(1, fpl)
        bpe = configs.bytes_per_estimate(cfg, np.dtype(dt).itemsize, 6)
        rate = n / us * 1e6
        print(json.dumps(dict(tag=a.tag, env=envset, case=name, est_per_s=round(rate), frac=round(rate * bpe / 1e9 / PEAK, 4),
                              us_per_launch=round(us_launch, 1), stages=info.n_stages, grid=info.grid)), flush=True)


if __name__ == "__main__":
    main()
