#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref).

Run in the dev container (where /root/reference exists and ``make -C oracle ref``
has built oracle/_ref/libref_*.so):

    python tests/golden/make_golden.py

Outputs (committed; they travel to the GPU box, the reference does not):
  single_calls.json + single_calls.npz : one pmu_estimate call per case on a fresh
      context — inputs stored bit-exactly, outputs = frames, first-ipDFT peak bin and
      the windowed spectrum the reference's pmue_fft_r produced.
  streams.json : consecutive-frame runs (ROCOF state) over varied C37.118.1-style
      streams; inputs are regenerated from the stored parameters by
      synchropmu_b200.signals, outputs stored.
  ini_configs.json : what the reference's pmu_init(CONFIG_FROM_INI) makes of the
      .ini files under config/ and tests/golden/ini/.
The signals restate the reference drivers' generators (test/test.c:21-28,60;
examples/simple_example.c:10-24; test/test2.c:89-99).
"""
from __future__ import annotations

import json
import math
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle_lib import CpuChecker  # noqa: E402
from synchropmu_b200 import configs, signals  # noqa: E402
from stream_cases import stream_case_windows, STREAM_CASES  # noqa: E402


def flt(a):
    """JSON-safe float list with full precision (repr round-trips doubles)."""
    out = []
    for v in np.asarray(a, dtype=np.float64).reshape(-1):
        out.append(None if math.isnan(v) else ("inf" if v == math.inf else ("-inf" if v == -math.inf else float(v))))
    return out


def single_cases():
    c1 = configs.DEFAULT_INI
    cases = []

    def tone(cfg, amp, f, ph, extra=(), n_ch=1, phases=None):
        phs = [ph] * n_ch if phases is None else phases
        return signals.steady(cfg, 0, [amp] * n_ch, [f] * n_ch, phs, harmonics=list(extra))

    ih = [(0.1, 75.0, 0.0)]
    cases.append(dict(name="config_ini_51hz_plus_75hz", cfg=c1, C=1, mid=0.0,
                      win=tone(c1, 2.0, 51.0, 0.0, ih)))
    cases.append(dict(name="config_ini_clean_50hz_ph1", cfg=c1, C=1, mid=0.0,
                      win=tone(c1, 2.0, 50.0, 1.0)))
    cases.append(dict(name="config_ini_clean_50hz_ph1_mid001", cfg=c1, C=1, mid=0.01,
                      win=tone(c1, 2.0, 50.0, 1.0)))
    cases.append(dict(name="p_class_51hz_plus_75hz", cfg=configs.P_CLASS_INI, C=1, mid=0.0,
                      win=tone(configs.P_CLASS_INI, 2.0, 51.0, 0.0, ih)))
    q13 = dict(c1, Q=13)
    cases.append(dict(name="test2_ncyc4_q13", cfg=q13, C=1, mid=0.0, win=tone(q13, 2.0, 51.0, 0.0, ih)))
    n2 = dict(c1, n_cycles=2)
    cases.append(dict(name="test2_ncyc2_q3", cfg=n2, C=1, mid=0.0, win=tone(n2, 2.0, 51.0, 0.0, ih)))
    n8 = dict(c1, n_cycles=8, fs=51200, n_bins=18, Q=13)
    cases.append(dict(name="test2_fs51200_ncyc8_k18_q13", cfg=n8, C=1, mid=0.0,
                      win=tone(n8, 2.0, 51.0, 0.0, ih)))
    cases.append(dict(name="test_c_q22", cfg=configs.TEST_C_Q22, C=1, mid=0.0,
                      win=tone(configs.TEST_C_Q22, 2.0, 51.0, 0.0, ih)))
    m = configs.M_CLASS_INI
    ph6 = [0.0, -2 * math.pi / 3, 2 * math.pi / 3, 0.0, -2 * math.pi / 3, 2 * math.pi / 3]
    cases.append(dict(name="m_class_6ch_51hz_plus_75hz", cfg=m, C=6, mid=0.0,
                      win=tone(m, 2.0, 51.0, 0.0, ih, n_ch=6, phases=ph6)))
    cases.append(dict(name="m_class_6ch_clean_50p2", cfg=m, C=6, mid=0.37,
                      win=tone(m, 325.27, 50.2, 0.0, (), n_ch=6, phases=ph6)))
    c5 = configs.CFG5_600
    cases.append(dict(name="cfg5_600_49p3hz_plus_h3", cfg=c5, C=1, mid=0.2,
                      win=tone(c5, 1.0, 49.3, -2.0, [(0.05, 147.9, 0.4)])))
    c60 = dict(c1, f0=60, fs=15360, frame_rate=60, n_bins=8)
    cases.append(dict(name="sys60hz_59p7", cfg=c60, C=1, mid=0.5, win=tone(c60, 120.0, 59.7, 2.5)))
    cdc = tone(c1, 1.0, 50.4, 0.7) + 0.25
    cases.append(dict(name="config_ini_dc_offset", cfg=c1, C=1, mid=0.0, win=cdc))
    lowf = tone(c1, 1.0, 5.0, 0.2)
    cases.append(dict(name="config_ini_peak_in_bin0", cfg=c1, C=1, mid=0.0, win=lowf))
    hif = tone(c1, 1.0, 128.0, 0.2)
    cases.append(dict(name="config_ini_peak_in_last_bin", cfg=c1, C=1, mid=0.0, win=hif))
    zeros = np.zeros((1, configs.win_len(c1)))
    cases.append(dict(name="config_ini_all_zero_window", cfg=c1, C=1, mid=0.0, win=zeros))
    return cases


def gen_single():
    meta, arrays = [], {}
    for case in single_cases():
        for dt in (np.float64, np.float32):
            tag = "f64" if dt == np.float64 else "f32"
            ck = CpuChecker("ref", dt, case["C"])
            h = ck.open(case["cfg"])
            assert h, case["name"]
            win = np.ascontiguousarray(case["win"], dtype=dt)
            frames, rc = ck.estimate(h, win, case["mid"])
            assert rc == 0
            key = f"{case['name']}__{tag}"
            arrays[key + "__win"] = win
            spec = ck.last_spectrum(h)
            meta.append(dict(name=case["name"], dtype=tag, C=case["C"], cfg=case["cfg"], mid=case["mid"],
                             key=key, frames=flt(frames), km_last=ck.last_km(h),
                             norm_factor=ck.norm_factor(h), win_len=ck.win_len(h),
                             spectrum_re=flt(spec.real), spectrum_im=flt(spec.imag)))
            ck.close(h)
    np.savez_compressed(HERE / "single_calls.npz", **arrays)
    (HERE / "single_calls.json").write_text(json.dumps(meta, indent=1))
    print("single_calls:", len(meta), "cases")


def gen_streams():
    out = []
    for sc in STREAM_CASES:
        cfg = configs.NAMED[sc["cfg"]]
        for dt in (np.float64, np.float32):
            tag = "f64" if dt == np.float64 else "f32"
            win, mid = stream_case_windows(sc, dtype=dt)           # [T, S, 1, N], [T, S]
            ck = CpuChecker("ref", dt, 1)
            frames, km, _ = ck.run(cfg, win, mid, n_threads=1)
            out.append(dict(case=sc["name"], cfg=sc["cfg"], dtype=tag,
                            frames=flt(frames), km=[int(v) for v in km.reshape(-1)],
                            shape=list(frames.shape)))
    (HERE / "streams.json").write_text(json.dumps(out))
    print("streams:", len(out), "runs")


def gen_ini():
    res = {}
    files = sorted((ROOT / "config").glob("*.ini")) + sorted((HERE / "ini").glob("*.ini"))
    ck = CpuChecker("ref", np.float64, 1)
    for f in files:
        rel = str(f.relative_to(ROOT))
        h = ck.open_ini(str(f))
        if not h:
            res[rel] = dict(accepted=False)
            continue
        n = ck.win_len(h)
        fs = 15360.0 if "60hz" in f.name else 25600.0
        t = np.arange(n) / fs
        x = (2.0 * np.cos(2 * math.pi * 51.0 * t) + 0.2 * np.cos(2 * math.pi * 75.0 * t)).reshape(1, -1)
        frames, rc = ck.estimate(h, x, 0.0)
        res[rel] = dict(accepted=True, win_len=n, n_bins=ck.n_bins(h), norm_factor=ck.norm_factor(h),
                        probe_fs=fs, frame=flt(frames))
        ck.close(h)
    (HERE / "ini_configs.json").write_text(json.dumps(res, indent=1))
    print("ini_configs:", {k: v["accepted"] for k, v in res.items()})


if __name__ == "__main__":
    gen_single()
    gen_streams()
    gen_ini()
