"""C37.118.1-style stream cases shared by the golden generator and the parity tests.

BASELINE.json config 4 ("synthetic sweep batch: off-nominal freq, harmonics,
interharmonics, AM/PM modulation, freq ramp") at a size the CPU checkers finish in
well under a second.  A case = one estimator configuration + a list of
single-channel streams; every stream is run for ``frames`` consecutive frames so
the ROCOF state machine (reference src/pmu_estimator.c:236-260) is exercised.
"""
from __future__ import annotations

import math

import numpy as np

from synchropmu_b200 import configs, signals


def _streams():
    s = []
    # off-nominal steady state, 45..55 Hz
    for f in (45.0, 47.5, 49.5, 50.5, 52.5, 55.0):
        s.append(dict(kind="steady", amp=1.0 + 0.01 * f, freq=f, phase=0.1 * f, harm=[]))
    # single harmonic at 1 % and 10 %
    for h, lvl in ((2, 0.01), (3, 0.10), (7, 0.10), (25, 0.01)):
        s.append(dict(kind="steady", amp=100.0, freq=50.1, phase=-1.0, harm=[(lvl, 50.1 * h, 0.3)]))
    # out-of-band interharmonics, 10 %
    for f, fi in ((47.5, 20.0), (50.0, 75.0), (52.5, 90.0), (51.0, 75.0)):
        s.append(dict(kind="steady", amp=2.0, freq=f, phase=0.5, harm=[(0.10, fi, 0.0)]))
    # amplitude + phase modulation
    for fm in (0.5, 2.0, 5.0):
        s.append(dict(kind="mod", amp=10.0, freq=50.0, fm=fm, phase=0.25))
    # frequency ramps, +-1 Hz/s
    s.append(dict(kind="steady", amp=1.0, freq=49.0, phase=0.0, harm=[], ramp=1.0))
    s.append(dict(kind="steady", amp=1.0, freq=51.0, phase=2.0, harm=[], ramp=-1.0))
    return s


STREAM_CASES = [
    dict(name="sweep_config_ini", cfg="config.ini", frames=8, streams=_streams()),
    dict(name="sweep_m_class", cfg="m_class_config.ini", frames=6, streams=_streams()),
    dict(name="sweep_cfg5_600", cfg="cfg5_600", frames=8, streams=_streams()),
    dict(name="sweep_p_class", cfg="p_class_config.ini", frames=6, streams=_streams()),
]


def stream_case_windows(sc: dict, dtype=np.float64):
    """-> windows [T, S, 1, N] (dtype), mid_fracsec [T, S] (float64)."""
    cfg = configs.NAMED[sc["cfg"]]
    T, S, N = sc["frames"], len(sc["streams"]), configs.win_len(cfg)
    win = np.zeros((T, S, 1, N), dtype=np.float64)
    mid = np.zeros((T, S), dtype=np.float64)
    for t in range(T):
        mid[t, :] = signals.mid_fracsec(cfg, t)
        for i, st in enumerate(sc["streams"]):
            if st["kind"] == "steady":
                ramp = [st["ramp"]] if "ramp" in st else None
                x = signals.steady(cfg, t, [st["amp"]], [st["freq"]], [st["phase"]],
                                   harmonics=st["harm"], ramp=ramp)
            else:
                x = signals.modulated(cfg, t, [st["amp"]], [st["freq"]], [st["fm"]],
                                      phase=[st["phase"]])
            win[t, i, 0, :] = x[0]
    return win.astype(dtype), mid
