"""Named estimator configurations of the pmu_estimate() path.

Each dict carries exactly the fields of ``estimator_config`` (reference
src/pmu_estimator.h:30-43).  The first three restate the shipped .ini files
(config/config.ini:13-32, config/p_class_config.ini, config/m_class_config.ini);
``CFG5_600`` is BASELINE.json config 5 (3 cycles @ 10 kHz, 600 samples) and
``TEST_C_Q22`` is the struct configuration of test/test.c:25-37.
"""
from __future__ import annotations

_ROCOF = dict(rocof_thresh=[3.0, 25.0, 0.035], rocof_low_pass_coeffs=[0.5913, 0.2043, 0.2043])

DEFAULT_INI = dict(n_cycles=4, f0=50, frame_rate=50, fs=25600, n_bins=11, P=3, Q=3,
                   iter_eipdft=1, interf_trig=0.0033, **_ROCOF)
P_CLASS_INI = dict(n_cycles=2, f0=50, frame_rate=50, fs=25600, n_bins=5, P=3, Q=1,
                   iter_eipdft=0, interf_trig=0.0033, **_ROCOF)
M_CLASS_INI = dict(n_cycles=6, f0=50, frame_rate=50, fs=25600, n_bins=11, P=3, Q=10,
                   iter_eipdft=1, interf_trig=0.0033, **_ROCOF)
CFG5_600 = dict(n_cycles=3, f0=50, frame_rate=50, fs=10000, n_bins=8, P=3, Q=3,
                iter_eipdft=1, interf_trig=0.0033, **_ROCOF)
TEST_C_Q22 = dict(n_cycles=4, f0=50, frame_rate=50, fs=25600, n_bins=11, P=3, Q=22,
                  iter_eipdft=1, interf_trig=0.0033, **_ROCOF)
# one point of the reference's wall-time sweep (test/test2.c:62-76: fs = 51.2 kHz, n_cycles = 16 -> 16384 samples,
# n_bins = n_cycles + 2 = 18, first Q of the sweep)
TEST2_16CYC = dict(n_cycles=16, f0=50, frame_rate=50, fs=51200, n_bins=18, P=3, Q=3,
                   iter_eipdft=1, interf_trig=0.0033, **_ROCOF)

NAMED = {
    "config.ini": DEFAULT_INI,
    "p_class_config.ini": P_CLASS_INI,
    "m_class_config.ini": M_CLASS_INI,
    "cfg5_600": CFG5_600,
    "test_c_q22": TEST_C_Q22,
    "test2_16cyc": TEST2_16CYC,
}


def win_len(cfg: dict) -> int:
    """``n_cycles * fs / f0`` in unsigned arithmetic (src/pmu_estimator.c:365)."""
    return int(cfg["n_cycles"]) * int(cfg["fs"]) // int(cfg["f0"])


def hop(cfg: dict) -> int:
    """Samples between consecutive frames, fs / frame_rate."""
    return int(cfg["fs"]) // int(cfg["frame_rate"])


def bytes_per_estimate(cfg: dict, itemsize: int = 8, n_channels: int = 1) -> float:
    """Algorithmic HBM bytes per phasor estimate:
    window samples read + frame written + ROCOF state read and written
    (+ the per-instance mid_fracsec shared by its channels)."""
    s = itemsize
    return win_len(cfg) * s + 4 * s + 2 * (3 * s + 1) + s / n_channels
