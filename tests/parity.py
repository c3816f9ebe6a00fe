"""Parity gates of the pmu_estimate() path (BASELINE.json north_star).

double build : |d amp|/amp <= 1e-9, |d ph| <= 1e-9 rad (mod 2 pi), |d f| <= 1e-7 Hz,
               |d ROCOF| <= 1e-4 Hz/s, first-ipDFT peak bin bit-exact, NaN == NaN.
float build  : checked against the float_p=float reference at TVE <= 1e-5 (relative
               amplitude 1e-5, phase 1e-5 rad), FE <= 1e-4 Hz, RFE <= 1e-2 Hz/s.
"""
from __future__ import annotations

import math

import numpy as np

TOL = {
    "f64": dict(amp_rel=1e-9, ph=1e-9, freq=1e-7, rocof=1e-4),
    "f32": dict(amp_rel=1e-5, ph=1e-5, freq=1e-4, rocof=1e-2),
    # checker-vs-checker (same arithmetic, same libm): essentially bit-exact
    "tight": dict(amp_rel=1e-12, ph=1e-12, freq=1e-11, rocof=1e-9),
    # two code paths of the float build with the same arithmetic: outputs may differ in the last float bit
    "f32_ulp": dict(amp_rel=2.5e-7, ph=1e-6, freq=1e-5, rocof=1e-3),
}


def golden_floats(lst):
    """Inverse of make_golden.flt(): None -> NaN."""
    conv = {None: math.nan, "inf": math.inf, "-inf": -math.inf}
    return np.array([conv[v] if (v is None or isinstance(v, str)) else v for v in lst], dtype=np.float64)


def phase_diff(a, b):
    d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
    return d - 2 * math.pi * np.rint(d / (2 * math.pi))


def frame_errors(got, want):
    """got/want [..., 4] = amp, ph, freq, rocof -> dict of max errors; asserts the
    NaN pattern is identical (NaN == NaN)."""
    g = np.asarray(got, dtype=np.float64).reshape(-1, 4)
    w = np.asarray(want, dtype=np.float64).reshape(-1, 4)
    assert g.shape == w.shape, (g.shape, w.shape)
    nan_g, nan_w = np.isnan(g), np.isnan(w)
    assert np.array_equal(nan_g, nan_w), f"NaN pattern differs at rows {np.argwhere(nan_g != nan_w)[:8].tolist()}"
    ok = ~nan_w.any(axis=1)
    g, w = g[ok], w[ok]
    if len(g) == 0:
        return dict(amp_rel=0.0, ph=0.0, freq=0.0, rocof=0.0, n=0)
    den = np.maximum(np.abs(w[:, 0]), 1e-300)
    return dict(
        amp_rel=float(np.max(np.abs(g[:, 0] - w[:, 0]) / den)),
        ph=float(np.max(np.abs(phase_diff(g[:, 1], w[:, 1])))),
        freq=float(np.max(np.abs(g[:, 2] - w[:, 2]))),
        rocof=float(np.max(np.abs(g[:, 3] - w[:, 3]))),
        n=int(len(g)),
    )


def assert_frames_close(got, want, tol="f64", what=""):
    t = TOL[tol] if isinstance(tol, str) else tol
    e = frame_errors(got, want)
    bad = {k: (e[k], t[k]) for k in t if e[k] > t[k]}
    assert not bad, f"{what}: parity gate exceeded {bad} (all errors {e})"
    return e


def rocof_ambiguous(want, cfg, margin=2e-3):
    """The ROCOF stage (src/pmu_estimator.c:236-260) is a two-state machine with strict threshold tests on
    r = (f - f_old) frame_rate and on its difference.  In the float build a one-ulp difference in the frequency estimate
    (3.8e-6 Hz at 60 Hz -> 2.3e-4 Hz/s in r) legitimately flips a decision that lies that close to a threshold, and the
    stream's ROCOF then differs until the states meet again.  want [T, S, C, 4] (oracle frames) -> bool [T, S, C]: True
    from the first frame on at which one of the oracle's own decisions is closer than `margin` to a threshold."""
    w = np.asarray(want, dtype=np.float64)
    fr = float(cfg["frame_rate"])
    thr = [float(v) for v in cfg["rocof_thresh"]]
    f_prev = np.full(w.shape[1:3], float(cfg["f0"]))
    r_prev = np.zeros(w.shape[1:3])
    amb = np.zeros(w.shape[:3], dtype=bool)
    seen = np.zeros(w.shape[1:3], dtype=bool)
    for t in range(w.shape[0]):
        r = (w[t, ..., 2] - f_prev) * fr
        rd = (r - r_prev) * fr
        near = (np.abs(np.abs(r) - thr[0]) < margin) | (np.abs(np.abs(rd) - thr[1]) < margin * fr) | \
               (np.abs(np.abs(r) - thr[2]) < margin)
        seen |= near
        amb[t] = seen
        f_prev, r_prev = w[t, ..., 2], r
    return amb
