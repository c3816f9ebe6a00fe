"""Synthetic IEEE C37.118.1-style window generators (host numpy or device torch).

The reference ships no data; its drivers synthesise one window in place
(test/test.c:49-62 — ``A cos(2 pi f t + ph) + A ki cos(2 pi fi t + ph)``;
examples/simple_example.c:20-24).  These helpers generate the same kinds of
signal for whole batches of streams and consecutive frames: frame ``m`` of a
stream covers samples ``m*hop .. m*hop + N - 1`` of that stream's waveform, and
its ``mid_fracsec`` is the time of the window centre modulo one second
(SURVEY.md section 8(d), config 3).

Every function takes ``xp`` (``numpy`` or ``torch``) so the bench can build its
batch directly in HBM with the same formulas the CPU tests use.
"""
from __future__ import annotations

import math

import numpy as np

from .configs import hop as _hop
from .configs import win_len as _win_len


def _arange(xp, n, like=None):
    if xp is np:
        return np.arange(n, dtype=np.float64)
    return xp.arange(n, dtype=xp.float64, device=like.device if like is not None else None)


def _col(xp, v, like=None):
    """Per-stream parameter -> column vector [S, 1] of float64."""
    if xp is np:
        return np.asarray(v, dtype=np.float64).reshape(-1, 1)
    if not hasattr(v, "device"):
        v = xp.as_tensor(np.asarray(v, dtype=np.float64),
                         device=like.device if like is not None else None)
    return v.to(xp.float64).reshape(-1, 1)


def frame_times(cfg: dict, frame: int, xp=np, like=None):
    """Sample times [N] (seconds) of frame ``frame``."""
    n = _win_len(cfg)
    return (_arange(xp, n, like) + float(frame * _hop(cfg))) / float(cfg["fs"])


def mid_fracsec(cfg: dict, frame: int) -> float:
    """Fraction of second at the centre of frame ``frame`` (window centre = N/2)."""
    t = (frame * _hop(cfg) + _win_len(cfg) // 2) / float(cfg["fs"])
    return t - math.floor(t)


def span_times(cfg: dict, first_sample: int, n_samples: int, xp=np, like=None):
    """Sample times [n_samples] (seconds) of the samples ``first_sample ...`` of the record frame 0 starts."""
    return (_arange(xp, n_samples, like) + float(first_sample)) / float(cfg["fs"])


def steady(cfg, frame, amp, freq, phase, xp=np, harmonics=(), ramp=None, like=None, span=None):
    """Windows [S, N] of ``amp cos(2 pi f t + pi R t^2 + ph)`` plus extra tones.

    amp/freq/phase/ramp are per-stream arrays of length S.  ``harmonics`` is a
    sequence of ``(rel_amp, freq_hz_or_array, phase_offset)`` added as
    ``amp*rel_amp*cos(2 pi f_h t + ph + phase_offset)``: (0.1, 75.0, 0.0)
    reproduces the interharmonic of test/test.c:60.

    ``span=(first_sample, n_samples)`` replaces the window of ``frame`` by any
    stretch of the same record (frame t is its samples ``t*hop ... t*hop+N-1``):
    the streaming front end is fed from one continuous record, hop by hop.
    """
    a = _col(xp, amp, like)
    if span is None:
        t = frame_times(cfg, frame, xp, a if xp is not np else None).reshape(1, -1)
    else:
        t = span_times(cfg, int(span[0]), int(span[1]), xp, a if xp is not np else None).reshape(1, -1)
    f = _col(xp, freq, a if xp is not np else None)
    p = _col(xp, phase, a if xp is not np else None)
    arg = 2.0 * math.pi * f * t + p
    if ramp is not None:
        r = _col(xp, ramp, a if xp is not np else None)
        arg = arg + math.pi * r * t * t
    x = a * xp.cos(arg)
    for rel, fh, dph in harmonics:
        fh_c = _col(xp, fh, a if xp is not np else None) if not np.isscalar(fh) else fh
        x = x + (a * rel) * xp.cos(2.0 * math.pi * fh_c * t + p + dph)
    return x


def modulated(cfg, frame, amp, f0, fm, ka=0.1, kp=0.1, phase=0.0, xp=np, like=None):
    """C37.118.1 AM+PM test signal,
    ``A (1 + ka cos(2 pi fm t)) cos(2 pi f0 t + kp cos(2 pi fm t - pi) + ph)``."""
    a = _col(xp, amp, like)
    like2 = a if xp is not np else None
    t = frame_times(cfg, frame, xp, like2).reshape(1, -1)
    f = _col(xp, f0, like2)
    m = _col(xp, fm, like2)
    p = _col(xp, phase, like2)
    return a * (1.0 + ka * xp.cos(2.0 * math.pi * m * t)) * xp.cos(
        2.0 * math.pi * f * t + kp * xp.cos(2.0 * math.pi * m * t - math.pi) + p)


def three_phase_params(n_inst: int, seed: int = 20240601):
    """Per-stream parameters for ``n_inst`` 6-channel instances (3 V + 3 I):
    f_i = 50 + U(-0.5, 0.5), amplitude scale U(0.9, 1.1), phi_i = U(-pi, pi),
    ramp R_i = U(-1, 1) Hz/s — SURVEY.md section 8(d) config 3 (numpy PCG64)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = 50.0 + rng.uniform(-0.5, 0.5, n_inst)
    scale = rng.uniform(0.9, 1.1, n_inst)
    phi = rng.uniform(-math.pi, math.pi, n_inst)
    ramp = rng.uniform(-1.0, 1.0, n_inst)
    base_amp = np.array([325.27, 325.27, 325.27, 14.14, 14.14, 14.14])
    base_ph = np.array([0.0, -2 * math.pi / 3, 2 * math.pi / 3,
                        -0.3, -2 * math.pi / 3 - 0.3, 2 * math.pi / 3 - 0.3])
    amp = (scale[:, None] * base_amp[None, :]).reshape(-1)
    freq = np.repeat(f, 6)
    phase = (phi[:, None] + base_ph[None, :]).reshape(-1)
    ramps = np.repeat(ramp, 6)
    return dict(amp=amp, freq=freq, phase=phase, ramp=ramps)
