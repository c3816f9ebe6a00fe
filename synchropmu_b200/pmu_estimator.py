"""Single-instance Python wrapper — host-side mirror of the reference's
``python/pmu_estimator.py`` (classes ``EstimatorConfig`` :73-97, ``PMUEstimator`` :99-168:
``configure_from_ini``, ``configure_from_class``, ``estimate``, ``deinit``).

Same names, argument meaning and error behaviour (non-zero init code returned,
``estimate`` -> dict or ``None``), but the ctypes structures here are layout-accurate for
any ``NUM_CHANLS`` / ``float_p`` (the reference's mirror is an oversized approximation that
only fits the double, 1-channel build — SURVEY.md appendix B), ``estimate`` accepts numpy
arrays as well as lists, and multi-channel libraries return one dict per channel.
The reference's own wrapper also works unchanged against ``lib/libpmu_estimator.so``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import api


def _structs(ftype, n_ch: int):
    """pmu_context & friends for one ABI variant (include/pmu_estimator.h)."""

    class Phasor(C.Structure):
        _fields_ = [("amp", ftype), ("ph", ftype), ("freq", ftype)]

    class PmuFrame(C.Structure):
        _fields_ = [("synchrophasor", Phasor), ("rocof", ftype)]

    class SynchParams(C.Structure):
        _fields_ = [("win_len", C.c_uint), ("n_cycles", C.c_uint), ("f0", C.c_uint), ("frame_rate", C.c_uint),
                    ("fs", C.c_uint), ("n_bins", C.c_uint), ("P", C.c_uint), ("Q", C.c_uint),
                    ("iter_eipdft_enabled", C.c_bool), ("interf_trig", ftype), ("df", ftype),
                    ("norm_factor", ftype), ("phasor", Phasor)]

    class Buffers(C.Structure):
        _fields_ = [("Xf", C.c_void_p), ("Xi", C.c_void_p), ("dftbins", C.c_void_p),
                    ("hann_window", C.c_void_p), ("signal_windows", C.c_void_p * n_ch)]

    class HannConsts(C.Structure):
        _fields_ = [("C0", ftype), ("C1", ftype), ("C2", ftype), ("C3", ftype), ("C4", ftype),
                    ("C5", ftype * 2), ("C6", ftype * 2), ("inv_norm_factor", ftype)]

    class RocofStates(C.Structure):
        _fields_ = [("freq_old", ftype * n_ch), ("thresholds", ftype * 3), ("low_pass_coeff", ftype * 3),
                    ("delay_line", (ftype * 2) * n_ch), ("state", C.c_bool * n_ch)]

    class PmuContext(C.Structure):
        _fields_ = [("synch_params", SynchParams), ("buff_params", Buffers), ("hann_params", HannConsts),
                    ("rocof_params", RocofStates), ("pmu_initialized", C.c_bool)]

    class EstimatorConfigC(C.Structure):
        _fields_ = [("n_cycles", C.c_uint), ("f0", C.c_uint), ("frame_rate", C.c_uint), ("fs", C.c_uint),
                    ("n_bins", C.c_uint), ("P", C.c_uint), ("Q", C.c_uint), ("iter_eipdft", C.c_bool),
                    ("interf_trig", ftype), ("rocof_thresh", ftype * 3), ("rocof_low_pass_coeffs", ftype * 3)]

    return PmuFrame, PmuContext, EstimatorConfigC


class EstimatorConfig:
    """Same constructor as the reference's ``EstimatorConfig`` (python/pmu_estimator.py:86-97)."""

    def __init__(self, n_cycles, f0, frame_rate, fs, n_bins, P, Q, interf_trig, rocof_thresh,
                 rocof_low_pass_coeffs, iter_eipdft=False):
        self.n_cycles, self.f0, self.frame_rate, self.fs = n_cycles, f0, frame_rate, fs
        self.n_bins, self.P, self.Q = n_bins, P, Q
        self.iter_eipdft = bool(iter_eipdft)
        self.interf_trig = float(interf_trig)
        self.rocof_thresh = [float(v) for v in rocof_thresh]
        self.rocof_low_pass_coeffs = [float(v) for v in rocof_low_pass_coeffs]

    @classmethod
    def from_dict(cls, d: dict) -> "EstimatorConfig":
        return cls(d["n_cycles"], d["f0"], d["frame_rate"], d["fs"], d["n_bins"], d["P"], d["Q"],
                   d["interf_trig"], d["rocof_thresh"], d["rocof_low_pass_coeffs"], bool(d.get("iter_eipdft", 0)))


class PMUEstimator:
    CONFIG_FROM_INI = True
    CONFIG_FROM_STRUCT = False

    def __init__(self, lib_path=None, n_channels: int = 1, dtype=np.float64):
        self.n_channels = int(n_channels)
        self.dtype = np.dtype(dtype)
        path = lib_path or api.lib_path(self.dtype, self.n_channels)
        if not os.path.exists(str(path)):
            raise FileNotFoundError(f"Library not found at: {path}")
        ftype = C.c_double if self.dtype == np.float64 else C.c_float
        self._Frame, self._Ctx, self._Cfg = _structs(ftype, self.n_channels)
        self._ftype = ftype
        self.lib = C.CDLL(str(path))
        self.lib.pmu_init.argtypes = [C.POINTER(self._Ctx), C.c_void_p, C.c_bool]
        self.lib.pmu_init.restype = C.c_int
        self.lib.pmu_estimate.argtypes = [C.POINTER(self._Ctx), C.c_void_p, ftype, C.c_void_p]
        self.lib.pmu_estimate.restype = C.c_int
        self.lib.pmu_deinit.argtypes = [C.POINTER(self._Ctx)]
        self.lib.pmu_deinit.restype = C.c_int
        self.ctx = self._Ctx()          # caller-owned, zeroed (reference python/pmu_estimator.py:138-139)

    def __del__(self):
        try:
            if self.ctx.pmu_initialized:
                self.lib.pmu_deinit(C.byref(self.ctx))
        except Exception:
            pass

    def deinit(self):
        return self.lib.pmu_deinit(C.byref(self.ctx))

    def configure_from_ini(self, ini_file_path):
        buf = C.create_string_buffer(os.fsencode(ini_file_path))
        return self.lib.pmu_init(C.byref(self.ctx), C.cast(buf, C.c_void_p), self.CONFIG_FROM_INI)

    def configure_from_class(self, config: EstimatorConfig):
        c = self._Cfg()
        for k in ("n_cycles", "f0", "frame_rate", "fs", "n_bins", "P", "Q", "iter_eipdft", "interf_trig"):
            setattr(c, k, getattr(config, k))
        c.rocof_thresh = (self._ftype * 3)(*config.rocof_thresh)
        c.rocof_low_pass_coeffs = (self._ftype * 3)(*config.rocof_low_pass_coeffs)
        return self.lib.pmu_init(C.byref(self.ctx), C.cast(C.pointer(c), C.c_void_p), self.CONFIG_FROM_STRUCT)

    @property
    def win_len(self) -> int:
        return int(self.ctx.synch_params.win_len)

    def estimate(self, input_signal_window, mid_window_fracsec):
        """window(s): ``n_channels * win_len`` values.  Returns the reference's dict
        {amp, ph, freq, rocof} (a list of dicts when ``n_channels > 1``) or None on error."""
        x = np.ascontiguousarray(input_signal_window, dtype=self.dtype).reshape(-1)
        if self.ctx.pmu_initialized and x.size != self.n_channels * self.win_len:
            raise ValueError(f"expected {self.n_channels} x {self.win_len} samples, got {x.size}")
        frames = (self._Frame * self.n_channels)()
        rc = self.lib.pmu_estimate(C.byref(self.ctx), x.ctypes.data, self._ftype(mid_window_fracsec),
                                   C.cast(frames, C.c_void_p))
        if rc != 0:
            return None
        out = [dict(amp=f.synchrophasor.amp, ph=f.synchrophasor.ph, freq=f.synchrophasor.freq, rocof=f.rocof)
               for f in frames]
        return out[0] if self.n_channels == 1 else out
