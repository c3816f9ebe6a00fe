"""ctypes loader for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``kind="ref"``  -> oracle/_ref/libref_<f64|f32>_c<C>.so, the unmodified reference
  compiled by oracle/Makefile (present when it was built in the dev container;
  the built files travel to the GPU box, /root/reference does not).
* ``kind="port"`` -> oracle/_build/liboracle_<f64|f32>.so, oracle/pmu_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"


class OrcConfig(C.Structure):
    """Mirror of ``orc_config`` (oracle/oracle_api.h)."""

    _fields_ = [
        ("n_cycles", C.c_uint), ("f0", C.c_uint), ("frame_rate", C.c_uint), ("fs", C.c_uint),
        ("n_bins", C.c_uint), ("P", C.c_uint), ("Q", C.c_uint), ("iter_eipdft", C.c_uint),
        ("interf_trig", C.c_double), ("rocof_thresh", C.c_double * 3),
        ("rocof_low_pass_coeffs", C.c_double * 3),
    ]


def make_orc_config(cfg: dict) -> OrcConfig:
    oc = OrcConfig()
    for k in ("n_cycles", "f0", "frame_rate", "fs", "n_bins", "P", "Q"):
        setattr(oc, k, int(cfg[k]))
    oc.iter_eipdft = 1 if cfg.get("iter_eipdft", 0) else 0
    oc.interf_trig = float(cfg["interf_trig"])
    oc.rocof_thresh = (C.c_double * 3)(*cfg["rocof_thresh"])
    oc.rocof_low_pass_coeffs = (C.c_double * 3)(*cfg["rocof_low_pass_coeffs"])
    return oc


def build_oracle(quiet: bool = True) -> None:
    """Compile oracle/ (port always; _ref when /root/reference is present)."""
    subprocess.run(["make", "-C", str(ORACLE_DIR), "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def lib_path(kind: str, dtype, n_channels: int) -> Path:
    tag = "f64" if np.dtype(dtype) == np.float64 else "f32"
    if kind == "ref":
        return ORACLE_DIR / "_ref" / f"libref_{tag}_c{n_channels}.so"
    if kind == "port":
        return ORACLE_DIR / "_build" / f"liboracle_{tag}.so"
    raise ValueError(kind)


def have(kind: str, dtype=np.float64, n_channels: int = 1) -> bool:
    return lib_path(kind, dtype, n_channels).exists()


class CpuChecker:
    """Uniform driver over oracle_api.h for either implementation."""

    def __init__(self, kind: str, dtype=np.float64, n_channels: int = 1):
        self.kind = kind
        self.dtype = np.dtype(dtype)
        self.C = int(n_channels)
        path = lib_path(kind, dtype, n_channels)
        if not path.exists() and kind == "port":
            build_oracle()
        if not path.exists():
            raise FileNotFoundError(path)
        self.lib = C.CDLL(str(path))
        p = "refh_" if kind == "ref" else "orc_"
        L = self.lib
        self._open = getattr(L, p + "open"); self._open.restype = C.c_void_p
        self._open.argtypes = [C.POINTER(OrcConfig), C.c_uint]
        self._open_ini = getattr(L, p + "open_ini"); self._open_ini.restype = C.c_void_p
        self._open_ini.argtypes = [C.c_char_p, C.c_uint]
        self._close = getattr(L, p + "close"); self._close.argtypes = [C.c_void_p]
        self._close.restype = None
        self._win_len = getattr(L, p + "win_len"); self._win_len.argtypes = [C.c_void_p]
        self._win_len.restype = C.c_uint
        self._n_bins = getattr(L, p + "n_bins"); self._n_bins.argtypes = [C.c_void_p]
        self._n_bins.restype = C.c_uint
        self._norm = getattr(L, p + "norm_factor"); self._norm.argtypes = [C.c_void_p]
        self._norm.restype = C.c_double
        self._estimate = getattr(L, p + "estimate"); self._estimate.restype = C.c_int
        self._estimate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        self._last_km = getattr(L, p + "last_km"); self._last_km.argtypes = [C.c_void_p]
        self._last_km.restype = C.c_int
        self._last_spec = getattr(L, p + "last_spectrum")
        self._last_spec.argtypes = [C.c_void_p, C.c_void_p]; self._last_spec.restype = None
        self._run = getattr(L, p + "run"); self._run.restype = C.c_double
        self._run.argtypes = [C.POINTER(OrcConfig), C.c_uint, C.c_long, C.c_long, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        info = getattr(L, p + "info")
        fs, nc = C.c_int(), C.c_int()
        info(C.byref(fs), C.byref(nc))
        assert fs.value == self.dtype.itemsize, (fs.value, self.dtype)
        assert nc.value in (0, self.C), (nc.value, self.C)

    # -- single instance ---------------------------------------------------
    def open(self, cfg: dict):
        h = self._open(C.byref(make_orc_config(cfg)), self.C)
        return h  # None when the configuration is rejected

    def open_ini(self, path: str):
        return self._open_ini(os.fsencode(path), self.C)

    def close(self, h) -> None:
        self._close(h)

    def win_len(self, h) -> int:
        return int(self._win_len(h))

    def n_bins(self, h) -> int:
        return int(self._n_bins(h))

    def norm_factor(self, h) -> float:
        return float(self._norm(h))

    def estimate(self, h, windows: np.ndarray, mid_fracsec: float):
        """windows [C, N] -> (frames [C, 4] = amp, ph, freq, rocof ; return code)."""
        w = np.ascontiguousarray(windows, dtype=self.dtype)
        out = np.zeros((self.C, 4), dtype=self.dtype)
        rc = self._estimate(h, w.ctypes.data, float(mid_fracsec), out.ctypes.data)
        return out, rc

    def last_km(self, h) -> int:
        return int(self._last_km(h))

    def last_spectrum(self, h) -> np.ndarray:
        nb = self.n_bins(h)
        buf = np.zeros(2 * nb, dtype=np.float64)
        self._last_spec(h, buf.ctypes.data)
        return buf[0::2] + 1j * buf[1::2]

    # -- batch -------------------------------------------------------------
    def run(self, cfg: dict, windows: np.ndarray, mid: np.ndarray, n_threads: int = 0,
            want_km: bool = True):
        """windows [T, I, C, N], mid [T, I] -> frames [T, I, C, 4], km [T, I, C], seconds."""
        w = np.ascontiguousarray(windows, dtype=self.dtype)
        T, I, Cc, _N = w.shape
        assert Cc == self.C
        m = np.ascontiguousarray(mid, dtype=np.float64).reshape(T, I)
        frames = np.zeros((T, I, Cc, 4), dtype=self.dtype)
        km = np.full((T, I, Cc), -1, dtype=np.int32) if want_km else None
        secs = self._run(C.byref(make_orc_config(cfg)), self.C, I, T, w.ctypes.data,
                         m.ctypes.data, frames.ctypes.data,
                         km.ctypes.data if want_km else None, int(n_threads))
        if secs < 0:
            raise RuntimeError(f"{self.kind} batch run failed ({secs})")
        return frames, km, secs
