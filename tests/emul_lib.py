"""Loader for tests/host_emul/_emul.so — the CUDA kernel's lane arithmetic compiled for the
CPU (TEST INFRASTRUCTURE; see tests/host_emul/emul.cpp).  Lets the GPU-less container check
the restructured DFT / sweep math against the oracle before any GPU time is spent."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "host_emul" / "emul.cpp"
SO = ROOT / "tests" / "host_emul" / "_emul.so"
CSRC = ROOT / "synchropmu_b200" / "csrc"


class CoreConfig(C.Structure):
    """Mirror of ``pmub_config`` (include/pmu_batch.h)."""
    _fields_ = [
        ("n_cycles", C.c_uint), ("f0", C.c_uint), ("frame_rate", C.c_uint), ("fs", C.c_uint),
        ("n_bins", C.c_uint), ("P", C.c_uint), ("Q", C.c_uint), ("iter_eipdft", C.c_uint),
        ("interf_trig", C.c_double), ("rocof_thresh", C.c_double * 3),
        ("rocof_low_pass_coeffs", C.c_double * 3),
    ]


def core_config(cfg: dict) -> CoreConfig:
    c = CoreConfig()
    for k in ("n_cycles", "f0", "frame_rate", "fs", "n_bins", "P", "Q"):
        setattr(c, k, int(cfg[k]))
    c.iter_eipdft = 1 if cfg.get("iter_eipdft", 0) else 0
    c.interf_trig = float(cfg["interf_trig"])
    c.rocof_thresh = (C.c_double * 3)(*cfg["rocof_thresh"])
    c.rocof_low_pass_coeffs = (C.c_double * 3)(*cfg["rocof_low_pass_coeffs"])
    return c


def build() -> None:
    deps = [SRC] + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("pmu_host.*"))
    if SO.exists() and all(SO.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return
    subprocess.run(["g++", "-O2", "-fPIC", "-std=c++17", "-shared", "-o", str(SO), str(SRC),
                    str(CSRC / "pmu_host.cpp")], check=True)


def load():
    build()
    lib = C.CDLL(str(SO))
    lib.emul_run.restype = C.c_int
    lib.emul_run.argtypes = [C.POINTER(CoreConfig), C.c_int, C.c_int, C.c_long, C.c_long, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emul_plan.restype = C.c_int
    lib.emul_plan.argtypes = [C.POINTER(CoreConfig), C.c_int, C.c_int, C.c_void_p]
    return lib


PLAN_FIELDS = ("M", "KC", "L", "W", "nslab", "pitch", "wp", "pk", "Kp", "n_stage", "stage_bytes", "n_cons", "n_raw", "raw_stride",
               "red_bytes", "off_raw", "off_red", "off_stage", "smem_bytes", "aligned16")


def plan(cfg: dict, dtype=np.float64, max_smem: int = 232448) -> dict | None:
    """The kernel launch plan make_plan() picks for a configuration (None if it is rejected)."""
    out = np.zeros(len(PLAN_FIELDS), dtype=np.int32)
    if load().emul_plan(C.byref(core_config(cfg)), np.dtype(dtype).itemsize, max_smem, out.ctypes.data) != 0:
        return None
    return dict(zip(PLAN_FIELDS, (int(v) for v in out)))


def run(cfg: dict, windows: np.ndarray, mid: np.ndarray, dtype=np.float64, n_channels: int = 1):
    """windows [T, S, N], mid [T, S // C] -> frames [T, S, 4], flags [T, S], spectrum [T, S, K], plan."""
    lib = load()
    w = np.ascontiguousarray(windows, dtype=dtype)
    T, S, _N = w.shape
    m = np.ascontiguousarray(mid, dtype=np.float64).reshape(T, S // n_channels)
    K = int(cfg["n_bins"])
    frames = np.zeros((T, S, 4))
    flags = np.zeros((T, S), dtype=np.int32)
    spec = np.zeros((T, S, K, 2))
    plan = np.zeros(5, dtype=np.int32)
    rc = lib.emul_run(C.byref(core_config(cfg)), np.dtype(dtype).itemsize, n_channels, S, T, w.ctypes.data,
                      m.ctypes.data, frames.ctypes.data, flags.ctypes.data, spec.ctypes.data, plan.ctypes.data)
    if rc != 0:
        raise RuntimeError("emul_run rejected the configuration")
    return frames, flags, spec[..., 0] + 1j * spec[..., 1], dict(M=int(plan[0]), KC=int(plan[1]),
                                                                   nslab=int(plan[2]), W=int(plan[3]), wp=int(plan[4]))
