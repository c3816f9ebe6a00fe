"""ctypes binding of the C ABI in include/pmu_batch.h (the batched entry).

This is plumbing over ``libpmu_estimator*.so``; all arithmetic happens in the CUDA
kernel.  There is no Python or CPU estimation path here: if the shared library is
missing or no CUDA device is present, construction raises.

Host-side mirror of the reference's operator interface for this path:
``BatchEstimator.estimate`` == ``pmu_estimate`` (reference src/pmu_estimator.c:170-271)
applied to ``n_instances x n_channels`` windows at once.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_DIR = PKG / "lib"
CSRC = PKG / "csrc"

PMUB_F32, PMUB_F64 = 4, 8
SAMPLES_NATIVE, SAMPLES_F32, SAMPLES_I16 = 0, 1, 2   # pmub_stream_push_ex sample formats
PUSH_ASYNC = 1
PUSH_IN_PLACE = 2


class CoreConfig(C.Structure):
    """``pmub_config`` — estimator_config (reference src/pmu_estimator.h:30-43) in doubles."""

    _fields_ = [
        ("n_cycles", C.c_uint), ("f0", C.c_uint), ("frame_rate", C.c_uint), ("fs", C.c_uint),
        ("n_bins", C.c_uint), ("P", C.c_uint), ("Q", C.c_uint), ("iter_eipdft", C.c_uint),
        ("interf_trig", C.c_double), ("rocof_thresh", C.c_double * 3),
        ("rocof_low_pass_coeffs", C.c_double * 3),
    ]

    @classmethod
    def from_dict(cls, cfg: dict) -> "CoreConfig":
        c = cls()
        for k in ("n_cycles", "f0", "frame_rate", "fs", "n_bins", "P", "Q"):
            setattr(c, k, int(cfg[k]))
        c.iter_eipdft = 1 if cfg.get("iter_eipdft", 0) else 0
        c.interf_trig = float(cfg["interf_trig"])
        c.rocof_thresh = (C.c_double * 3)(*cfg["rocof_thresh"])
        c.rocof_low_pass_coeffs = (C.c_double * 3)(*cfg["rocof_low_pass_coeffs"])
        return c

    def to_dict(self) -> dict:
        return dict(n_cycles=self.n_cycles, f0=self.f0, frame_rate=self.frame_rate, fs=self.fs,
                    n_bins=self.n_bins, P=self.P, Q=self.Q, iter_eipdft=int(self.iter_eipdft),
                    interf_trig=self.interf_trig, rocof_thresh=list(self.rocof_thresh),
                    rocof_low_pass_coeffs=list(self.rocof_low_pass_coeffs))


class BatchInfo(C.Structure):
    """``pmub_info``."""

    _fields_ = [
        ("win_len", C.c_uint), ("n_bins", C.c_uint), ("n_instances", C.c_uint), ("n_channels", C.c_uint),
        ("dtype", C.c_int), ("device", C.c_int), ("df", C.c_double), ("norm_factor", C.c_double),
        ("dft_radix", C.c_int), ("dft_bins", C.c_int), ("n_stages", C.c_int), ("stage_bytes", C.c_int),
        ("slabs_per_window", C.c_int), ("load_mode", C.c_int), ("grid", C.c_int), ("block", C.c_int),
        ("smem_bytes", C.c_int), ("bytes_per_estimate", C.c_double),
    ]


class C37Options(C.Structure):
    """``pmub_c37_options`` (include/pmu_batch.h)."""

    _fields_ = [("format", C.c_uint), ("idcode0", C.c_uint), ("soc", C.c_uint), ("fracsec", C.c_uint), ("stat", C.c_uint),
                ("f_nominal", C.c_uint), ("phunit", C.POINTER(C.c_uint)), ("phunit_is_current", C.POINTER(C.c_ubyte)),
                ("cfg_count", C.c_uint)]


C37_POLAR, C37_FLOAT_PHASORS, C37_FLOAT_FREQ = 1, 2, 8
C37_FLOAT_POLAR = C37_POLAR | C37_FLOAT_PHASORS | C37_FLOAT_FREQ


def _c37_options(n_ch, fmt, idcode0, soc, fracsec, stat, f_nominal, phunit, is_current, cfg_count=0):
    o = C37Options()
    o.format, o.idcode0, o.soc, o.fracsec, o.stat, o.f_nominal, o.cfg_count = fmt, idcode0, soc, fracsec, stat, f_nominal, cfg_count
    keep = []
    if phunit is not None:
        arr = (C.c_uint * n_ch)(*[int(u) for u in phunit])
        o.phunit = C.cast(arr, C.POINTER(C.c_uint))
        keep.append(arr)
    if is_current is not None:
        arr = (C.c_ubyte * n_ch)(*[1 if u else 0 for u in is_current])
        o.phunit_is_current = C.cast(arr, C.POINTER(C.c_ubyte))
        keep.append(arr)
    return o, keep


def lib_path(dtype=np.float64, n_channels: int = 1) -> Path:
    """The ABI variant for (float_p, NUM_CHANLS); the pmub_* core is identical in all of them."""
    name = "libpmu_estimator"
    if np.dtype(dtype) == np.float32:
        name += "_f32"
    if n_channels == 6:
        name += "_c6"
    elif n_channels != 1:
        raise ValueError("prebuilt drop-in variants exist for NUM_CHANLS 1 and 6")
    return LIB_DIR / f"{name}.so"


def build_library(verbose: bool = False) -> None:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    jobs = str(max(2, min(8, os.cpu_count() or 2)))
    subprocess.run(["make", "-C", str(CSRC), "-j", jobs], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


class _Missing:
    """Stands in for an entry point an older build of the library does not export (A/B runs against earlier builds):
    prototypes can be declared on it, calling it fails loudly."""

    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise AttributeError(f"this build of libpmu_estimator does not export {self._name}")


class _TolerantCDLL(C.CDLL):
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name.startswith("pmub_"):
                m = _Missing(name)
                setattr(self, name, m)
                return m
            raise


_LIB = None


def load_library(path: Path | None = None) -> C.CDLL:
    """dlopen libpmu_estimator.so and declare the pmub_* prototypes.  Raises if absent."""
    global _LIB
    if path is None and _LIB is not None:
        return _LIB
    p = Path(path) if path else lib_path()
    if not p.exists():
        raise FileNotFoundError(
            f"{p} not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the estimator)")
    lib = _TolerantCDLL(str(p))
    vp, ip = C.c_void_p, C.c_int
    lib.pmub_last_error.restype = C.c_char_p
    lib.pmub_validate.argtypes = [C.POINTER(CoreConfig)]
    lib.pmub_parse_ini.argtypes = [C.c_char_p, C.POINTER(CoreConfig)]
    lib.pmub_create.argtypes = [C.POINTER(vp), C.POINTER(CoreConfig), ip, C.c_uint, C.c_uint, ip]
    lib.pmub_create_ini.argtypes = [C.POINTER(vp), C.c_char_p, ip, C.c_uint, C.c_uint, ip]
    lib.pmub_destroy.argtypes = [vp]
    lib.pmub_get_info.argtypes = [vp, C.POINTER(BatchInfo)]
    lib.pmub_estimate.argtypes = [vp, vp, vp, vp]
    lib.pmub_estimate_async.argtypes = [vp, vp, vp, vp]
    lib.pmub_estimate_frames.argtypes = [vp, ip, vp, vp, vp]
    lib.pmub_estimate_frames_async.argtypes = [vp, ip, vp, vp, vp]
    lib.pmub_stream_begin.argtypes = [vp, C.c_uint, C.c_uint, vp, C.c_ulonglong]
    lib.pmub_stream_push.argtypes = [vp, ip, vp, vp, vp]
    lib.pmub_stream_push_ex.argtypes = [vp, ip, vp, ip, vp, vp, C.c_uint, C.POINTER(C.c_ulonglong)]
    lib.pmub_stream_set_scale.argtypes = [vp, vp]
    lib.pmub_stream_wait.argtypes = [vp, C.c_ulonglong]
    lib.pmub_stream_ring_info.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.pmub_stream_ring_info.restype = ip
    lib.pmub_sync.argtypes = [vp]
    lib.pmub_set_stream.argtypes = [vp, vp]
    lib.pmub_set_overlap.argtypes = [vp, C.c_int]
    lib.pmub_stream_ring_pitch.argtypes = [vp]
    lib.pmub_stream_ring_pitch.restype = C.c_longlong
    lib.pmub_launch_count.argtypes = [vp]
    lib.pmub_launch_count.restype = C.c_longlong
    lib.pmub_reset.argtypes = [vp]
    lib.pmub_get_state.argtypes = [vp, vp, vp, vp, vp]
    lib.pmub_set_state.argtypes = [vp, vp, vp, vp, vp]
    lib.pmub_set_debug.argtypes = [vp, ip]
    lib.pmub_fft_r.argtypes = [vp, vp, vp]
    lib.pmub_fft_r.restype = ip
    lib.pmub_get_debug.argtypes = [vp, vp, vp]
    lib.pmub_sequence_components.argtypes = [ip, vp, vp, C.c_uint, C.c_uint]
    lib.pmub_c37118_frame_bytes.argtypes = [C.c_uint]
    lib.pmub_c37118_pack.argtypes = [ip, vp, vp, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    lib.pmub_c37118_frame_bytes_ex.argtypes = [C.c_uint, C.c_uint]
    lib.pmub_c37118_pack_ex.argtypes = [ip, vp, vp, C.c_uint, C.c_uint, C.POINTER(C37Options), vp]
    lib.pmub_c37118_cfg2.argtypes = [vp, C.c_uint, C.c_uint, C.POINTER(C37Options), C.c_uint, C.c_char_p,
                                     C.POINTER(C.c_char_p), C.c_uint, ip]
    lib.pmub_sequence_components_on.argtypes = [ip, vp, vp, C.c_uint, C.c_uint, vp]
    lib.pmub_peer_alloc.argtypes = [C.c_size_t, ip, C.POINTER(vp), vp]
    lib.pmub_peer_open.argtypes = [vp, ip, C.POINTER(vp)]
    lib.pmub_peer_close.argtypes = [vp]
    lib.pmub_peer_read.argtypes = [vp, vp, C.c_size_t]
    lib.pmub_peer_free.argtypes = [vp]
    for fn in ("pmub_sequence_components", "pmub_sequence_components_on", "pmub_c37118_frame_bytes", "pmub_c37118_pack",
               "pmub_c37118_frame_bytes_ex", "pmub_c37118_pack_ex", "pmub_c37118_cfg2", "pmub_peer_alloc",
               "pmub_peer_open", "pmub_peer_close", "pmub_peer_read", "pmub_peer_free"):
        getattr(lib, fn).restype = ip
    for fn in ("pmub_validate", "pmub_parse_ini", "pmub_create", "pmub_create_ini", "pmub_destroy",
               "pmub_get_info", "pmub_estimate", "pmub_estimate_async", "pmub_estimate_frames",
               "pmub_estimate_frames_async", "pmub_stream_begin", "pmub_stream_push", "pmub_stream_push_ex",
               "pmub_stream_set_scale", "pmub_stream_wait", "pmub_sync", "pmub_set_stream", "pmub_set_overlap",
               "pmub_reset", "pmub_get_state", "pmub_set_state", "pmub_set_debug", "pmub_get_debug"):
        getattr(lib, fn).restype = ip
    if path is None:
        _LIB = lib
    return lib


def last_error(lib=None) -> str:
    lib = lib or load_library()
    return (lib.pmub_last_error() or b"").decode(errors="replace")


def validate(cfg: dict) -> bool:
    """check_config_validity (reference src/pmu_estimator.c:386-449).  Host only."""
    return load_library().pmub_validate(C.byref(CoreConfig.from_dict(cfg))) == 0


def parse_ini(path: str) -> dict | None:
    """config_from_file (reference src/pmu_estimator.c:451-484).  Host only."""
    c = CoreConfig()
    if load_library().pmub_parse_ini(os.fsencode(path), C.byref(c)) != 0:
        return None
    return c.to_dict()


def _ptr(x) -> int:
    """Address of a numpy array (host) or a torch tensor (host or CUDA)."""
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


def _ingest(x):
    """numpy arrays and torch tensors pass through; any other object that speaks ``__cuda_array_interface__``
    (CuPy, Numba, RAPIDS ...) or DLPack (``__dlpack__``: JAX, CuPy, TensorFlow ...) becomes a zero-copy torch view of
    the same memory, so device buffers of other frameworks can be handed to the estimator directly (the reference's
    wrapper copies a Python list element by element, python/pmu_estimator.py:153-156)."""
    if x is None or isinstance(x, np.ndarray) or hasattr(x, "data_ptr"):
        return x
    if hasattr(x, "__cuda_array_interface__") or hasattr(x, "__dlpack__"):
        import torch

        if hasattr(x, "__cuda_array_interface__"):
            return torch.as_tensor(x, device=torch.device("cuda", torch.cuda.current_device()))
        return torch.from_dlpack(x)
    return x


class BatchEstimator:
    """``n_instances`` independent estimators x ``n_channels`` channels on one GPU.

    ``estimate(windows, mid_fracsec)`` takes windows ``[n_instances, n_channels, win_len]``
    and ``mid_fracsec [n_instances]`` as numpy arrays (host) or torch tensors (host or
    CUDA) of the batch dtype and returns frames ``[n_instances, n_channels, 4]`` =
    ``(amp, ph, freq, rocof)`` in the same kind of container.  ROCOF state for every
    (instance, channel) stays resident in HBM between calls.
    """

    def __init__(self, cfg, n_instances: int, n_channels: int = 1, dtype=np.float64, device: int = -1,
                 lib: C.CDLL | None = None):
        self.lib = lib or load_library()
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise TypeError("dtype must be float64 or float32")
        self.n_instances, self.n_channels = int(n_instances), int(n_channels)
        h = C.c_void_p()
        code = PMUB_F64 if self.dtype == np.float64 else PMUB_F32
        if isinstance(cfg, (str, os.PathLike)):
            self.cfg = parse_ini(os.fspath(cfg)) or {}
            rc = self.lib.pmub_create_ini(C.byref(h), os.fsencode(cfg), code, self.n_instances,
                                          self.n_channels, device)
        else:
            self.cfg = dict(cfg)
            rc = self.lib.pmub_create(C.byref(h), C.byref(CoreConfig.from_dict(cfg)), code,
                                      self.n_instances, self.n_channels, device)
        if rc != 0 or not h.value:
            raise RuntimeError(f"pmub_create failed: {last_error(self.lib)}")
        self._h = h
        info = BatchInfo()
        self.lib.pmub_get_info(self._h, C.byref(info))
        self.info = info
        self.win_len = int(info.win_len)
        self.n_bins = int(info.n_bins)

    # -- lifecycle ----------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.pmub_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed: {last_error(self.lib)}")

    # -- hot path -----------------------------------------------------------
    def estimate(self, windows, mid_fracsec, out=None):
        """Synchronous estimate; host or device containers (see class docstring)."""
        n, c = self.n_instances, self.n_channels
        windows, mid_fracsec, out = _ingest(windows), _ingest(mid_fracsec), _ingest(out)
        is_torch = hasattr(windows, "data_ptr")
        if is_torch:
            import torch

            tdt = torch.float64 if self.dtype == np.float64 else torch.float32
            assert windows.dtype == tdt and windows.is_contiguous() and windows.numel() == n * c * self.win_len
            assert mid_fracsec.dtype == tdt and mid_fracsec.is_contiguous() and mid_fracsec.numel() == n
            if out is None:
                out = torch.empty((n, c, 4), dtype=tdt, device=windows.device)
        else:
            windows = np.ascontiguousarray(windows, dtype=self.dtype)
            mid_fracsec = np.ascontiguousarray(mid_fracsec, dtype=self.dtype)
            assert windows.size == n * c * self.win_len and mid_fracsec.size == n
            if out is None:
                out = np.empty((n, c, 4), dtype=self.dtype)
        self._check(self.lib.pmub_estimate(self._h, _ptr(windows), _ptr(mid_fracsec), _ptr(out)), "pmub_estimate")
        return out

    def estimate_frames(self, windows, mid_fracsec, out=None):
        """``n_frames`` consecutive frames in one call: windows ``[F, n_instances, n_channels, win_len]``,
        mid_fracsec ``[F, n_instances]`` -> frames ``[F, n_instances, n_channels, 4]``.  Same results as
        F calls of :meth:`estimate`; the kernel carries the ROCOF state from frame to frame."""
        n, c = self.n_instances, self.n_channels
        windows, mid_fracsec, out = _ingest(windows), _ingest(mid_fracsec), _ingest(out)
        is_torch = hasattr(windows, "data_ptr")
        if is_torch:
            import torch

            tdt = torch.float64 if self.dtype == np.float64 else torch.float32
            F = int(windows.shape[0])
            assert windows.dtype == tdt and windows.is_contiguous() and windows.numel() == F * n * c * self.win_len
            assert mid_fracsec.dtype == tdt and mid_fracsec.is_contiguous() and mid_fracsec.numel() == F * n
            if out is None:
                out = torch.empty((F, n, c, 4), dtype=tdt, device=windows.device)
        else:
            windows = np.ascontiguousarray(windows, dtype=self.dtype)
            mid_fracsec = np.ascontiguousarray(mid_fracsec, dtype=self.dtype)
            F = int(windows.shape[0])
            assert windows.size == F * n * c * self.win_len and mid_fracsec.size == F * n
            if out is None:
                out = np.empty((F, n, c, 4), dtype=self.dtype)
        self._check(self.lib.pmub_estimate_frames(self._h, F, _ptr(windows), _ptr(mid_fracsec), _ptr(out)),
                    "pmub_estimate_frames")
        return out

    # -- streaming front end --------------------------------------------------
    def stream_begin(self, hop: int = 0, max_frames_per_push: int = 8, history=None, first_sample_index: int = 0) -> int:
        """Keep every stream's recent samples in an HBM ring; ``history`` = the ``win_len - hop``
        samples ``[n_instances, n_channels, win_len - hop]`` that precede the first push (None = zeros).
        Returns the hop in samples."""
        if history is not None and not hasattr(history, "data_ptr"):
            history = np.ascontiguousarray(history, dtype=self.dtype)
        self._check(self.lib.pmub_stream_begin(self._h, int(hop), int(max_frames_per_push),
                                               _ptr(history) if history is not None else None,
                                               int(first_sample_index)), "pmub_stream_begin")
        # hop = 0 lets the library take one frame per reporting interval: fs / frame_rate samples
        self.hop = int(hop) if hop else int(self.cfg["fs"]) // int(self.cfg["frame_rate"])
        self.max_frames_per_push = max(1, int(max_frames_per_push))
        return self.hop

    def stream_set_scale(self, scale) -> None:
        """Per-stream factors of the int16 sample format: sample = scale[stream] * count (None = 1)."""
        if scale is None:
            self._check(self.lib.pmub_stream_set_scale(self._h, None), "pmub_stream_set_scale")
            return
        sc = np.ascontiguousarray(scale, dtype=np.float64).reshape(-1)
        assert sc.size == self.n_instances * self.n_channels
        self._check(self.lib.pmub_stream_set_scale(self._h, sc.ctypes.data), "pmub_stream_set_scale")

    def stream_push(self, new_samples, mid_fracsec=None, out=None, asynchronous: bool = False):
        """``new_samples [F, n_instances, n_channels, hop]`` -> frames ``[F, n_instances, n_channels, 4]``;
        ``mid_fracsec [F, n_instances]`` or None (computed on the device from the sample counter).

        The sample format follows the dtype of ``new_samples``: the batch dtype (native), float32 into a double
        batch (promoted on the device) or int16 ADC counts (times the factors of :meth:`stream_set_scale`).
        ``asynchronous=True`` queues the push and returns ``(out, ticket)``: ``out`` (and the inputs) must stay alive
        until :meth:`stream_wait` has returned for that ticket."""
        n, c = self.n_instances, self.n_channels
        if getattr(self, "hop", None) is None:
            raise RuntimeError("stream_push before stream_begin")
        new_samples, mid_fracsec, out = _ingest(new_samples), _ingest(mid_fracsec), _ingest(out)
        is_torch = hasattr(new_samples, "data_ptr")
        F = int(new_samples.shape[0])
        if not 1 <= F <= self.max_frames_per_push:
            raise ValueError(f"{F} frames in one push; stream_begin allowed 1..{self.max_frames_per_push}")
        if is_torch:
            import torch

            tdt = torch.float64 if self.dtype == np.float64 else torch.float32
            fmt = {tdt: SAMPLES_NATIVE, torch.float32: SAMPLES_F32, torch.int16: SAMPLES_I16}.get(new_samples.dtype)
            assert fmt is not None and new_samples.is_contiguous(), "samples must be the batch dtype, float32 or int16"
            assert new_samples.numel() == F * n * c * self.hop, "new_samples must be [F, n_instances, n_channels, hop]"
            if mid_fracsec is not None:
                assert mid_fracsec.dtype == tdt and mid_fracsec.is_contiguous() and mid_fracsec.numel() == F * n
            if out is None:
                out = torch.empty((F, n, c, 4), dtype=tdt, device=new_samples.device)
        else:
            new_samples = np.asarray(new_samples)
            if new_samples.dtype == np.int16:
                fmt = SAMPLES_I16
            elif new_samples.dtype == np.float32 and self.dtype == np.float64:
                fmt = SAMPLES_F32
            else:
                fmt = SAMPLES_NATIVE
                new_samples = new_samples.astype(self.dtype, copy=False)
            new_samples = np.ascontiguousarray(new_samples)
            assert new_samples.size == F * n * c * self.hop, "new_samples must be [F, n_instances, n_channels, hop]"
            if mid_fracsec is not None:
                mid_fracsec = np.ascontiguousarray(mid_fracsec, dtype=self.dtype)
                assert mid_fracsec.size == F * n
            if out is None:
                out = np.empty((F, n, c, 4), dtype=self.dtype)
        mid_p = _ptr(mid_fracsec) if mid_fracsec is not None else None
        if fmt == SAMPLES_NATIVE and not asynchronous:
            self._check(self.lib.pmub_stream_push(self._h, F, _ptr(new_samples), mid_p, _ptr(out)), "pmub_stream_push")
            return out
        ticket = C.c_ulonglong(0)
        self._check(self.lib.pmub_stream_push_ex(self._h, F, _ptr(new_samples), fmt, mid_p, _ptr(out),
                                                 PUSH_ASYNC if asynchronous else 0, C.byref(ticket)), "pmub_stream_push_ex")
        if asynchronous:
            # the library keeps at most four pushes in flight: hold on to the buffers of the last four
            self._keepalive = (getattr(self, "_keepalive", None) or [])[-3:] + [(new_samples, mid_fracsec, out)]
            return out, int(ticket.value)
        return out

    def stream_ring_info(self):
        """-> (device address of the rings, samples per stream ring, next write position): for device-side producers that
        write the new samples into the rings themselves and then call :meth:`stream_advance`."""
        base, cap, pos = C.c_void_p(), C.c_longlong(), C.c_longlong()
        self._check(self.lib.pmub_stream_ring_info(self._h, C.byref(base), C.byref(cap), C.byref(pos)), "pmub_stream_ring_info")
        return int(base.value), int(cap.value), int(pos.value)

    def stream_ring_pitch(self) -> int:
        """Samples between consecutive ring rows (>= the ring capacity): row ``w`` starts at ``base + w * pitch``."""
        return int(self.lib.pmub_stream_ring_pitch(self._h))

    def stream_advance(self, n_frames: int, out, mid_fracsec=None, asynchronous: bool = True):
        """``n_frames`` more frames from samples the caller's device code already wrote into the rings
        (``PMUB_PUSH_IN_PLACE``): no ingest pass, the estimator launch is the only kernel."""
        out, mid_fracsec = _ingest(out), _ingest(mid_fracsec)
        ticket = C.c_ulonglong(0)
        flags = PUSH_IN_PLACE | (PUSH_ASYNC if asynchronous else 0)
        self._check(self.lib.pmub_stream_push_ex(self._h, int(n_frames), None, SAMPLES_NATIVE,
                                                 _ptr(mid_fracsec) if mid_fracsec is not None else None, _ptr(out), flags,
                                                 C.byref(ticket)), "pmub_stream_push_ex")
        return int(ticket.value)

    def stream_wait(self, ticket: int) -> None:
        """Block until the asynchronous push that returned ``ticket`` has delivered its frames."""
        self._check(self.lib.pmub_stream_wait(self._h, int(ticket)), "pmub_stream_wait")

    def estimate_frames_async(self, n_frames: int, d_windows, d_mid_fracsec, d_out) -> None:
        self._check(self.lib.pmub_estimate_frames_async(self._h, int(n_frames), _ptr(d_windows), _ptr(d_mid_fracsec),
                                                        _ptr(d_out)), "pmub_estimate_frames_async")

    def estimate_async(self, d_windows, d_mid_fracsec, d_out) -> None:
        """Device tensors only; enqueue on the batch stream and return."""
        self._check(self.lib.pmub_estimate_async(self._h, _ptr(d_windows), _ptr(d_mid_fracsec), _ptr(d_out)),
                    "pmub_estimate_async")

    def sync(self) -> None:
        self._check(self.lib.pmub_sync(self._h), "pmub_sync")

    def set_stream(self, cuda_stream: int | None) -> None:
        """Run on the caller's CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        self._check(self.lib.pmub_set_stream(self._h, C.c_void_p(cuda_stream or 0)), "pmub_set_stream")

    def set_overlap(self, enable: bool = True) -> None:
        """Let consecutive launches of this batch overlap (``pmub_set_overlap``): the caller promises that nothing enqueued
        on the batch's stream between two launches writes what the second one reads (device windows, ``mid_fracsec``,
        ring samples written in place).  The kernel then waits for its predecessor at the ROCOF step instead of at its
        top, so a launch's transforms start while the previous launch's last CTAs finish.  Same results."""
        self._check(self.lib.pmub_set_overlap(self._h, 1 if enable else 0), "pmub_set_overlap")

    @property
    def launches(self) -> int:
        return int(self.lib.pmub_launch_count(self._h))

    # -- ROCOF state (checkpoint / resume) ------------------------------------
    def reset(self) -> None:
        self._check(self.lib.pmub_reset(self._h), "pmub_reset")

    def get_state(self) -> dict:
        n = self.n_instances * self.n_channels
        st = dict(freq_old=np.empty(n), dl0=np.empty(n), dl1=np.empty(n), state=np.empty(n, dtype=np.uint8))
        self._check(self.lib.pmub_get_state(self._h, st["freq_old"].ctypes.data, st["dl0"].ctypes.data,
                                            st["dl1"].ctypes.data, st["state"].ctypes.data), "pmub_get_state")
        return st

    def set_state(self, st: dict) -> None:
        a = {k: np.ascontiguousarray(st[k], dtype=np.uint8 if k == "state" else np.float64) for k in
             ("freq_old", "dl0", "dl1", "state")}
        self._check(self.lib.pmub_set_state(self._h, a["freq_old"].ctypes.data, a["dl0"].ctypes.data,
                                            a["dl1"].ctypes.data, a["state"].ctypes.data), "pmub_set_state")

    def fft_r(self, windows) -> np.ndarray:
        """The reference's ``pmue_fft_r`` for every window of the batch: the ``n_bins`` lowest DFT bins of the samples as
        given (no window function) -> complex ``[n_instances, n_channels, n_bins]``; the ROCOF state is not touched."""
        windows = _ingest(windows)
        if not hasattr(windows, "data_ptr"):
            windows = np.ascontiguousarray(windows, dtype=self.dtype)
        out = np.empty((self.n_instances, self.n_channels, self.n_bins, 2), dtype=self.dtype)
        self._check(self.lib.pmub_fft_r(self._h, _ptr(windows), out.ctypes.data), "pmub_fft_r")
        return out[..., 0] + 1j * out[..., 1]

    # -- parity observability -----------------------------------------------
    def set_debug(self, enable: bool = True) -> None:
        self._check(self.lib.pmub_set_debug(self._h, 1 if enable else 0), "pmub_set_debug")

    def get_debug(self):
        """-> (km [n_inst, n_ch], triggered [n_inst, n_ch] bool, spectrum [n_inst, n_ch, n_bins] complex)."""
        n = self.n_instances * self.n_channels
        flags = np.empty(n, dtype=np.int32)
        spec = np.empty((n, self.n_bins, 2))
        self._check(self.lib.pmub_get_debug(self._h, flags.ctypes.data, spec.ctypes.data), "pmub_get_debug")
        shape = (self.n_instances, self.n_channels)
        return ((flags & 0xFFFF).reshape(shape), ((flags >> 16) & 1).astype(bool).reshape(shape),
                (spec[..., 0] + 1j * spec[..., 1]).reshape(shape + (self.n_bins,)))


def sequence_components(frames, lib=None, out=None, cuda_stream: int | None = None):
    """Symmetrical components of every three-phase channel group: frames ``[n_inst, n_ch, 4]`` (numpy or
    torch, host or CUDA) -> ``[n_inst, n_ch // 3, 3, 2]`` = (amp, ph) of positive, negative, zero sequence.
    ``cuda_stream`` orders the kernel on the estimator's stream (device buffers: no synchronisation)."""
    lib = lib or load_library()
    frames = _ingest(frames)
    n_inst, n_ch = int(frames.shape[0]), int(frames.shape[1])
    if hasattr(frames, "data_ptr"):
        import torch

        code = PMUB_F64 if frames.dtype == torch.float64 else PMUB_F32
        if out is None:
            out = torch.empty((n_inst, n_ch // 3, 3, 2), dtype=frames.dtype, device=frames.device)
        frames = frames.contiguous()
    else:
        frames = np.ascontiguousarray(frames)
        code = PMUB_F64 if frames.dtype == np.float64 else PMUB_F32
        if out is None:
            out = np.empty((n_inst, n_ch // 3, 3, 2), dtype=frames.dtype)
    if lib.pmub_sequence_components_on(code, _ptr(frames), _ptr(out), n_inst, n_ch, C.c_void_p(cuda_stream or 0)) != 0:
        raise RuntimeError(f"pmub_sequence_components failed: {last_error(lib)}")
    return out


def c37118_pack(frames, idcode0: int = 1, soc: int = 0, fracsec: int = 0, stat: int = 0, lib=None, fmt: int = C37_FLOAT_POLAR,
                f_nominal: int = 50, phunit=None, cuda_stream: int | None = None, out=None):
    """One IEEE C37.118.2 data frame per instance -> uint8 ``[n_inst, frame_bytes]``.  ``fmt`` is the FORMAT word
    (``C37_POLAR | C37_FLOAT_PHASORS | C37_FLOAT_FREQ`` bits); ``phunit`` the per-channel conversion factors of the
    16-bit phasor formats.  Host frames give a host array; CUDA tensors may be packed into a CUDA ``out`` tensor on
    ``cuda_stream`` without synchronising."""
    lib = lib or load_library()
    frames = _ingest(frames)
    n_inst, n_ch = int(frames.shape[0]), int(frames.shape[1])
    if hasattr(frames, "data_ptr"):
        import torch

        code = PMUB_F64 if frames.dtype == torch.float64 else PMUB_F32
        frames = frames.contiguous()
    else:
        frames = np.ascontiguousarray(frames)
        code = PMUB_F64 if frames.dtype == np.float64 else PMUB_F32
    fb = lib.pmub_c37118_frame_bytes_ex(n_ch, fmt)
    if out is None:
        out = np.empty((n_inst, fb), dtype=np.uint8)
    o, _keep = _c37_options(n_ch, fmt, idcode0, soc, fracsec, stat, f_nominal, phunit, None)
    if lib.pmub_c37118_pack_ex(code, _ptr(frames), _ptr(out), n_inst, n_ch, C.byref(o), C.c_void_p(cuda_stream or 0)) != 0:
        raise RuntimeError(f"pmub_c37118_pack_ex failed: {last_error(lib)}")
    return out


def c37118_cfg2(n_channels: int, instance: int = 0, idcode0: int = 1, soc: int = 0, fracsec: int = 0, fmt: int = C37_FLOAT_POLAR,
                f_nominal: int = 50, phunit=None, is_current=None, station_name: str | None = None, channel_names=None,
                time_base: int = 1_000_000, data_rate: int = 50, cfg_count: int = 0, lib=None) -> bytes:
    """The CFG-2 configuration frame that announces the data frames of :func:`c37118_pack` for one instance."""
    lib = lib or load_library()
    o, _keep = _c37_options(n_channels, fmt, idcode0, soc, fracsec, 0, f_nominal, phunit, is_current, cfg_count)
    names = None
    if channel_names is not None:
        names = (C.c_char_p * n_channels)(*[n.encode() for n in channel_names])
    buf = (C.c_ubyte * 65535)()
    n = lib.pmub_c37118_cfg2(buf, 65535, n_channels, C.byref(o), instance, station_name.encode() if station_name else None,
                             names, time_base, data_rate)
    if n < 0:
        raise RuntimeError(f"pmub_c37118_cfg2 failed: {last_error(lib)}")
    return bytes(buf[:n])


class RawDevicePtr:
    """A bare device address handed to the C ABI (e.g. a slice of a peer-visible buffer)."""

    def __init__(self, address: int):
        self.address = int(address)

    def data_ptr(self) -> int:
        return self.address


def peer_alloc(n_bytes: int, device: int = -1, lib=None):
    """-> (device address, 64-byte IPC handle) of a zeroed buffer other processes can open."""
    lib = lib or load_library()
    ptr = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    if lib.pmub_peer_alloc(n_bytes, device, C.byref(ptr), handle) != 0:
        raise RuntimeError(f"pmub_peer_alloc failed: {last_error(lib)}")
    return int(ptr.value), bytes(handle)


def peer_open(handle: bytes, device: int = -1, lib=None) -> int:
    lib = lib or load_library()
    ptr = C.c_void_p()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    if lib.pmub_peer_open(buf, device, C.byref(ptr)) != 0:
        raise RuntimeError(f"pmub_peer_open failed: {last_error(lib)}")
    return int(ptr.value)


def peer_read(address: int, shape, dtype, lib=None) -> np.ndarray:
    """Synchronous copy of (part of) a peer-visible buffer into a new numpy array."""
    out = np.empty(shape, dtype=dtype)
    if (lib or load_library()).pmub_peer_read(out.ctypes.data, C.c_void_p(address), out.nbytes) != 0:
        raise RuntimeError("pmub_peer_read failed")
    return out


def peer_close(address: int, lib=None) -> None:
    (lib or load_library()).pmub_peer_close(C.c_void_p(address))


def peer_free(address: int, lib=None) -> None:
    (lib or load_library()).pmub_peer_free(C.c_void_p(address))


def shard_instances(n_instances: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block of instances owned by ``rank``: rank r owns
    ``[r*n/G, (r+1)*n/G)``.  No collective is needed on the hot path."""
    lo = n_instances * rank // world_size
    hi = n_instances * (rank + 1) // world_size
    return lo, hi
