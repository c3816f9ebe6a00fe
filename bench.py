#!/usr/bin/env python
"""bench.py — phasor estimates/s (windows x channels) of the pmu_estimate() hot path.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Headline workload (BASELINE.json config 3, the one the metric is quoted on): per GPU 4096 independent
estimator instances x 6 channels, config/config.ini parameters (2048-sample window, 11 bins, P=3, Q=3,
iterations enabled), streaming consecutive windows with ROCOF state carried in HBM.  One step = one
pmub_estimate_frames call = `--frames-per-step` (8) consecutive frames of every stream = 196608 estimates =
one kernel launch.  16 distinct consecutive frames (6.4 GB) are resident in HBM and cycled, so every step
reads inputs far larger than the 126 MB L2.  Weak scaling: every rank owns its own 4096 instances, no
collective on the hot path; the optional NCCL gather of frames to rank 0 is timed separately.

Prints ONE JSON line (rank 0):
  value         device-timed throughput of the K timed steps, inputs resident in HBM
  sustained     the same loop kept running for >= 1 s (CUDA-event timed; SM clock sampled over that second)
  roofline      algorithmic bytes / launch time against the measured HBM copy bandwidth
  workloads     the other configurations, each a short device-timed run with its own roofline fraction:
                config 5 (600-sample windows, 1M/8 channels per GPU), M-class (3072 samples), the float build,
                config 3 / M-class with EVERY stream triggering the interference iterations, the reference's
                test2.c point (16 cycles @ 51.2 kHz = 16384 samples, 18 bins), M-class through the streaming front end
  single_frame_launch   one frame per kernel launch (+ `overlapped`: the same with pmub_set_overlap)
  stream_device the streaming front end with its sample rings resident in HBM, fed from one record continuous in time
                (device-timed; `with_ingest`: the record pushed as device blocks, `sustained`: the in-place replay for >= 1 s)
  e2e           the same metric through the synchronous C ABI call with pinned HOST buffers (H2D + D2H inside)
  e2e_stream    ... through the streaming front end (only the new samples cross PCIe), double and int16 ingest
  cpu_baseline  the reference CPU path on this box's host cores (a bounded sample)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from synchropmu_b200 import configs, signals  # noqa: E402

METRIC = "phasor_estimates_per_sec"
UNIT = "estimates/s"
WORKLOADS = {
    "config3": dict(cfg="config.ini", n_inst=4096, n_ch=6, desc="4096 instances x 6 ch, 2048-sample window (config.ini)"),
    "config5": dict(cfg="cfg5_600", n_inst=21846, n_ch=6, desc="131076 channels of 600-sample windows (1M/8 per GPU)"),
    "mclass": dict(cfg="m_class_config.ini", n_inst=4096, n_ch=6, desc="4096 instances x 6 ch, 3072-sample M-class window"),
    "test2_n16384": dict(cfg="test2_16cyc", n_inst=1024, n_ch=6,
                         desc="reference test2.c sweep point: 16 cycles @ 51.2 kHz = 16384-sample window, 18 bins"),
}
TRIG_TONE = [(0.1, 75.0, 0.0)]   # the 10 % interharmonic of test/test.c:60: every stream runs the Q loop
# name -> (workload, dtype, every stream triggered, frames per launch)
EXTRA_LEGS = {
    "config5_f64": ("config5", "f64", False, 8),
    "mclass_f64": ("mclass", "f64", False, 8),
    "config3_f32": ("config3", "f32", False, 8),
    "config3_f64_all_triggered": ("config3", "f64", True, 8),
    "mclass_f64_all_triggered": ("mclass", "f64", True, 8),
    "test2_n16384_f64": ("test2_n16384", "f64", False, 4),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--frames-per-step", type=int, default=8,
                    help="consecutive frames of every stream handed to one pmub_estimate_frames call (= one step)")
    ap.add_argument("--sets", type=int, default=2, help="distinct input sets (each frames-per-step frames) cycled in HBM")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the `workloads` legs (other configurations)")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--extra-steps", type=int, default=50, help="timed launches per `workloads` leg")
    ap.add_argument("--interharmonic", action="store_true", help="add a 10 %% 75 Hz tone so the Q loop triggers")
    return ap.parse_args()


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash() -> str:
    """Hash of the CUDA sources the shipped library is built from: ncu-derived constants are only valid for it."""
    h = hashlib.sha256()
    for f in sorted((ROOT / "synchropmu_b200" / "csrc").glob("*")):
        if f.suffix in (".cu", ".cuh", ".cpp", ".hpp", ".c"):
            h.update(f.name.encode())
            h.update(f.read_bytes())
    return h.hexdigest()[:16]


def measured_traffic(key: str):
    """dram bytes per launch from the committed ncu capture - only if that capture was taken on THIS kernel source
    (profiles/traffic_per_launch.json carries the source hash); otherwise None, never a stale constant."""
    f = ROOT / "profiles" / "traffic_per_launch.json"
    try:
        d = json.loads(f.read_text())
        if d.get("kernel_source_hash") == kernel_source_hash():
            return d.get(key)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while a region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

    def summary(self, t0=None, t1=None):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1):
                continue
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, n_ch, dtype, sample_inst, sample_frames, steps, warmup, interharmonic):
    """Time the reference's CPU implementation (oracle/_ref, else the C port) on a bounded sample
    of the bench workload with all host threads.  Returns (est_per_s, info dict)."""
    import oracle_lib
    from oracle_lib import CpuChecker

    npdt = np.float64 if dtype == "f64" else np.float32
    kind = "ref" if oracle_lib.have("ref", npdt, n_ch) else "port"
    if kind == "port" and not oracle_lib.have("port", npdt, n_ch):
        oracle_lib.build_oracle()
    ck = CpuChecker(kind, npdt, n_ch)
    prm = signals.three_phase_params(sample_inst)
    N = configs.win_len(cfg)
    win = np.zeros((sample_frames, sample_inst, n_ch, N), dtype=npdt)
    mid = np.zeros((sample_frames, sample_inst))
    harm = TRIG_TONE if interharmonic else []
    for t in range(sample_frames):
        win[t] = signals.steady(cfg, t, prm["amp"], prm["freq"], prm["phase"], ramp=prm["ramp"],
                                harmonics=harm).reshape(sample_inst, n_ch, N)
        mid[t] = signals.mid_fracsec(cfg, t)
    cores = os.cpu_count() or 1
    n_est = sample_frames * sample_inst * n_ch
    for _ in range(warmup):
        ck.run(cfg, win[:1], mid[:1], n_threads=cores, want_km=False)
    secs = []
    for _ in range(max(1, steps)):
        _, _, s = ck.run(cfg, win, mid, n_threads=cores, want_km=False)
        secs.append(s)
    total = sum(secs)
    value = n_est * len(secs) / total
    cpu_model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                cpu_model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    info = dict(value=value, unit=UNIT, cores=cores, kind="reference" if kind == "ref" else "port",
                sample=f"{sample_inst} instances x {n_ch} ch x {sample_frames} frames of the bench workload "
                       f"({n_est} estimates per pass, OpenMP over instances)", cpu_model=cpu_model)
    return value, info, total / len(secs) * 1e3


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank's host threads (and so, by first touch, its pinned staging buffers) to the NUMA node the GPU
    hangs off: with 8 ranks on one box the H2D copies otherwise cross the socket interconnect."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


class Bench:
    def __init__(self, args):
        import torch

        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the estimator)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa_node = bind_to_gpu_numa_node(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.stream = torch.cuda.Stream(device=self.dev)   # a real (non-default) stream: the kernel and the timing
        self.peak, self.peak_src = measured_peak()

    # -- helpers ----------------------------------------------------------------------------
    def max_over_ranks(self, v: float) -> float:
        if self.dist is None:
            return v
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def make_inputs(self, cfg, n_inst, n_ch, tdt, n_frames, harm):
        """n_frames consecutive frames of this rank's shard of the (weak-scaled) population, built in HBM."""
        torch = self.torch
        prm = signals.three_phase_params(self.world * n_inst)
        sl = slice(self.rank * n_inst * n_ch, (self.rank + 1) * n_inst * n_ch)
        pt = {k: torch.as_tensor(v[sl], device=self.dev) for k, v in prm.items()}
        N = configs.win_len(cfg)
        win = torch.empty((n_frames, n_inst, n_ch, N), dtype=tdt, device=self.dev)
        mid = torch.empty((n_frames, n_inst), dtype=tdt, device=self.dev)
        for t in range(n_frames):
            x = signals.steady(cfg, t, pt["amp"], pt["freq"], pt["phase"], xp=torch, ramp=pt["ramp"], harmonics=harm,
                               like=pt["amp"])
            win[t] = x.to(tdt).reshape(n_inst, n_ch, N)
            mid[t] = signals.mid_fracsec(cfg, t)
            del x
        return win, mid

    def make_record(self, cfg, n_inst, n_ch, tdt, first_sample, n_rows, harm, samples_per_row=None):
        """[n_rows][n_inst][n_ch][samples_per_row] consecutive stretches (default: hops) of the SAME record make_inputs cuts
        its windows from, starting at sample `first_sample`: what a streaming caller would push, hop after hop."""
        torch = self.torch
        prm = signals.three_phase_params(self.world * n_inst)
        sl = slice(self.rank * n_inst * n_ch, (self.rank + 1) * n_inst * n_ch)
        pt = {k: torch.as_tensor(v[sl], device=self.dev) for k, v in prm.items()}
        L = configs.hop(cfg) if samples_per_row is None else int(samples_per_row)
        rec = torch.empty((n_rows, n_inst, n_ch, L), dtype=tdt, device=self.dev)
        for j in range(n_rows):
            x = signals.steady(cfg, 0, pt["amp"], pt["freq"], pt["phase"], xp=torch, ramp=pt["ramp"], harmonics=harm,
                               like=pt["amp"], span=(first_sample + j * L, L))
            rec[j] = x.to(tdt).reshape(n_inst, n_ch, L)
            del x
        return rec

    def stream_leg(self, wl_name, dtype, F, n_push):
        """The streaming front end on another configuration, samples resident in the HBM rings (one record continuous in
        time, pushed once through the ingest kernel, then replayed in place: see stream_legs): {value, frac, ...}."""
        from synchropmu_b200 import BatchEstimator

        torch = self.torch
        wl = WORKLOADS[wl_name]
        cfg = configs.NAMED[wl["cfg"]]
        n_inst, n_ch = wl["n_inst"], wl["n_ch"]
        tdt = torch.float64 if dtype == "f64" else torch.float32
        npdt = np.float64 if dtype == "f64" else np.float32
        N, hop = configs.win_len(cfg), configs.hop(cfg)
        warm = max(3, -(-(N // hop - 1) // F))                # the warm-up pushes cover the replay's seam
        K = warm + n_push
        history = self.make_record(cfg, n_inst, n_ch, tdt, 0, 1, [], samples_per_row=N - hop)[0]
        rec = self.make_record(cfg, n_inst, n_ch, tdt, N - hop, K * F, [])
        d_out = torch.empty((F, n_inst, n_ch, 4), dtype=tdt, device=self.dev)
        est = BatchEstimator(cfg, n_inst, n_ch, dtype=npdt, device=self.local_rank)
        est.set_stream(self.stream.cuda_stream)
        est.stream_begin(hop=hop, max_frames_per_push=K * F, history=history)
        for i in range(warm):
            est.stream_push(rec[i * F:(i + 1) * F], None, out=d_out, asynchronous=True)
        torch.cuda.synchronize()
        ms_ingest = self.timed_launches(lambda i: est.stream_push(rec[(warm + i) * F:(warm + i + 1) * F], None, out=d_out,
                                                                  asynchronous=True), n_push)
        est.sync()
        for i in range(warm):
            est.stream_advance(F, d_out)
        torch.cuda.synchronize()
        est.reset()
        ms = self.timed_launches(lambda i: est.stream_advance(F, d_out), n_push)
        est.sync()
        finite = bool(torch.isfinite(d_out).all().item())
        n_est = n_inst * n_ch * F
        value = self.world * n_est * n_push / (ms * 1e-3)
        bpe = configs.bytes_per_estimate(cfg, 8 if dtype == "f64" else 4, n_ch)
        res = dict(value=value, unit=UNIT, frac_of_bytes_requested=value / self.world * bpe / 1e9 / self.peak, ms_per_push=ms / n_push,
                   pushes=n_push, frames_per_push=F, with_ingest=self.world * n_est * n_push / (ms_ingest * 1e-3),
                   ring_pitch_samples=est.stream_ring_pitch(), ring_capacity_samples=est.stream_ring_info()[1], win_len=N, hop=hop,
                   dtype=dtype, outputs_finite=finite, workload=wl["desc"] + ", streaming front end, samples in the HBM rings")
        est.close()
        del rec, history, d_out
        torch.cuda.empty_cache()
        return res

    def timed_launches(self, fn, steps):
        """CUDA-event time of `steps` calls of fn(i) on self.stream, max over ranks, in ms."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for i in range(steps):
            fn(i)
        e1.record(self.stream)
        torch.cuda.synchronize()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def device_leg(self, wl_name, dtype, triggered, F, steps, warmup=5):
        """A short device-timed run of another configuration: {value, frac, ...}."""
        from synchropmu_b200 import BatchEstimator

        torch = self.torch
        wl = WORKLOADS[wl_name]
        cfg = configs.NAMED[wl["cfg"]]
        n_inst, n_ch = wl["n_inst"], wl["n_ch"]
        tdt = torch.float64 if dtype == "f64" else torch.float32
        npdt = np.float64 if dtype == "f64" else np.float32
        win, mid = self.make_inputs(cfg, n_inst, n_ch, tdt, F, TRIG_TONE if triggered else [])
        out = torch.empty((F, n_inst, n_ch, 4), dtype=tdt, device=self.dev)
        est = BatchEstimator(cfg, n_inst, n_ch, dtype=npdt, device=self.local_rank)
        est.set_stream(self.stream.cuda_stream)
        est.set_debug(True)                      # one untimed launch to count how many streams really trigger
        est.estimate_frames_async(1, win[:1], mid[:1], out[:1])
        torch.cuda.synchronize()
        trig_frac = float(est.get_debug()[1].mean())
        est.set_debug(False)
        est.reset()

        def go(i):
            est.estimate_frames_async(F, win, mid, out)

        for i in range(warmup):
            go(i)
        self.barrier()
        ms = self.timed_launches(go, steps)
        n_est = n_inst * n_ch * F
        value = self.world * n_est * steps / (ms * 1e-3)
        bpe = configs.bytes_per_estimate(cfg, 8 if dtype == "f64" else 4, n_ch)
        finite = bool(torch.isfinite(out).all().item())
        info = est.info
        res = dict(value=value, unit=UNIT, frac=value / self.world * bpe / 1e9 / self.peak, ms_per_launch=ms / steps,
                   launches=steps, estimates_per_launch=n_est, bytes_per_estimate=bpe, win_len=configs.win_len(cfg),
                   n_bins=cfg["n_bins"], Q=cfg["Q"], dtype=dtype, triggered_fraction=trig_frac, outputs_finite=finite,
                   workload=wl["desc"], ring_stages=int(info.n_stages))
        est.close()
        del win, mid, out
        torch.cuda.empty_cache()
        return res


def stream_legs(B, est, cfg, n_inst, n_ch, tdt, npdt, harm, args):
    """The streaming front end (pmub_stream_*): sample rings resident in HBM, each push hands over the `hop` new samples
    of every stream.  All legs are fed from ONE record that is continuous in time (round 2 cycled two blocks of new
    samples: the seam every other push put a phase jump into every window that straddled it, the interference loop
    fired on those windows and the legs measured a partly triggered workload).  Three measurements:
      stream_device   new samples already in HBM, F frames per push, CUDA-event timed (no host copy)
      e2e_stream      pinned host samples in the batch dtype, one frame per synchronous push (as in round 1)
      e2e_stream_i16  pinned host int16 ADC counts (promoted on the device, lossless), asynchronous pushes: the copy of
                      push t+1 overlaps the kernels of push t; every push's frames are read on the host
    """
    torch = B.torch
    world = B.world
    N = configs.win_len(cfg)
    hop = configs.hop(cfg)
    itemsize = 8 if tdt == torch.float64 else 4
    F = args.frames_per_step
    n_est_frame = n_inst * n_ch
    warm = 3
    n_push = max(10, min(args.steps, 16))
    K = warm + n_push
    # history = the record's first N - hop samples; rec[j] = its next hop samples: pushing rec[j] completes frame j
    history = B.make_record(cfg, n_inst, n_ch, tdt, 0, 1, harm, samples_per_row=N - hop)[0]
    rec = B.make_record(cfg, n_inst, n_ch, tdt, N - hop, K * F, harm)            # [K*F][n_inst][n_ch][hop]

    # ---- device-resident ------------------------------------------------------------------------------------------
    # The rings are sized for the whole record (max_frames_per_push = K*F, cap = K*F + N/hop - 1 hops), so that after
    # the ingest leg has pushed it once every sample of it is resident, contiguous in time, and the in-place leg can
    # walk it a second time without a producer kernel in the timed region.
    torch.cuda.set_stream(B.stream)
    est.set_stream(B.stream.cuda_stream)
    est.reset()
    est.stream_begin(hop=hop, max_frames_per_push=K * F, history=history)
    d_out = torch.empty((F, n_inst, n_ch, 4), dtype=tdt, device=B.dev)
    for i in range(warm):
        est.stream_push(rec[i * F:(i + 1) * F], None, out=d_out, asynchronous=True)
    torch.cuda.synchronize()
    ms_ingest = B.timed_launches(lambda i: est.stream_push(rec[(warm + i) * F:(warm + i + 1) * F], None, out=d_out,
                                                           asynchronous=True), n_push)
    est.sync()
    finite_ingest = bool(torch.isfinite(d_out).all().item())
    v_ingest = world * n_est_frame * F * n_push / (ms_ingest * 1e-3)
    # ... and with the samples already IN the rings (a device-side producer writes them there itself,
    # pmub_stream_ring_info + PMUB_PUSH_IN_PLACE): the estimator launch is the only kernel.  The write position is back
    # at 0 now (cap = K*F + N/hop - 1 hops), so the next frames replay the record: N/hop - 1 windows straddle its
    # end/start seam - they fall into the warm-up pushes - then frames 0, 1, ... again.
    _, cap, wpos = est.stream_ring_info()
    ring_exact = (wpos == 0)                                # (false when hop does not divide win_len: cap is rounded up to whole hops
                                                            #  and the replay's seam drifts through the timed pushes)
    for i in range(warm):
        est.stream_advance(F, d_out)
    torch.cuda.synchronize()
    est.reset()       # ROCOF memory only (a window across the seam may have left a NaN frequency in a stream's delay line)
    launches0 = est.launches
    ms = B.timed_launches(lambda i: est.stream_advance(F, d_out), n_push)
    est.sync()
    finite = bool(torch.isfinite(d_out).all().item())
    launches = int(est.launches - launches0)
    est.set_debug(True)                                      # how many streams run the interference loop on the next frame
    est.stream_advance(1, d_out[:1])
    est.sync()
    trig_frac = float(est.get_debug()[1].mean())
    est.set_debug(False)
    v = world * n_est_frame * F * n_push / (ms * 1e-3)
    # the same launches for >= 1 s (the power cap sets the clocks): the rings are replayed over and over, so N/hop - 1 of every
    # K*F frames straddle the record's seam and run the interference loop (2 %)
    sustained = None
    if not args.no_sustained:
        n_sus = max(n_push, int(math.ceil(1200.0 / (ms / n_push))))
        ms_sus = B.timed_launches(lambda i: est.stream_advance(F, d_out), n_sus)
        est.sync()
        sustained = dict(value=world * n_est_frame * F * n_sus / (ms_sus * 1e-3), unit=UNIT, pushes=n_sus,
                         seconds=ms_sus * 1e-3, ms_per_push=ms_sus / n_sus,
                         seam_frames_fraction=(N // hop - 1) / float(K * F))
    s = itemsize
    alg = hop * s + 4 * s + 2 * (3 * s + 1)                  # new samples in, frame out, ROCOF state in and out
    moved = N * s + 4 * s + 2 * (3 * s + 1)                  # what the estimator launch requests: the whole window from its ring
    traffic = measured_traffic("stream_device_config3_f64") if (N, F, itemsize) == (2048, 8, 8) else None
    stream_device = dict(value=v, unit=UNIT, frames_per_push=F, pushes=n_push, ms_per_push=ms / n_push,
                         gpu_launches=launches, outputs_finite=finite, triggered_fraction=trig_frac, sustained=sustained,
                         record_fills_rings_exactly=ring_exact,
                         with_ingest=dict(value=v_ingest, unit=UNIT, ms_per_push=ms_ingest / n_push, outputs_finite=finite_ingest,
                                          note="the same record pushed as device blocks of F x n_streams x hop new samples, "
                                               "scattered into the rings by the ingest kernel (+ 2 x hop x %d B per estimate)" % s),
                         roofline=dict(bound="hbm", algorithmic_bytes_per_estimate=alg, frac=v / world * alg / 1e9 / B.peak,
                                       requested_bytes_per_estimate=moved, frac_of_bytes_requested=v / world * moved / 1e9 / B.peak,
                                       traffic=traffic, traffic_bytes_per_estimate=(traffic / (n_est_frame * F) if traffic else None),
                                       peak=B.peak, unit="GB/s"),
                         note="sample rings resident in HBM, the new samples already in them, one record continuous in time.  "
                              "The estimator requests each whole window from its ring, but a streaming launch walks its items "
                              "group-major (the F consecutive windows of a group of 32 streams back to back), so the "
                              "win_len - hop samples that consecutive windows share are re-read from L2: `traffic` is the DRAM "
                              "traffic of the estimator launch measured by ncu.  `frac` is against the streaming-algorithmic "
                              "bytes (hop new samples + frame + state), `frac_of_bytes_requested` against what the window-mode "
                              "launch would have read from HBM")
    del d_out

    # ---- end to end, batch dtype, synchronous, one frame per push ---------------------------------------------------
    torch.cuda.set_stream(torch.cuda.default_stream(B.dev))
    est.set_stream(None)
    est.reset()
    est.stream_begin(hop=hop, max_frames_per_push=1, history=history)
    n_push = min(args.e2e_steps * 4, K * F - 3)
    n_host = n_push + 3                                       # frames of the record staged in pinned host memory
    h_new = torch.empty((n_host, n_inst, n_ch, hop), dtype=tdt).pin_memory()
    h_new.copy_(rec[:n_host])
    h_out = [torch.empty((1, n_inst, n_ch, 4), dtype=tdt).pin_memory() for _ in range(2)]
    hn = h_new.numpy()
    ho = [h.numpy() for h in h_out]
    for i in range(3):
        est.stream_push(hn[i:i + 1], None, out=ho[0])
    B.barrier()
    t0 = time.perf_counter()
    for i in range(n_push):
        est.stream_push(hn[3 + i:4 + i], None, out=ho[0])
        _ = float(ho[0][0, 0, 0, 2])
    dt = B.max_over_ranks(time.perf_counter() - t0)
    e2e_stream = dict(value=world * n_est_frame * n_push / dt, unit=UNIT,
                      h2d_bytes_per_step=int(n_est_frame * hop * itemsize), d2h_bytes_per_step=int(n_est_frame * 4 * itemsize),
                      steps=n_push, ms_per_step=dt / n_push * 1e3, host_memory="pinned", numa_node=B.numa_node,
                      outputs_finite=bool(np.isfinite(ho[0]).all()),
                      note="pmub_stream_push (synchronous): sample rings resident in HBM, one frame per call, only the "
                           "%d new samples per stream are copied in, as %d-byte values" % (hop, itemsize))
    del h_new, hn

    # ---- end to end, int16 ADC counts, asynchronous double-buffered pushes --------------------------------------------
    est.reset()
    est.stream_begin(hop=hop, max_frames_per_push=1, history=history)
    amax = torch.maximum(history.abs().amax(dim=2), rec[:n_host].abs().amax(dim=3).amax(dim=0)).clamp_min(1e-30)   # per stream
    scale = (amax / 32000.0).double()
    est.stream_set_scale(scale.reshape(-1).cpu().numpy())
    q_new = torch.empty((n_host, n_inst, n_ch, hop), dtype=torch.int16).pin_memory()
    for j in range(n_host):
        q_new[j].copy_(torch.round(rec[j].double() / scale[:, :, None]).clamp_(-32768, 32767).to(torch.int16))
    qn = q_new.numpy()
    prev = None
    for i in range(3):
        _, tk = est.stream_push(qn[i:i + 1], None, out=ho[i % 2], asynchronous=True)
        est.stream_wait(tk)
    B.barrier()
    t0 = time.perf_counter()
    for i in range(n_push):
        _, tk = est.stream_push(qn[3 + i:4 + i], None, out=ho[i % 2], asynchronous=True)
        if prev is not None:
            est.stream_wait(prev[0])
            _ = float(ho[prev[1]][0, 0, 0, 2])     # the previous push's frames, read on the host while this one runs
        prev = (tk, i % 2)
    est.stream_wait(prev[0])
    _ = float(ho[prev[1]][0, 0, 0, 2])
    dt = B.max_over_ranks(time.perf_counter() - t0)
    finite = bool(np.isfinite(ho[0]).all() and np.isfinite(ho[1]).all())
    e2e_stream_i16 = dict(value=world * n_est_frame * n_push / dt, unit=UNIT,
                          h2d_bytes_per_step=int(n_est_frame * hop * 2), d2h_bytes_per_step=int(n_est_frame * 4 * itemsize),
                          steps=n_push, ms_per_step=dt / n_push * 1e3, host_memory="pinned", numa_node=B.numa_node,
                          outputs_finite=finite,
                          note="pmub_stream_push_ex(PMUB_SAMPLES_I16, PMUB_PUSH_ASYNC): int16 ADC counts x per-stream scale, promoted "
                               "to the batch dtype on the device (exact), consecutive hops of one pinned record; every push's frames "
                               "are read on the host one push later")
    est.stream_set_scale(None)
    del rec, history, q_new, qn
    torch.cuda.empty_cache()
    return stream_device, e2e_stream, e2e_stream_i16


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    cfg = configs.NAMED[wl["cfg"]]
    n_inst, n_ch = wl["n_inst"], wl["n_ch"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    itemsize = 8 if args.dtype == "f64" else 4
    N = configs.win_len(cfg)
    bpe = configs.bytes_per_estimate(cfg, itemsize, n_ch)
    config = dict(workload=f"{args.workload}: {wl['desc']}", n_instances_per_gpu=n_inst, n_channels=n_ch,
                  win_len=N, n_bins=cfg["n_bins"], P=cfg["P"], Q=cfg["Q"], iter_eipdft=cfg["iter_eipdft"],
                  frames_per_step=args.frames_per_step, frames_resident=args.frames_per_step * args.sets, parallelism=f"instances sharded over {world} GPU(s), no collective",
                  l2_policy="inputs larger than L2 (%.0f MB per frame, %d distinct frames cycled)" % (
                      n_inst * n_ch * N * itemsize / 1e6, args.frames_per_step * args.sets),
                  interharmonic=bool(args.interharmonic))

    # ---------------- reference arm: the CPU implementation on the host cores ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        # one step = one pass over a bounded sample (256 instances x 6 ch x 4 frames = 6144 estimates, ~0.1 s on
        # 16-24 cores); K steps are honoured up to a bound that keeps the run within a few minutes
        sample_inst = 256
        ref_steps = max(1, min(args.steps, 1500))
        value, info, ms = cpu_reference_run(cfg, n_ch, args.dtype, sample_inst, 4, ref_steps,
                                            min(args.warmup, 5), args.interharmonic)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=ref_steps, warmup=min(args.warmup, 5),
                    ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype,
                    data="synthetic", config=config, impl="reference", cpu_baseline=info,
                    e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line))
        return

    # ---------------- our arm ----------------------------------------------------------------
    B = Bench(args)
    torch, dist, dev, local_rank = B.torch, B.dist, B.dev, B.local_rank
    from synchropmu_b200 import BatchEstimator

    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    npdt = np.float64 if args.dtype == "f64" else np.float32
    harm = TRIG_TONE if args.interharmonic else []
    F = args.frames_per_step
    n_frames_res = F * args.sets
    win_all, mid_all = B.make_inputs(cfg, n_inst, n_ch, tdt, n_frames_res, harm)
    out = torch.empty((F, n_inst, n_ch, 4), dtype=tdt, device=dev)
    est = BatchEstimator(cfg, n_inst, n_ch, dtype=npdt, device=local_rank)
    stream = B.stream
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)            # events below are recorded on the same stream
    est.set_stream(stream.cuda_stream)
    n_est_step = n_inst * n_ch * F

    def step(i):
        s0 = (i % args.sets) * F
        est.estimate_frames_async(F, win_all[s0:s0 + F], mid_all[s0:s0 + F], out)

    warm = max(3, args.warmup)
    for i in range(warm):
        step(i)
    B.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    launches0 = est.launches
    t_timed0 = time.perf_counter()
    ms_total = B.timed_launches(step, args.steps)
    launches = est.launches - launches0
    ms_per_step = ms_total / args.steps
    value = world * n_est_step * args.steps / (ms_total * 1e-3)

    # ---------------- the same loop kept running for >= 1 s: what the part sustains once the power cap bites ------
    sustained = None
    t_sus0 = t_sus1 = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(math.ceil(1200.0 / ms_per_step)))
        t_sus0 = time.perf_counter()
        ms_sus = B.timed_launches(step, n_sus)
        t_sus1 = time.perf_counter()
        v_sus = world * n_est_step * n_sus / (ms_sus * 1e-3)
        sustained = dict(value=v_sus, unit=UNIT, frac=v_sus / world * bpe / 1e9 / B.peak, steps=n_sus, seconds=ms_sus * 1e-3,
                         ms_per_step=ms_sus / n_sus)
    elif rank == 0 and ms_total < 400.0:
        # no sustained leg: still keep the kernel running long enough for nvidia-smi's sampling period
        t_end = time.time() + 1.0
        i = 0
        while time.time() < t_end:
            for _ in range(200):
                step(i)
                i += 1
            torch.cuda.synchronize()
    clocks = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary(t_timed0, None)
        clocks["interval"] = "the K timed steps and the sustained loop that follows them without a gap (same kernel, same inputs)"
        if sustained is not None:
            cs = sampler.summary(t_sus0 + 0.1, t_sus1)
            sustained.update(sm_mhz=cs["sm_mhz"], reasons=cs["reasons"], clock_samples=cs["samples"])
    finite = bool(torch.isfinite(out).all().item())

    # optional gather of the frames to rank 0 over NCCL/NVLink (off the hot path, timed separately)
    gather_ms = None
    if dist is not None:
        last = out[F - 1].contiguous()
        bufs = [torch.empty_like(last) for _ in range(world)] if rank == 0 else None
        dist.gather(last, bufs, dst=0)          # untimed: the first call sets up the NCCL connections
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.gather(last, bufs, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)

    # ---------------- the same kernel launched one frame at a time (pmub_estimate_async) ----------
    single = None
    if F > 1:
        out1 = out[0]
        n1 = max(20, min(args.steps * F, 400))
        for i in range(5):
            est.estimate_async(win_all[i % n_frames_res], mid_all[i % n_frames_res], out1)
        ms1 = B.timed_launches(lambda i: est.estimate_async(win_all[i % n_frames_res], mid_all[i % n_frames_res], out1), n1) / n1
        v1 = world * n_inst * n_ch / (ms1 * 1e-3)
        single = dict(value=v1, unit=UNIT, frac=v1 / world * bpe / 1e9 / B.peak, ms_per_launch=ms1,
                      note="one frame (%d estimates) per kernel launch" % (n_inst * n_ch))
        # the same launches with pmub_set_overlap: every input is resident before the first launch, so a launch may start
        # on the SMs its predecessor has already left (the kernel waits for it at the ROCOF step only)
        est.set_overlap(True)
        for i in range(5):
            est.estimate_async(win_all[i % n_frames_res], mid_all[i % n_frames_res], out1)
        ms1o = B.timed_launches(lambda i: est.estimate_async(win_all[i % n_frames_res], mid_all[i % n_frames_res], out1), n1) / n1
        est.set_overlap(False)
        v1o = world * n_inst * n_ch / (ms1o * 1e-3)
        single["overlapped"] = dict(value=v1o, unit=UNIT, frac=v1o / world * bpe / 1e9 / B.peak, ms_per_launch=ms1o,
                                    note="pmub_set_overlap(1): consecutive launches overlap")

    # ---------------- frames written straight into rank 0's HBM by the kernel (P2P stores over NVLink,
    # CUDA IPC buffer): the gather without a collective ------------------------------------------------
    p2p = None
    if dist is not None:
        from synchropmu_b200 import api as _api

        slice_bytes = F * n_inst * n_ch * 4 * itemsize
        base, handle = (None, None)
        if rank == 0:
            base, handle = _api.peer_alloc(world * slice_bytes, local_rank)
        box = [handle]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            base = _api.peer_open(box[0], local_rank)
        remote = _api.RawDevicePtr(base + rank * slice_bytes)
        est.reset()
        for i in range(3):
            est.estimate_frames_async(F, win_all[0:F], mid_all[0:F], remote)
        torch.cuda.synchronize()
        dist.barrier()
        n2 = max(10, args.steps // 4)

        def step_remote(i):
            s0 = (i % args.sets) * F
            est.estimate_frames_async(F, win_all[s0:s0 + F], mid_all[s0:s0 + F], remote)

        tp = B.timed_launches(step_remote, n2)
        # one more pass on a known set, into rank 0 and locally, to check that the remote copy is the same data
        est.reset()
        est.estimate_frames_async(F, win_all[0:F], mid_all[0:F], remote)
        est.reset()
        est.estimate_frames_async(F, win_all[0:F], mid_all[0:F], out)
        torch.cuda.synchronize()
        mine = out.double().sum(dim=(0, 1, 2)).cpu()
        dist.barrier()
        all_sums = [None] * world
        dist.all_gather_object(all_sums, mine.tolist())
        ok = None
        if rank == 0:
            got = _api.peer_read(base, (world, F, n_inst, n_ch, 4), npdt).astype(np.float64).sum(axis=(1, 2, 3))
            ok = bool(np.allclose(got, np.array(all_sums), rtol=1e-12, atol=0))
        p2p = dict(value=world * n_est_step * n2 / (tp * 1e-3), unit=UNIT, steps=n2, verified=ok,
                   note="same launches with `frames` pointing into rank 0's HBM (CUDA IPC): frame stores cross NVLink, no collective")
        dist.barrier()
        if rank != 0:
            _api.peer_close(base)

    # ---------------- e2e: pinned host buffers through the synchronous C-ABI call ----------------
    e2e = None
    if not args.no_e2e:
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        est.set_stream(None)
        est.reset()
        host_in = [torch.empty((n_inst, n_ch, N), dtype=tdt).pin_memory() for _ in range(2)]
        for i in range(2):
            host_in[i].copy_(win_all[i % n_frames_res])
        host_mid = [mid_all[i % n_frames_res].cpu().pin_memory() for i in range(2)]
        host_out = torch.empty((n_inst, n_ch, 4), dtype=tdt).pin_memory()
        hin = [h.numpy() for h in host_in]
        hmid = [h.numpy() for h in host_mid]
        hout = host_out.numpy()
        for i in range(2):
            est.estimate(hin[i % 2], hmid[i % 2], out=hout)
        B.barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            est.estimate(hin[i % 2], hmid[i % 2], out=hout)
            _ = float(hout[0, 0, 2])                      # the step's result is read on the host
        dt = B.max_over_ranks(time.perf_counter() - t0)
        e2e = dict(value=world * n_inst * n_ch * args.e2e_steps / dt, unit=UNIT,
                   h2d_bytes_per_step=int(n_inst * n_ch * N * itemsize + n_inst * itemsize),
                   d2h_bytes_per_step=int(n_inst * n_ch * 4 * itemsize), steps=args.e2e_steps,
                   ms_per_step=dt / args.e2e_steps * 1e3, host_memory="pinned", numa_node=B.numa_node,
                   note="one frame per synchronous pmub_estimate call, H2D of the windows and D2H of the frames inside")
        del host_in, hin

    # ---------------- the streaming front end: device-resident, then end to end ----------------------------------
    stream_device = None
    e2e_stream = None
    e2e_stream_i16 = None
    del win_all, mid_all, out
    torch.cuda.empty_cache()
    if not args.no_e2e:
        stream_device, e2e_stream, e2e_stream_i16 = stream_legs(B, est, cfg, n_inst, n_ch, tdt, npdt, harm, args)

    est.close()
    torch.cuda.empty_cache()

    # ---------------- the other configurations (short runs) ---------------------------------------------------------
    workloads = None
    if not args.no_workloads:
        workloads = {}
        torch.cuda.set_stream(stream)
        for name, (wl_name, dt_name, trig, f_leg) in EXTRA_LEGS.items():
            try:
                workloads[name] = B.device_leg(wl_name, dt_name, trig, f_leg, args.extra_steps)
            except Exception as e:   # a leg that cannot run is reported, it does not take the headline down
                workloads[name] = dict(error=str(e)[:300])
        # the streaming front end on a plan with several slabs per window (mirrored rings, tensor-map tiles)
        try:
            workloads["mclass_f64_stream"] = B.stream_leg("mclass", "f64", 8, 8)
        except Exception as e:
            workloads["mclass_f64_stream"] = dict(error=str(e)[:300])

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    achieved = bpe * n_est_step / (ms_per_step * 1e-3) / 1e9
    roofline = dict(bound="hbm", achieved=achieved, peak=B.peak, unit="GB/s", frac=achieved / B.peak,
                    traffic=measured_traffic(f"{args.workload}_{args.dtype}"), traffic_source="profiles/traffic_per_launch.json "
                    "(ncu dram__bytes_read+write of one launch; used only when its kernel_source_hash matches the tree)",
                    peak_source=B.peak_src, kernel="pmu_fused_kernel", bytes_per_estimate=bpe,
                    estimates_per_launch=n_est_step, launch_ms=ms_per_step, kernel_source_hash=kernel_source_hash())
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        _, cpu_baseline, _ = cpu_reference_run(cfg, n_ch, args.dtype, 256, 4, 3, 1, args.interharmonic)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=warm,
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype,
                data="synthetic", config=config, roofline=roofline, cpu_baseline=cpu_baseline, e2e=e2e, clocks=clocks,
                gpu_launches=int(launches), outputs_finite=finite, impl="ours", sustained=sustained,
                single_frame_launch=single, stream_device=stream_device, e2e_stream=e2e_stream,
                e2e_stream_i16=e2e_stream_i16, workloads=workloads)
    if gather_ms is not None:
        line["frame_gather_ms"] = gather_ms
        line["frames_p2p_to_rank0"] = p2p
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
