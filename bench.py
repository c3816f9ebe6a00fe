#!/usr/bin/env python
"""bench.py — phasor estimates/s (windows x channels) of the pmu_estimate() hot path.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (BASELINE.json config 3, the one the metric is quoted on): per GPU 4096 independent
estimator instances x 6 channels, config/config.ini parameters (2048-sample window, 11 bins,
P=3, Q=3, iterations enabled), streaming consecutive windows with ROCOF state carried in HBM.
One step = one pmub_estimate_frames call = `--frames-per-step` (8) consecutive frames of every
stream = 196608 estimates = one kernel launch; `single_frame_launch` in the output is the same
kernel launched one frame (24576 estimates) at a time.  16 distinct consecutive frames (6.4 GB)
are resident in HBM and cycled, so every step reads inputs far larger than the 126 MB L2.  Weak scaling: every rank owns its own 4096 instances, no collective on the hot
path; the optional NCCL gather of frames to rank 0 is timed separately.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM;
`e2e` = the same metric through the C ABI with pinned HOST buffers (H2D and D2H inside the
timed region); `roofline` = algorithmic bytes / launch time against the measured HBM copy
bandwidth; `cpu_baseline` = the reference CPU path on this box's host cores (a bounded sample).
"""
from __future__ import annotations

import argparse
import json
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
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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
    harm = [(0.1, 75.0, 0.0)] if interharmonic else []
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


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    cfg = configs.NAMED[wl["cfg"]]
    n_inst, n_ch = wl["n_inst"], wl["n_ch"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
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
    import torch

    from synchropmu_b200 import BatchEstimator

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the estimator)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    npdt = np.float64 if args.dtype == "f64" else np.float32

    # this rank's shard of the (weak-scaled) population: world * n_inst instances in total
    prm = signals.three_phase_params(world * n_inst)
    sl = slice(rank * n_inst * n_ch, (rank + 1) * n_inst * n_ch)
    pt = {k: torch.as_tensor(v[sl], device=dev) for k, v in prm.items()}
    harm = [(0.1, 75.0, 0.0)] if args.interharmonic else []
    F = args.frames_per_step
    n_frames_res = F * args.sets
    win_all = torch.empty((n_frames_res, n_inst, n_ch, N), dtype=tdt, device=dev)
    mid_all = torch.empty((n_frames_res, n_inst), dtype=tdt, device=dev)
    for t in range(n_frames_res):
        x = signals.steady(cfg, t, pt["amp"], pt["freq"], pt["phase"], xp=torch, ramp=pt["ramp"], harmonics=harm,
                           like=pt["amp"])
        win_all[t] = x.to(tdt).reshape(n_inst, n_ch, N)
        mid_all[t] = signals.mid_fracsec(cfg, t)
        del x
    out = torch.empty((F, n_inst, n_ch, 4), dtype=tdt, device=dev)
    est = BatchEstimator(cfg, n_inst, n_ch, dtype=npdt, device=local_rank)
    stream = torch.cuda.Stream(device=dev)   # a real (non-default) stream: the kernel and the timing
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)            # events below are recorded on the same stream
    est.set_stream(stream.cuda_stream)
    n_est_step = n_inst * n_ch * F

    def step(i):
        s0 = (i % args.sets) * F
        est.estimate_frames_async(F, win_all[s0:s0 + F], mid_all[s0:s0 + F], out)

    for i in range(max(3, args.warmup)):
        step(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    launches0 = est.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step(i)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    launches = est.launches - launches0
    if dist is not None:
        tmax = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms_total = float(tmax.item())
    clocks = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
    if rank == 0:
        if ms_total < 400.0:
            # the timed region is too short for nvidia-smi's sampling period: keep the same kernel
            # running (untimed) for ~1 s so that the clocks are sampled under the same load
            t_end = time.time() + 1.0
            i = 0
            while time.time() < t_end:
                for _ in range(200):
                    step(i)
                    i += 1
                torch.cuda.synchronize()
        sampler.stop()
        clocks = sampler.summary()
    ms_per_step = ms_total / args.steps
    value = world * n_est_step * args.steps / (ms_total * 1e-3)
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
        s0e, s1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0e.record(stream)
        for i in range(n1):
            est.estimate_async(win_all[i % n_frames_res], mid_all[i % n_frames_res], out1)
        s1e.record(stream)
        torch.cuda.synchronize()
        ms1 = s0e.elapsed_time(s1e) / n1
        single = dict(value=world * n_inst * n_ch / (ms1 * 1e-3), unit=UNIT, ms_per_launch=ms1,
                      note="one frame (%d estimates) per kernel launch" % (n_inst * n_ch))

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
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for i in range(n2):
            s0 = (i % args.sets) * F
            est.estimate_frames_async(F, win_all[s0:s0 + F], mid_all[s0:s0 + F], remote)
        p1.record(stream)
        torch.cuda.synchronize()
        tp = torch.tensor([p0.elapsed_time(p1)], device=dev)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
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
        p2p = dict(value=world * n_est_step * n2 / (float(tp.item()) * 1e-3), unit=UNIT, steps=n2, verified=ok,
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
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            est.estimate(hin[i % 2], hmid[i % 2], out=hout)
            _ = float(hout[0, 0, 2])                      # the step's result is read on the host
        dt = time.perf_counter() - t0
        if dist is not None:
            tm = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dt = float(tm.item())
        e2e = dict(value=world * n_inst * n_ch * args.e2e_steps / dt, unit=UNIT,
                   h2d_bytes_per_step=int(n_inst * n_ch * N * itemsize + n_inst * itemsize),
                   d2h_bytes_per_step=int(n_inst * n_ch * 4 * itemsize), steps=args.e2e_steps,
                   ms_per_step=dt / args.e2e_steps * 1e3, host_memory="pinned",
                   note="one frame per synchronous pmub_estimate call, H2D of the windows and D2H of the frames inside")

    # ---------------- e2e through the streaming front end: only the hop new samples cross PCIe ----
    e2e_stream = None
    if not args.no_e2e:
        hop = configs.hop(cfg)
        est.stream_begin(hop=hop, max_frames_per_push=1, history=win_all[0][:, :, :N - hop].contiguous())
        h_new = [torch.empty((1, n_inst, n_ch, hop), dtype=tdt).pin_memory() for _ in range(2)]
        for i in range(2):
            h_new[i][0].copy_(win_all[(i + 1) % n_frames_res][:, :, N - hop:])
        h_out = torch.empty((1, n_inst, n_ch, 4), dtype=tdt).pin_memory()
        hn = [h.numpy() for h in h_new]
        ho = h_out.numpy()
        for i in range(2):
            est.stream_push(hn[i % 2], None, out=ho)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        n_push = args.e2e_steps * 4
        for i in range(n_push):
            est.stream_push(hn[i % 2], None, out=ho)
            _ = float(ho[0, 0, 0, 2])
        dt = time.perf_counter() - t0
        if dist is not None:
            tm = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dt = float(tm.item())
        e2e_stream = dict(value=world * n_inst * n_ch * n_push / dt, unit=UNIT,
                          h2d_bytes_per_step=int(n_inst * n_ch * hop * itemsize),
                          d2h_bytes_per_step=int(n_inst * n_ch * 4 * itemsize), steps=n_push,
                          ms_per_step=dt / n_push * 1e3, host_memory="pinned",
                          note="pmub_stream_push: sample rings resident in HBM, one frame per call, only the "
                               "%d new samples per stream are copied in" % hop)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    achieved = bpe * n_est_step / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tr_file = ROOT / "profiles" / "traffic_per_launch.json"
    if tr_file.exists():
        try:
            traffic = json.loads(tr_file.read_text()).get(f"{args.workload}_{args.dtype}")
        except Exception:
            traffic = None
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    peak_source=peak_src, kernel="pmu_fused_kernel", bytes_per_estimate=bpe,
                    estimates_per_launch=n_est_step, launch_ms=ms_per_step)
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        _, cpu_baseline, _ = cpu_reference_run(cfg, n_ch, args.dtype, 256, 4, 3, 1, args.interharmonic)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype,
                data="synthetic", config=config, roofline=roofline, cpu_baseline=cpu_baseline, e2e=e2e, clocks=clocks,
                gpu_launches=int(launches), outputs_finite=finite, impl="ours", single_frame_launch=single,
                e2e_stream=e2e_stream)
    if gather_ms is not None:
        line["frame_gather_ms"] = gather_ms
        line["frames_p2p_to_rank0"] = p2p
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
