#!/usr/bin/env python
"""tools/collect_final.py — copy what tools/gpu_final.sh left in gpurun_out/ into profiles/ (round-2 names) and print the
headline numbers; the launch list is trimmed to start at the first estimator launch."""
import csv
import json
import shutil
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

for src, dst in (("traffic_per_launch.json", "traffic_per_launch.json"), ("final_bench.json", "r2_bench_final.json"),
                 ("final_bench_default.json", "r2_bench_default_400steps.json"), ("final_bench_ref.json", "r2_bench_reference_arm.json"),
                 ("final_sweeps.jsonl", "r2_sweeps_final.jsonl"), ("final_latency.txt", "r2_latency.txt")):
    shutil.copy(G / src, P / dst)
(P / "r2_ncu_full_fused_kernel.json").write_text(
    subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(G / "final_ncu_full_raw.csv")], capture_output=True,
                   text=True, check=True).stdout)
rows = list(csv.reader(open(G / "final_launches_ncu.csv", errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
ik = hdr.index("Kernel Name")
data = [r for r in rows if len(r) == len(hdr) and r != hdr]
first = next(i for i, r in enumerate(data) if "pmu_fused" in r[ik])
c = Counter(r[ik].split("(")[0][:60] for r in data[:first])
(P / "r2_launches_ncu.csv").write_text(
    "# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline "
    "--no-workloads --no-sustained\n# the %d launches before the first estimator launch are torch element-wise kernels building the "
    "synthetic windows in HBM: %s\n# from the first estimator launch on (3 warm-up + 4 timed launches of 8 frames, a finiteness check, then 5 + 20 "
    "single-frame launches and the same 5 + 20 with pmub_set_overlap), verbatim:\n" % (first, dict(c)) + "\n".join(",".join('"%s"' % x for x in r) for r in [hdr] + data[first:]) + "\n")
print("tree hash", bench.kernel_source_hash(), "traffic file", json.load(open(P / "traffic_per_launch.json"))["kernel_source_hash"])
for f in ("r2_bench_final.json", "r2_bench_default_400steps.json"):
    d = json.loads((P / f).read_text())
    print(f, "value %.4g frac %.3f traffic %s clocks %s" % (d["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["clocks"]["sm_mhz"]))
    for k in ("sustained", "single_frame_launch", "e2e", "e2e_stream", "e2e_stream_i16"):
        v = d.get(k) or {}
        print("   %-20s %.4g %s" % (k, v.get("value", 0), ("frac %.3f" % v["frac"]) if "frac" in v else ""))
    sd = d["stream_device"]
    print("   stream_device        %.4g (with ingest %.4g), %.0f B DRAM per estimate" % (sd["value"], sd["with_ingest"]["value"],
                                                                                     sd["roofline"].get("traffic_bytes_per_estimate") or 0))
    for k, v in d["workloads"].items():
        print("   %-28s %.4g frac %.3f" % (k, v.get("value", 0), v.get("frac", v.get("frac_of_bytes_requested", 0))))
    print("   cpu_baseline         %.4g on %d cores" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
