#!/usr/bin/env python
"""tools/update_traffic.py — measure dram bytes per launch with ncu and tie them to the kernel source they were measured
on (bench.py only reports `roofline.traffic` when profiles/traffic_per_launch.json carries the hash of the tree it runs).

Run on the GPU box:   python tools/update_traffic.py          (writes profiles/traffic_per_launch.json)
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"


def capture(cmd, skip, count, log):
    full = ["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", "regex:pmu_fused", "-s", str(skip), "-c", str(count),
            "--csv", "--log-file", str(log)] + cmd
    subprocess.run(full, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=ROOT)
    rows = list(csv.reader(open(log, errors="replace")))
    hdr = next(r for r in rows if "Metric Name" in r)
    per_launch = {}
    for r in rows:
        if len(r) == len(hdr) and r != hdr:
            d = dict(zip(hdr, r))
            per_launch.setdefault(d["ID"], {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    return list(per_launch.values())


def main():
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    res = {"kernel_source_hash": bench.kernel_source_hash(),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE pmu_fused_kernel launch (8 frames x 24576 streams), "
                   "`ncu --metrics ... --clock-control none`; bench.py uses these only while kernel_source_hash matches"}
    # explicit windows: the bench's timed launches (skip the warm-up launches)
    l = capture([sys.executable, "bench.py", "--steps", "4", "--warmup", "3", "--no-e2e", "--no-cpu-baseline", "--no-workloads",
                 "--no-sustained"], 4, 2, out / "traffic_explicit.csv")
    res["config3_f64"] = l[-1]["dram__bytes_read.sum"] + l[-1]["dram__bytes_write.sum"]
    res["config3_f64_detail"] = l[-1]
    res["config3_f64_algorithmic"] = 196608 * (2048 * 8 + 32 + 50 + 8 / 6)
    # streaming front end, rings in HBM, 8 frames per push
    # (3 + 6 pushes through the ingest kernel, then 3 + 6 in place: take two of the timed in-place launches)
    l = capture([sys.executable, "tools/stream_dev.py", "config.ini", "8", "6", "-", "both"], 13, 2, out / "traffic_stream.csv")
    res["stream_device_config3_f64"] = l[-1]["dram__bytes_read.sum"] + l[-1]["dram__bytes_write.sum"]
    res["stream_device_config3_f64_detail"] = l[-1]
    (ROOT / "profiles" / "traffic_per_launch.json").write_text(json.dumps(res, indent=1) + "\n")
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
