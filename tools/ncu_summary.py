#!/usr/bin/env python
"""tools/ncu_summary.py — the handful of numbers we read from an `ncu --page raw --csv` export, as JSON.

    python tools/ncu_summary.py gpurun_out/ncu_<tag>_<case>_raw.csv [...]  > profiles/r2_ncu_<...>.json"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__block_size": "block_size",
    "launch__grid_size": "grid_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warp_occupancy_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "smsp__sass_inst_executed_op_local_ld.sum": "local_loads",
    "smsp__sass_inst_executed_op_local_st.sum": "local_stores",
}


def summarize(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"file": path.split("/")[-1], "kernel": None, "stall_per_issue": {}}
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name":
            out["kernel"] = v
        if h in KEYS:
            try:
                out[KEYS[h]] = [float(v.replace(",", "")), u]
            except ValueError:
                out[KEYS[h]] = [v, u]
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
            out["stall_per_issue"][name] = round(float(v), 3)
    out["stall_per_issue"] = dict(sorted(out["stall_per_issue"].items(), key=lambda kv: -kv[1]))
    return out


if __name__ == "__main__":
    print(json.dumps([summarize(p) for p in sys.argv[1:]], indent=1))
