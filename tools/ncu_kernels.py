#!/usr/bin/env python3
"""Key metrics of every kernel in an ncu report as JSON lines: python tools/ncu_kernels.py rep.ncu-rep > out.jsonl"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us", "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "registers",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "smsp__inst_executed.sum": "warp_inst", "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct", "smsp__average_warp_latency_per_inst_issued.ratio": "cycles_per_issued_inst",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {"kernel": vals[hdr.index("Kernel Name")].split("(")[0], "report": sys.argv[1].split("/")[-1]}
    for i, h in enumerate(hdr):
        if h in KEYS:
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                continue
            if units[i] in UNIT:
                v *= UNIT[units[i]]
            d[KEYS[h]] = round(v, 4)
    if "dram_read" in d:
        d["dram_bytes"] = d["dram_read"] + d.get("dram_write", 0.0)
        d["dram_gbs"] = round(d["dram_bytes"] / d["duration_us"] / 1e3, 1)
    print(json.dumps(d))
