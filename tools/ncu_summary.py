#!/usr/bin/env python3
"""Compact summary (JSON) of the first kernel in an ncu report: python tools/ncu_summary.py rep.ncu-rep out.json [frames]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration", "smsp__inst_executed.sum": "warp_inst_executed",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst", "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block", "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__inst_executed.avg.per_cycle_elapsed": "ipc_per_sm",
    "smsp__average_warp_latency_per_inst_issued.ratio": "cycles_per_issued_inst", "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "sm__cycles_elapsed.max": "sm_cycles",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum": "local_load_sectors", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum": "local_store_sectors",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    frames = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")], "report": rep}
    for i, h in enumerate(hdr):
        if h in KEYS:
            v = float(vals[i].replace(",", ""))
            if units[i] in UNIT:
                v *= UNIT[units[i]]
            d[KEYS[h]] = v
    if "dram_read" in d:
        d["dram_bytes_per_launch"] = d["dram_read"] + d.get("dram_write", 0.0)
    if frames:
        d["env_frames_in_launch"] = frames
        d["thread_inst_per_env_frame"] = d["warp_inst_executed"] * d["threads_per_inst"] / frames
        d["warp_inst_per_6507_inst"] = d["thread_inst_per_env_frame"] / 6740.0
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
