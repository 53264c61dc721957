#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report as JSON (the numbers DESIGN.md / profiles/ quote).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [label] >> profiles/r02_ncu_summary.jsonl

One object per profiled launch: duration, DRAM bytes (= roofline.traffic), registers, achieved occupancy, active lanes
per instruction, issue utilisation, cache hit rates, local-memory (stack / spill) instructions and the top stall reasons.
"""
import csv
import json
import subprocess
import sys

WANT = {
    "time_ms": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_pct_of_peak": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_pct_of_peak": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_pct_of_peak": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "regs": "launch__registers_per_thread",
    "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warp_inst": "smsp__inst_executed.sum",
    "local_ld_inst": "smsp__sass_inst_executed_op_local_ld.sum",
    "local_st_inst": "smsp__sass_inst_executed_op_local_st.sum",
    "local_ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "local_st_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "stall_math_pipe": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_branch": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "stall_lg_throttle": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "stall_no_inst": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    rep = sys.argv[1]
    label = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        o = {"label": label, "kernel": r[col["Kernel Name"]].split("(")[0]}
        for key, metric in WANT.items():
            i = col.get(metric)
            if i is None or r[i] == "":
                continue
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if key.endswith("_bytes") or key == "time_ms":
                v *= SCALE.get(u, 1.0)
            o[key] = v
        if "dram_read_bytes" in o:
            o["dram_bytes"] = o["dram_read_bytes"] + o.get("dram_write_bytes", 0.0)
        print(json.dumps(o))


if __name__ == "__main__":
    main()
