#!/usr/bin/env python
"""Summarise an .ncu-rep (via `ncu -i <rep> --page raw --csv`) into the few numbers we act on:
duration, DRAM bytes/throughput, occupancy limiters, pipe utilisation and the warp-stall breakdown.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.txt]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in data:
        print("==", r[idx["Kernel Name"]][:70], " id", r[idx["ID"]])
        for k in KEYS:
            if k in idx:
                print("   %-70s %s %s" % (k, r[idx[k]], units[idx[k]]))
        vals = sorted(((float(r[idx[h]].replace(",", "") or 0), h) for h in stall), reverse=True)[:6]
        print("   top stalls (warps per issue-active cycle): " + ", ".join(
            "%s=%.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in vals))


if __name__ == "__main__":
    main()
