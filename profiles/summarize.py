#!/usr/bin/env python3
"""Extract the metrics DESIGN.md / bench.py quote from ncu reports (run in the build container:
ncu reads .ncu-rep files without a GPU).  usage: summarize.py out.csv name=report.ncu-rep ..."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    out = csv.writer(open(sys.argv[1], "w", newline=""))
    out.writerow(["kernel", "metric", "unit", "value"])
    for arg in sys.argv[2:]:
        name, path = arg.split("=", 1)
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
        for k in KEYS:
            if k in d:
                out.writerow([name, k, d[k][0], d[k][1]])
        for h in rows[0]:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                out.writerow([name, h, d[h][0], d[h][1]])


if __name__ == "__main__":
    main()
