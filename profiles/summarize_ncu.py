#!/usr/bin/env python
"""Dump the metrics we quote from an .ncu-rep (ncu -i ... --page raw --csv) as a small CSV.
usage: python profiles/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu_summary.csv"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
w.writerow(["kernel", "metric", "value", "unit"])
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    for i, h in enumerate(hdr):
        if h in WANT or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
            w.writerow([name, h, r[i], units[i]])
