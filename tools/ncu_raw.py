#!/usr/bin/env python3
"""Key per-kernel metrics from an .ncu-rep (ncu --page raw --csv).  usage: python tools/ncu_raw.py rep.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:60])
    for w in want:
        if w in hdr:
            print("   %-62s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
