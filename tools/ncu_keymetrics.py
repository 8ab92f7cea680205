#!/usr/bin/env python3
"""Compact key,value CSV of the metrics the roofline discussion uses, from an .ncu-rep
(ncu --set full).  usage: tools/ncu_keymetrics.py REPORT.ncu-rep [KERNEL_INDEX] > profiles/NAME_metrics.csv
(KERNEL_INDEX: which captured launch, default the last)"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + int(sys.argv[2])] if len(sys.argv) > 2 else rows[-1]
KEYS = ("Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__icc_requests_lookup_hit", "sm__icc_requests_lookup_miss",
        "smsp__average_warp", "smsp__average_warps_issue_stalled", "smsp__warp_issue_stalled", "local_", "sm__sass_inst_executed_op_shared",
        "smsp__sass_thread_inst_executed_op_dfma", "smsp__sass_thread_inst_executed_op_dmul", "smsp__sass_thread_inst_executed_op_dadd")
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit", "value"])
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in KEYS) and v not in ("", "n/a"):
        w.writerow([h, u, v])
