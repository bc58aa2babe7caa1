#!/usr/bin/env python
"""Markdown table of selected metrics from `ncu -i X.ncu-rep --page raw --csv`.  usage: ncu_summary.py raw.csv"""
import csv, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {}
for w in WANT:
    for i, h in enumerate(hdr):
        if h == w or h.endswith("." + w):
            idx[w] = i
            break
cols = [w for w in WANT if w in idx]
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
k = hdr.index("Kernel Name")
for r in rows[2:]:
    if len(r) <= k:
        continue
    print("| " + r[k].replace("void ", "").replace("(ModParams)", "") + " | " + " | ".join(f"{r[idx[c]]} {units[idx[c]]}" for c in cols) + " |")
