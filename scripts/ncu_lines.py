#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu_lines.py dump.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        sa = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
        continue
    if hdr and len(r) > ie and r[0].isdigit():
        try:
            data.append((int(r[sa]), int(r[ie]), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
toti = sum(d[1] for d in data) or 1
print(f"total samples {tot}, warp instructions {toti}")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100*d[0]/tot:5.1f}% smp {100*d[1]/toti:5.1f}% ins  {d[2]}:{d[3]:<4} {d[4]}")
