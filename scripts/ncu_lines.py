#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump by CUDA source line,
per kernel.  usage: ncu_lines.py dump.csv [top] [kernel-substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
kern, fname, hdr = "", "", None
data = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        kern = r[1]; continue
    if r[0] == "Line No":
        hdr = r; sa = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > ie and r[0].isdigit():
        try:
            data.setdefault(kern, []).append((int(r[sa]), int(r[ie]), fname, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
for k, d in data.items():
    if want not in k:
        continue
    tot = sum(x[0] for x in d) or 1
    toti = sum(x[1] for x in d) or 1
    print(f"== {k}: total samples {tot}, warp instructions {toti}")
    for x in sorted(d, reverse=True)[:top]:
        print(f"{100*x[0]/tot:5.1f}% smp {100*x[1]/toti:5.1f}% ins  {x[2]}:{x[3]:<4} {x[4]}")
