#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump into the phases of the
FormerModule kernel (source-line ranges of kasf_module.cu), with the dominant stall reasons of each phase.
usage: ncu_phases.py dump.csv [kernel-substring]"""
import csv, sys
csv.field_size_limit(1 << 30)
BUCKETS = [  # (name, first line, last line) in kasf_module.cu
    ("gelu helpers", 104, 152), ("ln_stats", 202, 222), ("ln_write", 223, 255), ("attention core", 256, 430),
    ("similarity mma", 431, 514), ("similarity topk", 515, 576), ("gather/read_staged", 578, 634),
    ("arrive", 635, 648), ("setup", 650, 688), ("producer+mma warps", 689, 775), ("limb / load", 776, 853),
    ("qkv drain", 854, 904), ("gcn aggregation", 905, 987), ("mixer epilogue + ln2", 988, 1036),
    ("mlp epilogue", 1037, 1118), ("out epilogue", 1119, 1160)]
want = sys.argv[2] if len(sys.argv) > 2 else ""
fname, kern, hdr = "", "", None
agg = {}
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        kern = r[1]; continue
    if r[0] == "Line No":
        hdr = r
        sa, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
        st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if not hdr or not r[0].isdigit() or len(r) <= ie:
        continue
    ln = int(r[0])
    if fname == "kasf_module.cu":
        b = next((n for n, a, z in BUCKETS if a <= ln <= z), "other module.cu")
    else:
        b = fname
    try:
        smp, ins = int(r[sa]), int(r[ie])
    except ValueError:
        continue
    d = agg.setdefault(kern, {}).setdefault(b, {"smp": 0, "ins": 0, "st": {}})
    d["smp"] += smp; d["ins"] += ins
    for i, h in st:
        try:
            d["st"][h] = d["st"].get(h, 0) + int(r[i])
        except ValueError:
            pass
for k, bs in agg.items():
    if want not in k:
        continue
    ts = sum(b["smp"] for b in bs.values()) or 1
    ti = sum(b["ins"] for b in bs.values()) or 1
    print(f"== {k}: samples {ts}, warp instructions {ti}")
    for n, b in sorted(bs.items(), key=lambda x: -x[1]["smp"]):
        top = sorted(b["st"].items(), key=lambda x: -x[1])[:4]
        tops = ", ".join(f"{h[6:]} {100 * v / max(b['smp'], 1):.0f}%" for h, v in top)
        print(f"  {n:24s} {100 * b['smp'] / ts:5.1f}% smp {100 * b['ins'] / ti:5.1f}% ins   {tops}")
