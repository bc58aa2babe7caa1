#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump into the phases of the
FormerModule kernel (source-line ranges of kasf_module.cu), with the dominant stall reasons of each phase.
usage: ncu_phases.py dump.csv [kernel-substring]"""
import csv, sys
csv.field_size_limit(1 << 30)
import os
# phase = source-line range of kasf_module.cu, located by the first line containing a marker (the source must be the
# one the capture was built from)
MARKERS = [("gelu helpers", "TWICE the erf-GELU"), ("tile helpers", "fp32 [128][128] tile with XOR-swizzled"),
           ("ln_stats", "void ln_stats("), ("ln_write", "template <bool Z_TO_AUX>"),
           ("attention core", "warp-level MMAs"), ("similarity mma", "temporal adjacency"),
           ("similarity topk", "a row is spread over the 4 lanes of a quad"), ("tile geometry", "tile geometry"),
           ("gather/read_staged", "Row gather of a tile, issued by the compute warps"), ("arrive", "void warp_arrive("),
           ("setup", "__global__ void __launch_bounds__(MOD_THREADS, 1) former_module_kernel"),
           ("producer+mma warps", "service warpgroup (hands its registers"), ("limb / load + ln1", "compute warps ====="),
           ("qkv drain", "Q,K,V: TMEM -> bf16 smem"), ("attention call", "attention_core<MODE, TC>(sm"),
           ("gcn mixer", "GCN mixer ====="), ("mixer epilogue + ln2", "x1 = x + ls1 * mixer"),
           ("mlp epilogue", "MLP epilogues: hidden chunk c"), ("out epilogue", "B1|B2 are free: request the next tile"),
           ("kernel tail", "tc_fence_before();\n    __syncthreads();\n    if (warp == 0) tmem_dealloc"),
           ("launch + split path", "static int launch_one(")]
_src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kasportsformer_b200", "csrc", "kasf_module.cu")).read()
_starts = []
for name, mark in MARKERS:
    i = _src.find(mark)
    if i >= 0:
        _starts.append((_src.count("\n", 0, i) + 1, name))
_starts.sort()
BUCKETS = [(n, a, (_starts[i + 1][0] - 1) if i + 1 < len(_starts) else 10 ** 9) for i, (a, n) in enumerate(_starts)]
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
