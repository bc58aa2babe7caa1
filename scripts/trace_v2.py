"""Event timeline of CTA 0 of the two-tiles-in-flight kernel (PROF build): python scripts/trace_v2.py attention spatial"""
import sys, torch
sys.path.insert(0, ".")
from kasportsformer_b200 import _capi, synthetic
kind, mode = sys.argv[1], sys.argv[2]
B, T = 1024, 27
cfg = dict(n_layers=1, n_frames=T, dim_feat=128, dim_rep=512, num_heads=8, mlp_ratio=4, num_joints=17, neighbour_num=4)
dev = torch.device("cuda:0")
state = synthetic.make_state(cfg, 0, "default")
blob = _capi.pack_state(cfg, {k: v for k, v in state.items() if v.is_floating_point()}, dev)
v = torch.randn(B, T, 17, 128, device=dev)
xl = torch.randn(B, T, 17, 128, device=dev)
_capi.former_module(cfg, blob, 0, kind, mode, v, xl, use_limb_tiles=True)
_capi.former_module_phases_v2(cfg, blob, 0, kind, mode, v, xl)
tr = sorted(_capi.former_module_phases_v2.trace, key=lambda e: e[1])
def name(tag):
    if tag < 6: return ["G0 rows landed", "G0 LN1 written", "G0 qkv drained", "G0 core done", "G0 epilogue waits over", "G0 epilogue done"][tag]
    if 100 <= tag < 110: return f"ISS mixer op {tag-100} triggers ready"
    if 110 <= tag < 120: return f"ISS mixer op {tag-110} issued"
    if tag == 140: return "ISS mlp fc1(0), fc1(1) issued"
    if 141 <= tag < 160: return f"ISS mlp fc2({tag-141}) + fc1({tag-139}) issued"
    if tag == 160: return "ISS mlp LN2 ready"
    if 161 <= tag < 180: return f"ISS mlp GELU piece {tag-161} ready"
    if 200 <= tag < 210: return f"G1 HFULL seen q={tag-200}"
    if 210 <= tag < 220: return f"G1 H loaded q={tag-210}"
    if 220 <= tag < 230: return f"G1 GELU stored q={tag-220}"
    if tag == 230: return "G1 x1 ready seen"
    if tag == 231: return "G1 LN2 written"
    if tag == 232: return "G1 out epilogue done"
    if 300 <= tag < 320: return f"PRODM fill op {(tag-300)//2} half {(tag-300)%2}"
    if 320 <= tag < 340: return f"PRODP fill op {tag-320}"
    return str(tag)
t0 = tr[0][1]
for tagk, t in tr:
    print(f"{t - t0:8d}  tile {tagk >> 16}  {name(tagk & 0xffff)}")
