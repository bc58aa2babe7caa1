"""Per-phase SM cycles of the two-tiles-in-flight FormerModule kernel (kasf_former_module_profiled_lt): cycles per tile
spent by the mixer group and by the MLP group in each phase (the two run concurrently: the period of a tile is about
the larger of the two sums)."""
import json, sys, torch
sys.path.insert(0, ".")
from kasportsformer_b200 import _capi, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 27
cfg = dict(n_layers=1, n_frames=T, dim_feat=128, dim_rep=512, num_heads=8, mlp_ratio=4, num_joints=17, neighbour_num=4)
dev = torch.device("cuda:0")
state = synthetic.make_state(cfg, 0, "default")
blob = _capi.pack_state(cfg, {k: v for k, v in state.items() if v.is_floating_point()}, dev)
v = torch.randn(B, T, 17, 128, device=dev)
xl = torch.randn(B, T, 17, 128, device=dev)
out = {}
MIX = ("rows_wait", "ln1", "qkv_wait_drain", "mixer_core", "epilogue_waits", "mixer_epilogue")
for kind in ("attention", "graph", "bone"):
    for mode in ("spatial", "temporal"):
        _capi.former_module(cfg, blob, 0, kind, mode, v, xl, use_limb_tiles=True)      # warm
        ph, tiles = _capi.former_module_phases_v2(cfg, blob, 0, kind, mode, v, xl)
        a = sum(x for k, x in ph.items() if k in MIX)
        b = sum(x for k, x in ph.items() if k not in MIX)
        out[f"{kind}_{mode}"] = {"tiles": tiles, "mixer_group": round(a), "mlp_group": round(b), **{k: round(x) for k, x in ph.items()}}
        print(kind, mode, "tiles", tiles, "mixer", round(a), "mlp", round(b), {k: round(x) for k, x in ph.items()})
json.dump(out, open("gpurun_out/phases_v2.json", "w"), indent=1)
