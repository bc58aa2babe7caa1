"""Per-phase SM cycles of the fused FormerModule kernels (kasf_former_module_profiled)."""
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
for kind in ("attention", "graph", "bone"):
    for mode in ("spatial", "temporal"):
        _capi.former_module(cfg, blob, 0, kind, mode, v, xl)      # warm
        ph, tiles = _capi.former_module_phases(cfg, blob, 0, kind, mode, v, xl)
        tot = sum(x for k, x in ph.items() if not k.startswith("mmawarp_"))
        out[f"{kind}_{mode}"] = {"tiles": tiles, "cycles_per_tile": round(tot), **{k: round(x) for k, x in ph.items() if x > 0}}
        print(kind, mode, "tiles", tiles, "cyc/tile", round(tot), {k: round(x) for k, x in ph.items() if x > 0})
json.dump(out, open("gpurun_out/phases.json", "w"), indent=1)
